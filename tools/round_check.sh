# end-of-milestone check on the GPU box: full GPU test suite, smoke, bench (both arms), ncu launch list
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
python bench.py > gpurun_out/bench_1g.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -c 600 gpurun_out/pytest.log gpurun_out/smoke.log
