for w in extract match_c4 match_c5; do python bench.py --workload $w --no-cpu > gpurun_out/bench_$w.log 2>&1; tail -1 gpurun_out/bench_$w.log | cut -c1-700; echo; done
