python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or count or partition or sharded or large" 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/bench_1g.log 2>&1; tail -1 gpurun_out/bench_1g.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms']); print(d['e2e'])"
