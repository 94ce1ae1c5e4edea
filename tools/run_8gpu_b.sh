N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29921 bench.py --gpus $N --no-cpu --no-e2e > gpurun_out/bench_8gpu_v18_c2.log 2> gpurun_out/bench_8gpu_v18_c2.err
timeout 200 $TR --master-port 29922 bench.py --gpus $N --config c3 --no-cpu --no-e2e --steps 10 > gpurun_out/bench_8gpu_v18_c3.log 2> gpurun_out/bench_8gpu_v18_c3.err
python - <<PY
import json
for f in ("c2","c3"):
    try:
        d=json.loads(open(f"gpurun_out/bench_8gpu_v18_{f}.log").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],2), "%.4g"%d["value"], {k:round(v,2) for k,v in (d.get("phases_ms") or {}).items()}, "tier2", d["config"]["tier2_kmers"], "rl_step", round(d["roofline_step"]["frac"],3), {k:v for k,v in d["parity"].items() if k.endswith("_ok")})
    except Exception as ex:
        print(f, "ERR", ex)
PY
