# headline config and C3 on N GPUs (one gpurun --gpus N call).  usage: bash tools/run_ngpu_short.sh N [c3]
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29911 bench.py --gpus $N --no-cpu > gpurun_out/bench_${N}gpu_v18_c2.log 2> gpurun_out/bench_${N}gpu_v18_c2.err
[ "$2" = "c3" ] && timeout 300 $TR --master-port 29912 bench.py --gpus $N --config c3 --no-cpu > gpurun_out/bench_${N}gpu_v18_c3.log 2> gpurun_out/bench_${N}gpu_v18_c3.err
python - <<PY
import json
for f in ("c2","c3"):
    try:
        d=json.loads(open(f"gpurun_out/bench_${N}gpu_v18_{f}.log").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, round(d["ms_per_step"],2), "%.4g"%d["value"], {k:round(v,2) for k,v in (d.get("phases_ms") or {}).items()}, "tier2", d["config"]["tier2_kmers"], "rl_step", round(d["roofline_step"]["frac"],3), {k:v for k,v in d["parity"].items() if k.endswith("_ok")})
        print("   e2e", "%.4g"%e.get("value",0), round(e.get("ms_per_step",0),1), e.get("d2h_copy_gb_per_s_per_rank"), e.get("parity"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
