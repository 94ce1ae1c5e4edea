# round-2 multi-GPU measurements (one gpurun --gpus N call): headline config, C3, the 10 GB k=21 shape, and the headline without
# NUMA binding of the host side.  usage: bash tools/run_8gpu.sh N
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29901 bench.py --gpus $N > gpurun_out/bench_${N}gpu_c2.log 2> gpurun_out/bench_${N}gpu_c2.err
timeout 300 $TR --master-port 29902 bench.py --gpus $N --config c3 --no-cpu > gpurun_out/bench_${N}gpu_c3.log 2> gpurun_out/bench_${N}gpu_c3.err
timeout 300 $TR --master-port 29903 bench.py --gpus $N --config c2x10 --no-cpu > gpurun_out/bench_${N}gpu_c2x10.log 2> gpurun_out/bench_${N}gpu_c2x10.err
KMER_NO_BIND=1 timeout 300 $TR --master-port 29904 bench.py --gpus $N --no-cpu --steps 5 > gpurun_out/bench_${N}gpu_c2_nobind.log 2> gpurun_out/bench_${N}gpu_c2_nobind.err
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
lscpu | grep -i "numa\|socket\|model name" > gpurun_out/lscpu_${N}gpu.txt 2>&1
python - <<PY
import json
for f in ("c2","c3","c2x10","c2_nobind"):
    try:
        d=json.loads(open(f"gpurun_out/bench_${N}gpu_{f}.log").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, round(d["ms_per_step"],2), "%.3g"%d["value"], {k:round(v,2) for k,v in (d.get("phases_ms") or {}).items()}, "rl_step", round(d["roofline_step"]["frac"],3), "parity", {k:v for k,v in d["parity"].items() if k.endswith("_ok")})
        print("   e2e", "%.3g"%e.get("value",0), round(e.get("ms_per_step",0),1), e.get("d2h_copy_gb_per_s_per_rank"), e.get("host_numa_node_per_rank"), e.get("host_bound_to_gpu_node_per_rank"), e.get("parity"))
    except Exception as ex:
        print(f, "ERR", ex)
        import subprocess; print(subprocess.run(["tail","-c","600",f"gpurun_out/bench_${N}gpu_{f}.err"],capture_output=True,text=True).stdout)
PY
