python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v18.log 2>gpurun_out/bench_ref_v18.err
python bench.py > gpurun_out/bench_1g_v18.log 2>gpurun_out/bench_1g_v18.err; tail -c 300 gpurun_out/bench_1g_v18.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v18.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bucket_count|partition_kernel" -c 2 -o gpurun_out/prof_r02_v18_full_1g python tools/part_experiment.py 1000000 > gpurun_out/ncu_v18_full.log 2>&1; tail -1 gpurun_out/ncu_v18_full.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_1g_v18.log").read().strip().splitlines()[-1])
e=d["e2e"]
print(round(d["ms_per_step"],3), "%.4g"%d["value"], d["phases_ms"], d["roofline"]["frac"], d["roofline_step"]["frac"], d["config"]["tier2_kmers"], d["parity"].get("checksum_ok"), d["parity"].get("unique_ok"))
print("e2e", "%.4g"%e["value"], e["ms_per_step"], e["parity"], "pageable", (e.get("pageable_input") or {}).get("ms_per_step"))
print("secondary", {k:(v.get("frac") if isinstance(v,dict) else v) for k,v in (d.get("secondary") or {}).items()})
print(open("gpurun_out/bench_ref_v18.log").read()[:600])
PY
