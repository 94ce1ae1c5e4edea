set -x
for d in 0 8 16 32 4 24 48; do KMER_CUDA_DEBUG_PARTITION=$d python tools/part_experiment.py 1000000 2>&1 | tail -1; done > gpurun_out/leaf_exp_v6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bucket_count|partition_kernel' -c 2 -o gpurun_out/prof_r01_v6 python tools/part_experiment.py 200000 > gpurun_out/ncu_v6.log 2>&1
