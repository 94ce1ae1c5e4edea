# one `ncu --set full` capture of the two count kernels at the bench size (1 GB, k=21); DRAM bytes per launch -> profiles/
ncu --set full --clock-control none --import-source on -k regex:'bucket_count|partition_kernel' -c 2 -o gpurun_out/prof_r01_v11_full_1g python tools/part_experiment.py 1000000 > gpurun_out/ncu_v11_full.log 2>&1
tail -2 gpurun_out/ncu_v11_full.log
