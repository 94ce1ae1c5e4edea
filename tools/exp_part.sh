python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sharded_count or partition or all_k" 2>&1 | tail -3
for d in 0; do KMER_CUDA_DEBUG_PARTITION=$d python tools/part_experiment.py 1000000 2>&1 | tail -1; done
