for d in 0 3; do KMER_CUDA_DEBUG_PARTITION=$d python tools/part_experiment.py 1000000 2>&1 | tail -1; done
