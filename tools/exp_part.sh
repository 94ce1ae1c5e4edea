for d in 0 8; do KMER_CUDA_DEBUG_PARTITION=$d python tools/part_experiment.py 1000000 2>&1 | tail -1 | cut -c1-160; done
