for nb in 1024 16384 131072; do KMER_CUDA_DEBUG_NBUCKETS=$nb timeout 120 python tools/part_experiment.py 1000000 2>&1 | tail -1 | cut -c1-200; done
