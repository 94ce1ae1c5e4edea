for tg in 1100 1300 1400; do KMER_CUDA_BUCKET_KMERS=$tg python tools/part_experiment.py 1000000 2>&1 | tail -1; done
