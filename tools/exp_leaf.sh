python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "count or partition or kat_group or sharded or large or empty" 2>&1 | tail -3
for tg in 1200 1100; do KMER_CUDA_BUCKET_KMERS=$tg python tools/part_experiment.py 1000000 2>&1 | tail -1; done
