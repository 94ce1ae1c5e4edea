for k in 14 15; do KMER_ALGO=1 KMER_K=$k python tools/part_experiment.py 1000000 2>&1 | tail -1; done
for k in 15 16 18 20 27 28; do KMER_K=$k python tools/part_experiment.py 1000000 2>&1 | tail -1; done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "all_k or sharded or tiers or overflow or split" 2>&1 | tail -3
