python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "minimizer_window or all_k or sharded" 2>&1 | tail -4
KMER_CUDA_DEBUG_W=6 python tools/part_experiment.py 1000000 2>&1 | tail -1
