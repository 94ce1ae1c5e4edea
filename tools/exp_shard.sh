python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sharded or count or partition" 2>&1 | tail -4
