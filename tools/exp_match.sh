python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "match" 2>&1 | tail -3
python bench.py --workload match_c4 --no-cpu 2>&1 | tail -1 | cut -c1-330
