python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiers or overflow or all_k or boundaries or window or repetitive" --durations=5 2>&1 | tail -12
python tools/part_experiment.py 1000000 2>&1 | tail -1
