python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "count or partition or kat_group or sharded or large or empty or split or window" 2>&1 | tail -3
python tools/part_experiment.py 1000000 2>&1 | tail -1
