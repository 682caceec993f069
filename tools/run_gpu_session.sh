python -m pytest tests/test_gpu_production.py -m gpu -q -x -k full_size_c5 --durations=3 2>&1 | tail -12
