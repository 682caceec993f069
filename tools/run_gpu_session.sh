python -m pytest tests/test_gpu_validate.py -m gpu -q -x -k full_size --durations=3 2>&1 | tail -8
