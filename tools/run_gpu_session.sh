timeout 120 python -m pytest tests/test_gpu_production.py -m gpu -x -q -k "exactly_spp" 2>&1 | tail -8
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
