python -m pytest tests/test_bench_contract.py -m gpu -q -x 2>&1 | tail -5
