python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/abtest.py 2>&1 | tee gpurun_out/abtest_v8.txt
