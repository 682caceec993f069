python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" > gpurun_out/pytest_gpu_s28.log; echo "pytest rc=$?"; grep "identical\|passed\|failed" gpurun_out/pytest_gpu_s28.log | tail -12
python tools/abtest.py 2>&1 | head -2
