python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s44.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_s44.log
python tools/abtest.py > gpurun_out/abtest_s44.log 2>&1; cat gpurun_out/abtest_s44.log
