set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_s12.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_s12.log
python tools/abtest.py > gpurun_out/abtest_s12.log 2>&1; cat gpurun_out/abtest_s12.log
