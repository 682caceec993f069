set -x
PTB200_JIT_VERBOSE=1 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s7.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu_s7.log
python tools/abtest.py > gpurun_out/abtest_s7.log 2>&1; cat gpurun_out/abtest_s7.log
