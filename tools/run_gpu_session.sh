python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s24.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_s24.log
