python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s29.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_s29.log
