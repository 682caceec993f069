python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s31.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_s31.log
