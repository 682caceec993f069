python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu_s52.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu_s52.log
