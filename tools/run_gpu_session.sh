python tools/e2e_breakdown.py
VIEW=1 python tools/e2e_breakdown.py
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s26.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_s26.log
