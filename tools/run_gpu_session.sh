for i in 1 2 3; do python tools/repro_seq.py A:16:1 B:16:1; echo "== A16,B16 rc=$?"; done
python tools/repro_seq.py --peak A:16:1 B:16:1 A:1:0 B:2:1; echo "== seq rc=$?"
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s21.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_s21.log
python tools/run_configs.py > gpurun_out/configs_s21.json 2> gpurun_out/configs_s21.log; echo "rc=$?"; tail -12 gpurun_out/configs_s21.log
