python tools/run_configs.py > gpurun_out/configs_v7.log 2>&1; echo "configs rc=$?"; tail -10 gpurun_out/configs_v7.log | cut -c1-200
