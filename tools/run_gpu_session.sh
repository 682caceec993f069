set -x
python tools/sweep_tail.py > gpurun_out/sweep_tail_s9.log 2>&1; cat gpurun_out/sweep_tail_s9.log
