set -x
python tools/sweep_jit_opts.py > gpurun_out/sweep_jit_opts_s17.log 2>&1; cat gpurun_out/sweep_jit_opts_s17.log
