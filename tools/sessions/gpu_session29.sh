# round 2, session 29: 1024-thread blocks on every config without a lockstep scan
mkdir -p gpurun_out
{
python tools/ab_jit_opts.py c1 - "-DPT_BLOCK=1024"
python tools/ab_jit_opts.py c3 - "-DPT_BLOCK=1024"
python tools/ab_jit_opts.py c1b - "-DPT_BLOCK=1024"
python tools/ab_jit_opts.py c2b - "-DPT_BLOCK=1024"
python tools/ab_jit_opts.py c3b - "-DPT_BLOCK=1024"
AB_WORLD=8 python tools/ab_jit_opts.py c5 - "-DPT_BLOCK=1024"
AB_WORLD=2 python tools/ab_jit_opts.py c5 - "-DPT_BLOCK=1024"
} > gpurun_out/s29_ab.log 2>&1
cat gpurun_out/s29_ab.log
