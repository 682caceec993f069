# round 2, session 28: 512- and 1024-thread blocks on scenes A and B
mkdir -p gpurun_out
{
python tools/ab_jit_opts.py c5 "-DPT_BLOCK=1024"
python tools/ab_jit_opts.py c2 - "-DPT_BLOCK=512" "-DPT_BLOCK=1024"
python tools/ab_jit_opts.py c1 - "-DPT_BLOCK=512"
python tools/ab_jit_opts.py c3 - "-DPT_BLOCK=512"
python tools/ab_jit_opts.py c2b - "-DPT_BLOCK=512"
AB_WORLD=8 python tools/ab_jit_opts.py c5 - "-DPT_BLOCK=512"
} > gpurun_out/s28_ab.log 2>&1
cat gpurun_out/s28_ab.log
