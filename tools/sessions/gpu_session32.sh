# round 2, session 32: lockstep x block size (C4 without lockstep in big blocks; scene A with lockstep in 1024-thread blocks)
mkdir -p gpurun_out
{
python tools/ab_jit_opts.py c4 - "-DPT_NO_LOCKSTEP -DPT_BLOCK=1024" "-DPT_NO_LOCKSTEP -DPT_BLOCK=512"
python tools/ab_jit_opts.py c5 - "-DPT_LOCKSTEP"
python tools/ab_jit_opts.py c2 - "-DPT_LOCKSTEP"
} > gpurun_out/s32_ab.log 2>&1
cat gpurun_out/s32_ab.log
