# round 2, session 23: C4 - block size with the lockstep scan (instruction-cache sharing across more warps)
mkdir -p gpurun_out
{
python tools/ab_jit_opts.py c4 - "-DPT_BLOCK=512" "-DPT_BLOCK=1024" "-DPT_BLOCK=128"
} > gpurun_out/s23_ab.log 2>&1
cat gpurun_out/s23_ab.log
