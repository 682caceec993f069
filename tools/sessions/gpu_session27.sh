# round 2, session 27: threads per block on scene A (warps of a block wait for its slowest warp at the end-of-launch barrier)
mkdir -p gpurun_out
{
python tools/ab_jit_opts.py c5 - "-DPT_BLOCK=128" "-DPT_BLOCK=64" "-DPT_BLOCK=512"
python tools/ab_jit_opts.py c2 - "-DPT_BLOCK=128" "-DPT_BLOCK=64"
python tools/ab_jit_opts.py c1 - "-DPT_BLOCK=128"
} > gpurun_out/s27_ab.log 2>&1
cat gpurun_out/s27_ab.log
PTB200_JIT_OPTS="-DPT_DEBUG_BOUNDS" timeout 600 python tools/sanitize_target.py > gpurun_out/s27_debug_bounds.log 2>&1; echo "debug bounds rc=$?"; tail -2 gpurun_out/s27_debug_bounds.log
