# round 2, session 16: regeneration flags as compile-time constants (potential of specialising on render parameters); own bounds checks
mkdir -p gpurun_out
PTB200_JIT_OPTS="-DPT_DEBUG_BOUNDS" timeout 600 python tools/sanitize_target.py > gpurun_out/s16_debug_bounds.log 2>&1; echo "debug bounds rc=$?"; tail -3 gpurun_out/s16_debug_bounds.log
{
echo "== c2"; python tools/ab_jit_opts.py c2 - "-DPT_BAKE_WRAP_ONCE=1 -DPT_BAKE_MAGIC=1 -DPT_BAKE_WORLD1=1 -DPT_NO_ROW_BLOCKS -DPT_BAKE_RUNS=0"
echo "== c5 1/8 share"; AB_WORLD=8 python tools/ab_jit_opts.py c5 - "-DPT_BAKE_WRAP_ONCE=1 -DPT_BAKE_MAGIC=1 -DPT_BAKE_WORLD1=0 -DPT_BAKE_RUNS=0"
echo "== c5"; python tools/ab_jit_opts.py c5 - "-DPT_BAKE_WRAP_ONCE=1 -DPT_BAKE_MAGIC=1 -DPT_BAKE_WORLD1=1 -DPT_BAKE_RUNS=1"
echo "== c1"; python tools/ab_jit_opts.py c1 - "-DPT_BAKE_WRAP_ONCE=1 -DPT_BAKE_MAGIC=1 -DPT_BAKE_WORLD1=1 -DPT_NO_ROW_BLOCKS -DPT_BAKE_RUNS=0"
} > gpurun_out/s16_ab.log 2>&1
cat gpurun_out/s16_ab.log
