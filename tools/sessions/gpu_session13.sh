# round 2, session 13: runs + single samples at the end (two segments)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_production.py tests/test_gpu_sharding.py tests/test_gpu_glossy.py -m gpu -q -x > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s13_pytest.log
{
echo "== c5 full"; python tools/ab_jit_opts.py c5 -
echo "== c5 full tail_runs 1.0"; PTB200_RUN_TAIL=1.0 python tools/ab_jit_opts.py c5 -
echo "== c5 full tail_runs 2.5"; PTB200_RUN_TAIL=2.5 python tools/ab_jit_opts.py c5 -
echo "== c5 full iters_tail 4"; PTB200_ITERS_TAIL=4 python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 share"; AB_WORLD=2 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 share"; AB_WORLD=4 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share (run 64)"; AB_WORLD=8 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share run 128"; AB_WORLD=8 PTB200_RUN=128 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share run 64 tail 2.5"; AB_WORLD=8 PTB200_RUN_TAIL=2.5 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share run 32"; AB_WORLD=8 PTB200_RUN=32 python tools/ab_jit_opts.py c5 -
echo "== c2"; python tools/ab_jit_opts.py c2 -
echo "== c4"; python tools/ab_jit_opts.py c4 -
} > gpurun_out/s13_ab.log 2>&1
cat gpurun_out/s13_ab.log
