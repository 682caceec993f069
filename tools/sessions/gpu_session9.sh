# round 2, session 9: shadow rays without slot codes; sample runs vs slots in flight; one-eighth share of C5
mkdir -p gpurun_out
python -m pytest tests/test_gpu_production.py tests/test_gpu_units.py -m gpu -q -x > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s9_pytest.log
{
echo "== shadow raw on/off, run default"; python tools/ab_jit_opts.py c5 - ; PTB200_NO_SHADOW_RAW=1 python tools/ab_jit_opts.py c5 -
echo "== c2 run 1 shadow raw on/off"; PTB200_RUN=1 python tools/ab_jit_opts.py c2 - ; PTB200_RUN=1 PTB200_NO_SHADOW_RAW=1 python tools/ab_jit_opts.py c2 -
echo "== c5 one wave x 3072 bounces"; AB_CAP=151552 AB_ITERS=3072 PTB200_ITERS_DRAIN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 two waves x 1536 bounces"; AB_CAP=303104 AB_ITERS=1536 PTB200_ITERS_DRAIN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share, default"; AB_WORLD=8 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share, run 1"; AB_WORLD=8 PTB200_RUN=1 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share, run 64"; AB_WORLD=8 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share, one wave x 3072, run 64"; AB_WORLD=8 AB_CAP=151552 AB_ITERS=3072 PTB200_ITERS_DRAIN=64 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share, two waves x 1536, run 64"; AB_WORLD=8 AB_CAP=303104 AB_ITERS=1536 PTB200_ITERS_DRAIN=64 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
} > gpurun_out/s9_ab.log 2>&1
cat gpurun_out/s9_ab.log
