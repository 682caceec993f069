# round 2, session 12: guided chunks down to single runs near the end; run thresholds
mkdir -p gpurun_out
python -m pytest tests/test_gpu_production.py -m gpu -q -x -k "run_length or exactly_spp or sharded or reproducible" > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s12_pytest.log
{
echo "== c5 full"; python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 share"; AB_WORLD=2 python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 share run 128"; AB_WORLD=2 PTB200_RUN=128 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 share run 64"; AB_WORLD=4 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 share run 128"; AB_WORLD=4 PTB200_RUN=128 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share run 64"; AB_WORLD=8 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share run 128"; AB_WORLD=8 PTB200_RUN=128 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share run 1"; AB_WORLD=8 PTB200_RUN=1 python tools/ab_jit_opts.py c5 -
echo "== c2 run 64"; PTB200_RUN=64 python tools/ab_jit_opts.py c2 -
echo "== c2 run 1"; PTB200_RUN=1 python tools/ab_jit_opts.py c2 -
} > gpurun_out/s12_ab.log 2>&1
cat gpurun_out/s12_ab.log
