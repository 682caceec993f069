# round 2, session 26: sample runs without reservation near the end of the indices: tails and thresholds
mkdir -p gpurun_out
python -m pytest tests/test_gpu_production.py -m gpu -q -x -k "run_length or exactly_spp or sharded or reproducible or beyond" > gpurun_out/s26_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s26_pytest.log
{
echo "== c5 full (run 128)"; python tools/ab_jit_opts.py c5 -
echo "== c5 full reserve 1"; PTB200_RUN_RESERVE=1 python tools/ab_jit_opts.py c5 -
echo "== c5 full reserve 1000 (always reserve)"; PTB200_RUN_RESERVE=0 python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 run 64 (default)"; AB_WORLD=2 python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 run 128"; AB_WORLD=2 PTB200_RUN=128 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 run 1 (default)"; AB_WORLD=4 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 run 64"; AB_WORLD=4 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 run 128"; AB_WORLD=4 PTB200_RUN=128 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 run 1 (default)"; AB_WORLD=8 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 run 64"; AB_WORLD=8 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 run 32"; AB_WORLD=8 PTB200_RUN=32 python tools/ab_jit_opts.py c5 -
echo "== c2 run 16"; PTB200_RUN=16 python tools/ab_jit_opts.py c2 -
echo "== c2 run 1"; python tools/ab_jit_opts.py c2 -
} > gpurun_out/s26_ab.log 2>&1
cat gpurun_out/s26_ab.log
