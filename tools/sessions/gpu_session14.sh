# round 2, session 14: single-segment sample runs (reverted from the two-segment trial): full tests, tail depth, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s14_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s14_pytest.log
{
echo "== c5 full"; python tools/ab_jit_opts.py c5 -
echo "== c5 full iters_tail 16"; PTB200_ITERS_TAIL=16 python tools/ab_jit_opts.py c5 -
echo "== c5 full iters_tail 32"; PTB200_ITERS_TAIL=32 python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 share"; AB_WORLD=2 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share"; AB_WORLD=8 python tools/ab_jit_opts.py c5 -
echo "== c2"; python tools/ab_jit_opts.py c2 -
echo "== c4"; python tools/ab_jit_opts.py c4 -
} > gpurun_out/s14_ab.log 2>&1
cat gpurun_out/s14_ab.log
python bench.py > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s14_bench.err
