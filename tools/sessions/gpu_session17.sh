# round 2, session 17: scene-specialised modules built for the render's regeneration flags: tests, A/B, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s17_pytest.log
{
echo "== c5 full"; python tools/ab_jit_opts.py c5 -
echo "== c5 full, no render flags"; PTB200_JIT_NO_RENDER_FLAGS=1 python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 share"; AB_WORLD=2 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 share"; AB_WORLD=4 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share"; AB_WORLD=8 python tools/ab_jit_opts.py c5 -
echo "== c2"; python tools/ab_jit_opts.py c2 -
echo "== c1"; python tools/ab_jit_opts.py c1 -
echo "== c3"; python tools/ab_jit_opts.py c3 -
echo "== c4"; python tools/ab_jit_opts.py c4 -
echo "== c2b"; python tools/ab_jit_opts.py c2b -
} > gpurun_out/s17_ab.log 2>&1
cat gpurun_out/s17_ab.log
python bench.py > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s17_bench.err
