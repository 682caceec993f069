# round 2, session 19: film coordinates at the camera ray (run mode), refine_hit without the IEEE division sequence
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/s19_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s19_pytest.log
{
echo "== c5"; python tools/ab_jit_opts.py c5 - "-DPT_REFINE_IEEE_DIV"
echo "== c2"; python tools/ab_jit_opts.py c2 - "-DPT_REFINE_IEEE_DIV"
echo "== c1"; python tools/ab_jit_opts.py c1 - "-DPT_REFINE_IEEE_DIV"
echo "== c5 1/8"; AB_WORLD=8 python tools/ab_jit_opts.py c5 -
} > gpurun_out/s19_ab.log 2>&1
cat gpurun_out/s19_ab.log
