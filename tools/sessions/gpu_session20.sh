# round 2, session 20: hit point as o + d*t (no FMA) for all three components vs the select form
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/s20_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s20_pytest.log
{
echo "== c5"; python tools/ab_jit_opts.py c5 - "-DPT_REFINE_SELECT_X"
echo "== c2"; python tools/ab_jit_opts.py c2 - "-DPT_REFINE_SELECT_X"
echo "== c1"; python tools/ab_jit_opts.py c1 - "-DPT_REFINE_SELECT_X"
} > gpurun_out/s20_ab.log 2>&1
cat gpurun_out/s20_ab.log
