# round 2, session 24: 512-thread blocks for the lockstep scenes: tests, C4
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/s24_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s24_pytest.log
{
python tools/ab_jit_opts.py c4 -
PTB200_NO_BLOCK512=1 python tools/ab_jit_opts.py c4 -
python tools/ab_jit_opts.py c5 -
} > gpurun_out/s24_ab.log 2>&1
cat gpurun_out/s24_ab.log
