# round 2, session 33: pt_readback_owned by tile height (one GPU, one-eighth share of a 4K image); final GPU tests of the committed state
mkdir -p gpurun_out
python tools/readback_owned_timing.py > gpurun_out/s33_readback_owned.log 2>&1; cat gpurun_out/s33_readback_owned.log
python -m pytest tests -m gpu -q -x > gpurun_out/s33_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s33_pytest.log
