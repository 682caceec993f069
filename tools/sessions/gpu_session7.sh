# round 2, session 7: -fmad=false in both builds (explicit fmaf everywhere): tests, A/B, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s7_pytest.log
python tools/ab_jit_opts.py c5 - > gpurun_out/s7_ab.log 2>&1
python tools/ab_jit_opts.py c2 - "--fmad=true" >> gpurun_out/s7_ab.log 2>&1
python tools/ab_jit_opts.py c4 - >> gpurun_out/s7_ab.log 2>&1
cat gpurun_out/s7_ab.log
python bench.py > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s7_bench.err
