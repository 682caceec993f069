# round 2, session 6: row-block path order v2, lockstep for big sphere tables: tests, A/B on one box, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s6_pytest.log
PTB200_BLOCK_PIXELS=100000000 python tools/ab_jit_opts.py c5 - > gpurun_out/s6_ab.log 2>&1
python tools/ab_jit_opts.py c5 - >> gpurun_out/s6_ab.log 2>&1
PTB200_BLOCK_PIXELS=2097152 python tools/ab_jit_opts.py c5 - >> gpurun_out/s6_ab.log 2>&1
python tools/ab_jit_opts.py c2 - >> gpurun_out/s6_ab.log 2>&1
python tools/ab_jit_opts.py c4 - "-DPT_NO_LOCKSTEP" >> gpurun_out/s6_ab.log 2>&1
cat gpurun_out/s6_ab.log
python bench.py > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s6_bench.err
