# round 2, session 25: final state - all GPU tests, smoke, the bench line, the generic (ahead-of-time) kernel on C2
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s25_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s25_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s25_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s25_smoke.log
python tools/ab_jit_opts.py c2 generic - > gpurun_out/s25_ab.log 2>&1; cat gpurun_out/s25_ab.log
python bench.py > gpurun_out/s25_bench.json 2> gpurun_out/s25_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s25_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s25_bench_ref.json 2> gpurun_out/s25_bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/s25_bench_ref.json | cut -c1-400
