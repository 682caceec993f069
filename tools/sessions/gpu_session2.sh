# round 2, session 2: fixed packed-FP32 microbenchmark, GPU tests incl. the mirror/glass suite, utilisation timelines, bench
mkdir -p gpurun_out
./expt/ubench_f32x2 > gpurun_out/s2_ubench_f32x2.log 2>&1; echo "ubench rc=$?"
python -m pytest tests -m gpu -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s2_pytest.log
for wl in c1 c3 c1b; do python tools/timeline.py $wl 40 >> gpurun_out/s2_timeline.log 2>&1; done
PTB200_DUMP_LAUNCHES=gpurun_out/s2_launches_c3.txt python tools/timeline.py c3 40 > /dev/null 2>&1
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/s2_bench.err
cat gpurun_out/s2_ubench_f32x2.log; cat gpurun_out/s2_timeline.log
