# round 2, session 1: packed-FP32 microbenchmark, GPU tests, the restructured bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
./expt/ubench_f32x2 > gpurun_out/s1_ubench_f32x2.log 2>&1; echo "ubench rc=$?"
python -m pytest tests -m gpu -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s1_pytest.log
python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/s1_bench.err
cat gpurun_out/s1_ubench_f32x2.log
