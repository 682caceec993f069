set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s6.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_s6.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s6_c2.json 2> gpurun_out/bench_s6_c2.err; echo "bench rc=$?"; cat gpurun_out/bench_s6_c2.json
python bench.py --steps 2 --warmup 3 --workload c5 --no-cpu-baseline > gpurun_out/bench_s6_c5.json 2> gpurun_out/bench_s6_c5.err; cat gpurun_out/bench_s6_c5.json
PTB200_WAVES=2,4,8 PTB200_ITERS=8,16,32,64 PTB200_BPS=4 python tools/sweep2.py > gpurun_out/sweep2_s6.log 2>&1; cat gpurun_out/sweep2_s6.log
