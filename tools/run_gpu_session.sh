set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s5.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_s5.log
PTB200_WAVES=2,4 PTB200_ITERS=1,8,32 python tools/sweep2.py c4/8,c2 > gpurun_out/sweep2_s5.log 2>&1; cat gpurun_out/sweep2_s5.log
PTB200_WAVES=2,4 PTB200_ITERS=1,8,32 PTB200_LIB=expt/libptb200_occ4.so PTB200_BPS=4 python tools/sweep2.py c4/8,c2 > gpurun_out/sweep2_s5_occ4.log 2>&1; cat gpurun_out/sweep2_s5_occ4.log
ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 12 -c 1 -o gpurun_out/prof_s5_c4 -f python bench.py --steps 1 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/ncu_s5_c4.log 2>&1
