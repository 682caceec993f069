export PTB200_JIT_KEEP_SRC=/root/repo/gpurun_out/pt_kernel_jit_c2.cu
ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 24 -c 1 -o gpurun_out/prof_s40_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_f_s40.log 2>&1
tail -2 gpurun_out/ncu_f_s40.log | cut -c1-200
