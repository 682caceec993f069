# round 2: ncu evidence for the CURRENT kernels (one gpurun call: every ncu run follows the same command run plainly, exit 0)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-ffma-peak"
$B > gpurun_out/ncu_plain_c5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c5.csv $B > gpurun_out/ncu_launches_c5.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/ncu_plain_c5b.log 2>&1 && \
PTB200_CACHE_DIR=off PTB200_JIT_KEEP_SRC=gpurun_out/pt_kernel_jit_c5.cu ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 33 -c 2 -f -o gpurun_out/r02_ncu_c5 $B > gpurun_out/ncu_full_c5.log 2>&1
echo "full c5 rc=$?"
B4="python bench.py --workload c4 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --no-ffma-peak"
$B4 > gpurun_out/ncu_plain_c4.log 2>&1 && \
PTB200_CACHE_DIR=off PTB200_JIT_KEEP_SRC=gpurun_out/pt_kernel_jit_c4.cu ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 90 -c 1 -f -o gpurun_out/r02_ncu_c4 $B4 > gpurun_out/ncu_full_c4.log 2>&1
echo "full c4 rc=$?"
ls -la gpurun_out/*.ncu-rep
