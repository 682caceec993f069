# round 2, session 31: ahead-of-time (generic) kernel with bigger blocks; ncu of C5 / C2 with the final kernel (1024-thread blocks)
mkdir -p gpurun_out
ABTEST_EXPT_SPEC=0 python tools/abtest.py > gpurun_out/s31_abtest_generic_blocks.log 2>&1; cat gpurun_out/s31_abtest_generic_blocks.log | cut -c1-260
B="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-ffma-peak"
$B > gpurun_out/ncu_plain_c5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_c5.csv $B > gpurun_out/ncu_launches_c5.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/ncu_plain_c5b.log 2>&1 && \
PTB200_CACHE_DIR=off PTB200_JIT_KEEP_SRC=gpurun_out/pt_kernel_jit_c5.cu ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 110 -c 1 -f -o gpurun_out/r02_ncu_c5_final $B > gpurun_out/ncu_full_c5.log 2>&1
echo "full c5 rc=$?"
B2="python bench.py --workload c2 --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-ffma-peak"
$B2 > gpurun_out/ncu_plain_c2.log 2>&1 && \
PTB200_CACHE_DIR=off ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 18 -c 1 -f -o gpurun_out/r02_ncu_c2_final $B2 > gpurun_out/ncu_full_c2.log 2>&1
echo "full c2 rc=$?"
