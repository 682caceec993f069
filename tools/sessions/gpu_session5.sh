# round 2, session 5: pixel-block path order (L2-resident accumulators): tests, bench, ncu of C5 after the change; lockstep experiment on C4
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s5_pytest.log
python tools/ab_jit_opts.py c4 - "-DPT_LOCKSTEP" > gpurun_out/s5_ab.log 2>&1
python tools/ab_jit_opts.py c2 - "-DPT_LOCKSTEP" >> gpurun_out/s5_ab.log 2>&1
PTB200_BLOCK_PIXELS=100000000 python tools/ab_jit_opts.py c5 - >> gpurun_out/s5_ab.log 2>&1
python tools/ab_jit_opts.py c5 - >> gpurun_out/s5_ab.log 2>&1
cat gpurun_out/s5_ab.log
python bench.py > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s5_bench.err
B="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-ffma-peak"
$B > gpurun_out/ncu_plain_c5c.log 2>&1 && \
PTB200_CACHE_DIR=off PTB200_JIT_KEEP_SRC=gpurun_out/pt_kernel_jit_c5.cu ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 33 -c 1 -f -o gpurun_out/r02_ncu_c5_blocked $B > gpurun_out/ncu_full_c5c.log 2>&1
echo "ncu rc=$?"
$B > gpurun_out/ncu_plain_c5d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c5.csv $B > gpurun_out/ncu_launches_c5.log 2>&1
echo "launch list rc=$?"
