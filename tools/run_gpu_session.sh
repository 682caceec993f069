python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/abtest.py 2>&1 | tee gpurun_out/abtest_v7.txt
python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_v7.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-ffma-peak > gpurun_out/launches_v7.log 2>&1; echo "launch list rc=$?"
PTB200_JIT_KEEP_SRC=/root/repo/gpurun_out/pt_kernel_jit_c2.cu timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 18 -c 6 -o gpurun_out/ncu_c2_v7 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_c2_v7.log 2>&1; echo "ncu rc=$?"
