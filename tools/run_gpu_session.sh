set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s2.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s2_c2.json 2> gpurun_out/bench_s2_c2.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/bench_s2_c4.json 2> gpurun_out/bench_s2_c4.err; echo "bench c4 rc=$?"
python bench.py --steps 2 --warmup 3 --workload c5 --no-cpu-baseline > gpurun_out/bench_s2_c5.json 2> gpurun_out/bench_s2_c5.err; echo "bench c5 rc=$?"
python tools/sweep.py > gpurun_out/sweep_s2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_s2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_s2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 150 -c 2 -o gpurun_out/prof_s2_bounce -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_s2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 40 -c 2 -o gpurun_out/prof_s2_bounce_c4 -f python bench.py --steps 1 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/ncu_f_s2_c4.log 2>&1
tail -3 gpurun_out/pytest_gpu_s2.log; cat gpurun_out/bench_s2_c2.json; cat gpurun_out/sweep_s2.log
