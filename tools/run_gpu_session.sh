ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_s48.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-ffma-peak > gpurun_out/ncu_l_s48.log 2>&1
tail -1 gpurun_out/ncu_l_s48.log | cut -c1-300
