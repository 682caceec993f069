# round 2, session 21: launch timeline of C5 with sample runs (where do the 5.4 ms at the end go?)
mkdir -p gpurun_out
rm -f gpurun_out/s21_launches.txt
PTB200_DUMP_LAUNCHES=gpurun_out/s21_launches.txt python tools/ab_jit_opts.py c5 - > gpurun_out/s21_ab.log 2>&1
cat gpurun_out/s21_ab.log
tail -30 gpurun_out/s21_launches.txt
