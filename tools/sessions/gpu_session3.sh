# round 2, session 3: tests, latency breakdown of one bounce, A/B of the huge-sphere forms, generic vs specialised
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s3_pytest.log
python tools/section_clocks.py A 2 > gpurun_out/s3_section_clocks.log 2>&1
python tools/section_clocks.py A 0 >> gpurun_out/s3_section_clocks.log 2>&1
python tools/section_clocks.py B 1 >> gpurun_out/s3_section_clocks.log 2>&1
cat gpurun_out/s3_section_clocks.log
for wl in c2b c1b; do python tools/ab_jit_opts.py $wl - "-DPT_HUGE_FP64" >> gpurun_out/s3_ab.log 2>&1; done
python tools/ab_jit_opts.py c2 - generic "-DPT_NO_SPLIT" >> gpurun_out/s3_ab.log 2>&1
python tools/ab_jit_opts.py c1 - generic >> gpurun_out/s3_ab.log 2>&1
python tools/ab_jit_opts.py c4 - "-DPT_NO_SPLIT" >> gpurun_out/s3_ab.log 2>&1
cat gpurun_out/s3_ab.log
