# round 2, session 8: sample runs: tests, A/B over the run length
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s8_pytest.log
for r in 1 16 64 128 256; do echo "run $r"; PTB200_RUN=$r python tools/ab_jit_opts.py c5 - ; done > gpurun_out/s8_ab.log 2>&1
echo "default" >> gpurun_out/s8_ab.log; python tools/ab_jit_opts.py c5 - >> gpurun_out/s8_ab.log 2>&1
for r in 1 2 4 8; do echo "run $r"; PTB200_RUN=$r python tools/ab_jit_opts.py c2 - ; done >> gpurun_out/s8_ab.log 2>&1
echo "default" >> gpurun_out/s8_ab.log; python tools/ab_jit_opts.py c2 - >> gpurun_out/s8_ab.log 2>&1
python tools/ab_jit_opts.py c4 - >> gpurun_out/s8_ab.log 2>&1
cat gpurun_out/s8_ab.log
