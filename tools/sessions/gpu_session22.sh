# round 2, session 22: maximum-size test; ncu of the small renders C1 (16 spp cosine) and C3 (32 spp uniform), specialised kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_production.py -m gpu -q -x -k "beyond_2_pow_24 or edge_cases or exactly_spp" > gpurun_out/s22_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s22_pytest.log
for wl in c1 c3; do
  B="python bench.py --workload $wl --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-ffma-peak"
  PTB200_JIT=2 $B > gpurun_out/ncu_plain_$wl.log 2>&1 && \
  PTB200_JIT=2 PTB200_CACHE_DIR=off ncu --set full --clock-control none --import-source on -k regex:k_bounce -s 12 -c 8 -f -o gpurun_out/r02_ncu_${wl}_final $B > gpurun_out/ncu_full_$wl.log 2>&1
  echo "full $wl rc=$?"
done
ls -la gpurun_out/r02_ncu_c1_final.ncu-rep gpurun_out/r02_ncu_c3_final.ncu-rep
