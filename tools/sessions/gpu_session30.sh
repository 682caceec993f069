# round 2, session 30: slots in flight and bounces per launch with 1024-thread blocks
mkdir -p gpurun_out
{
for w in 3 4 8 12; do echo "== c5 $w waves"; AB_CAP=$((148*1024*w)) python tools/ab_jit_opts.py c5 - ; done
echo "== c5 1024 bounces"; AB_ITERS=1024 PTB200_ITERS_DRAIN=64 PTB200_ITERS_TAIL=64 python tools/ab_jit_opts.py c5 -
echo "== c5 256 bounces"; AB_ITERS=256 PTB200_ITERS_DRAIN=64 PTB200_ITERS_TAIL=64 python tools/ab_jit_opts.py c5 -
for w in 3 12; do echo "== c2 $w waves"; AB_CAP=$((148*1024*w)) python tools/ab_jit_opts.py c2 - ; done
} > gpurun_out/s30_ab.log 2>&1
cat gpurun_out/s30_ab.log
