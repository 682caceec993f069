# round 2, session 10: where does a 1/8 share of C5 lose 8 % against the whole of it?
mkdir -p gpurun_out
{
echo "== c5 full"; PTB200_DUMP_LAUNCHES=gpurun_out/s10_launches_full.txt python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share"; PTB200_DUMP_LAUNCHES=gpurun_out/s10_launches_w8.txt AB_WORLD=8 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 share"; AB_WORLD=4 python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 share"; AB_WORLD=2 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share, 3 waves"; AB_WORLD=8 AB_CAP=454656 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share, 256 bounces per launch"; AB_WORLD=8 AB_ITERS=256 PTB200_ITERS_DRAIN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share, 1024 bounces per launch"; AB_WORLD=8 AB_ITERS=1024 PTB200_ITERS_DRAIN=64 python tools/ab_jit_opts.py c5 -
echo "== c2"; python tools/ab_jit_opts.py c2 -
} > gpurun_out/s10_ab.log 2>&1
cat gpurun_out/s10_ab.log
tail -25 gpurun_out/s10_launches_w8.txt
