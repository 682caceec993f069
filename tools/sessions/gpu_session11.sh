# round 2, session 11: tail of renders with sample runs (64 bounces per tail launch); runs on a 1/8 share
mkdir -p gpurun_out
{
echo "== c5 full"; python tools/ab_jit_opts.py c5 -
echo "== c5 full, tail 128"; PTB200_ITERS_TAIL=128 PTB200_ITERS_DRAIN=128 python tools/ab_jit_opts.py c5 -
echo "== c5 1/2 share"; AB_WORLD=2 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 share default"; AB_WORLD=4 python tools/ab_jit_opts.py c5 -
echo "== c5 1/4 share run 64"; AB_WORLD=4 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share default"; AB_WORLD=8 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share run 64"; AB_WORLD=8 PTB200_RUN=64 python tools/ab_jit_opts.py c5 -
echo "== c5 1/8 share run 32"; AB_WORLD=8 PTB200_RUN=32 python tools/ab_jit_opts.py c5 -
echo "== c2 run 64"; PTB200_RUN=64 python tools/ab_jit_opts.py c2 -
echo "== c2 run 32"; PTB200_RUN=32 python tools/ab_jit_opts.py c2 -
} > gpurun_out/s11_ab.log 2>&1
cat gpurun_out/s11_ab.log
