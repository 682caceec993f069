set -x
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_s14_n$N.json 2> gpurun_out/bench_s14_n$N.err; echo "bench n$N rc=$?"; tail -1 gpurun_out/bench_s14_n$N.json
done
