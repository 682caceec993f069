# round 2: eight GPUs - the N = 8, 4, 2 bench lines of the final kernel (C5 at every N)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m8_smi.txt 2>&1
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/m8_bench_n$n.json 2> gpurun_out/m8_bench_n$n.err; echo "bench n$n rc=$?"; tail -c 300 gpurun_out/m8_bench_n$n.err
  tail -1 gpurun_out/m8_bench_n$n.json | cut -c1-600
done
