for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2959$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_s54_n$N.json 2> gpurun_out/bench_s54_n$N.err; echo "bench n$N rc=$?"; tail -1 gpurun_out/bench_s54_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['mrays_per_s'], d['config']['assembly'][:20], d.get('strong_scaling'))"
done
