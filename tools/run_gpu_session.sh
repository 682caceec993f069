python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29594 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_v9_n4.json 2> gpurun_out/bench_v9_n4.err; echo "bench n4 rc=$?"; tail -1 gpurun_out/bench_v9_n4.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['mrays_per_s'], d['config']['assembly'][:20], d.get('strong_scaling'))"
