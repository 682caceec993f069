python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_v8_n2.json 2> gpurun_out/bench_v8_n2.err; echo "bench n2 rc=$?"; tail -1 gpurun_out/bench_v8_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['mrays_per_s'], d['config']['assembly'][:20], d.get('strong_scaling'))"
python -m pytest tests/test_gpu_sharding.py -m gpu -x -q 2>&1 | tail -2
