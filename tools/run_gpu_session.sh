python -m pytest tests/test_gpu_sharding.py -m gpu -q -x > gpurun_out/pytest_gpu_s33.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_s33.log
./small-pathtracer_b200/smallpt 256 --size 1920x1080 --out /tmp/one.ppm
./small-pathtracer_b200/smallpt 256 --size 1920x1080 --gpus 2 --out /tmp/two.ppm
cmp /tmp/one.ppm /tmp/two.ppm && echo "IMAGES IDENTICAL"
./small-pathtracer_b200/smallpt 1024 --size 3840x2160 --gpus 2 --out /tmp/c5.ppm
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_s33_n2.json 2> gpurun_out/bench_s33_n2.err; tail -1 gpurun_out/bench_s33_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['assembly'], d['config']['tile_rows'], d.get('strong_scaling'))"
