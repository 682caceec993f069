# round 2: two GPUs - sharding tests (CUDA IPC, shared host image, pt_render_multi) and the N = 2 bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m2_smi.txt 2>&1
python -m pytest tests/test_gpu_sharding.py tests/test_gpu_grid.py -m gpu -q > gpurun_out/m2_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/m2_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/m2_bench_n2.json 2> gpurun_out/m2_bench_n2.err; echo "bench n2 rc=$?"; tail -c 500 gpurun_out/m2_bench_n2.err
tail -1 gpurun_out/m2_bench_n2.json | cut -c1-1500
