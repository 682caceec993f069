# round 2: eight GPUs - row-tile height and the balance between ranks (C5, N = 8)
mkdir -p gpurun_out
for t in 10 4 2 1; do
  BENCH_TILE_ROWS=$t python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$t bench.py --gpus 8 --steps 5 --warmup 3 --no-secondary > gpurun_out/m8b_tile$t.json 2> gpurun_out/m8b_tile$t.err; echo "tile $t rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/m8b_tile$t.json").read().strip().splitlines()[-1])
print("tile $t", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["phases"].get("own_render_by_rank_ms"), {k:(round(v["max_over_ranks_ms"],3), round(v["min_over_ranks_ms"],3)) for k,v in d["phases"].items() if isinstance(v,dict)})
PY
done
python -m pytest tests/test_gpu_sharding.py -m gpu -q > gpurun_out/m8b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/m8b_pytest.log
