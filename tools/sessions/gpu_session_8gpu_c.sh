# round 2: eight GPUs - end-to-end step with owned-rows-only resolve and scaling; tile heights 10 / 6 / 5
mkdir -p gpurun_out
i=0
for t in 10 6 5; do
  i=$((i+1))
  BENCH_TILE_ROWS=$t python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$i bench.py --gpus 8 --steps 5 --warmup 3 --no-secondary > gpurun_out/m8c_tile$t.json 2> gpurun_out/m8c_tile$t.err; echo "tile $t rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/m8c_tile$t.json").read().strip().splitlines()[-1])
print("tile $t", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), round(d["e2e"]["value"]/d["value"],4), d["e2e"].get("host_image_equals_device_image"), d["phases"].get("own_render_by_rank_ms"))
PY
done
