# round 2: eight GPUs - 10-row vs 2-row tiles, alternating (is the end-to-end difference real or noise?)
mkdir -p gpurun_out
i=0
for t in 10 2 10 2 3; do
  i=$((i+1))
  BENCH_TILE_ROWS=$t python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 8 --steps 8 --warmup 3 --no-secondary > gpurun_out/m8d_run${i}_tile$t.json 2> gpurun_out/m8d_run${i}_tile$t.err; echo "run $i tile $t rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/m8d_run${i}_tile$t.json").read().strip().splitlines()[-1])
print("tile $t", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["value"]/d["value"],4), d["e2e"].get("host_image_equals_device_image"), d["phases"].get("own_render_by_rank_ms"))
PY
done
