python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s53.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_s53.log
python bench.py --steps 10 > gpurun_out/bench_s53_c2.json 2> gpurun_out/bench_s53_c2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s53_c2.json').read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "mrays", d["mrays_per_s"], "frac", d["roofline"]["frac"], "c4", d["secondary"]["c4"]["value"], d["secondary"]["c4"]["roofline"]["frac"], "launches/step", d["roofline"]["launches_per_step"], "traffic", d["roofline"]["traffic"])
PY
python tools/run_configs.py --full > gpurun_out/configs_s53.json 2> gpurun_out/configs_s53.log; cat gpurun_out/configs_s53.log
