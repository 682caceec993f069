python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_s34.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_s34.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_s34_c2.json 2> gpurun_out/bench_s34_c2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s34_c2.json').read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "c4", d["secondary"]["c4"]["value"], d["secondary"]["c4"]["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "launches", d["gpu_launches"], d["clocks"])
PY
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
