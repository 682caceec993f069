"""CPU: bench.py's reference arm prints ONE JSON line with the keys the round driver reads, and our own arm refuses
to run without a CUDA device (no CPU fallback).  The GPU arm's line is checked on the B200 by test_bench_line_on_gpu."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "e2e", "cpu_baseline"}


def _run(args, **kw):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT, **kw)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c1"], timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "Mpaths/s" and d["unit"] == "Mpaths/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "workload" in d["config"]


def test_own_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = _run(["--steps", "1", "--warmup", "3"], timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_bench_line_on_gpu():
    r = _run(["--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-secondary"], timeout=900)
    assert r.returncode == 0, r.stderr
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    assert (BASE_KEYS - {"cpu_baseline"}) | {"clocks", "gpu_launches", "roofline", "mrays_per_s"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["dtype"] == "f32" and d["gpu_launches"] > 0
    assert d["config"]["workload"].startswith("C5") and d["value"] > 1000 and 0 < d["e2e"]["value"] <= d["value"] * 1.02
    assert d["scaling"] == "strong"             # the same workload at every N
    assert d["e2e"]["d2h_bytes_per_step"] == 3840 * 2160 * 3 * 8 and d["e2e"]["h2d_bytes_per_step"] > 0
    ph = d["phases"]
    assert ph["generating"]["max_over_ranks_ms"] > 0 and ph["tail"]["max_over_ranks_ms"] >= 0 and ph["resolve"]["max_over_ranks_ms"] > 0
    assert ph["generating"]["max_over_ranks_ms"] + ph["tail"]["max_over_ranks_ms"] + ph["resolve"]["max_over_ranks_ms"] <= d["ms_per_step"] * 1.01
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
