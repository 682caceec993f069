"""Uniform grid vs brute-force scan (ahead-of-time build of the FP32 engine) on sphere clouds of growing size: where the
acceleration structure of SURVEY 8 f4 starts to pay.   python tools/grid_bench.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("PTB200_CACHE_DIR", "off")
from _pkg import ptb
from test_gpu_grid import _sphere_cloud
w, h, spp = 960, 540, 16
print(f"{w}x{h}, {spp} spp, cosine mode; Mrays/s (ms)")
for n in (64, 128, 256, 512, 1024, 4096, 15000):
    sc = _sphere_cloud(n, 5, w, h, rmin=0.4, rmax=1.5 if n > 1024 else 2.0)
    row = [f"{n:6d} spheres"]
    with ptb.Context(sc) as c:
        for accel, spec, name in ((0, 2, "scan, specialised"), (0, 0, "scan, generic"), (2, 0, "grid")):
            if accel == 0 and n > 512:
                row.append(f"{name}: -")
                continue
            if spec == 2 and n > 256:
                row.append(f"{name}: -")
                continue
            c.set_specialisation(spec)
            c.set_acceleration(accel)
            best = None
            for _ in range(3):
                c.render(ptb.params(w, h, spp, mode=1, seed=1))
                st = c.stats()
                if best is None or st.render_ms < best.render_ms:
                    best = st
            row.append(f"{name}: {best.rays / best.render_ms * 1e-3:8.0f} ({best.render_ms:7.2f} ms, {best.rays / best.paths:.1f} rays/path)")
    print(" | ".join(row), flush=True)
