"""Small renders that touch every device code path, for compute-sanitizer (one tool per gpurun call):
  compute-sanitizer --tool racecheck python tools/sanitize_target.py
  compute-sanitizer --tool memcheck  python tools/sanitize_target.py
Covers: both engines; the ahead-of-time and the scene-specialised build of k_bounce; compaction and regeneration over several
launches (tiny queue, few bounces per launch); the REFR branch stacks; statistics (shared-memory histograms); the uniform grid;
row-tile sharding with owned_rows_only; k_resolve / k_resolve_owned; pt_debug_intersect."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PTB200_CACHE_DIR", "off")
from _pkg import ptb
w, h = 48, 32
n = 0
for scene, modes in (("A", (0, 1, 2)), ("B", (1, 3)), ("G", (1,)), ("synthetic", (1, 3))):
    sc = ptb.builtin_scene(scene, w, h)
    with ptb.Context(sc) as c:
        for spec in (0, 2):
            c.set_specialisation(spec)
            for mode in modes:
                for kw in ({}, {"queue_capacity": 1024, "bounces_per_launch": 3}, {"collect_stats": 1}, {"tile_rows": 5, "rank": 1, "world": 3}):
                    c.render(ptb.params(w, h, 4, mode=mode, seed=n, **kw))
                    img, st = c.readback()
                    assert np.isfinite(img).all()
                    n += 1
        if scene in ("G", "synthetic"):
            c.set_specialisation(0)
            c.set_acceleration(2)
            c.render(ptb.params(w, h, 4, mode=1, seed=1))
            assert c.stats().accel_structure == 1
            c.set_acceleration(1)
            n += 1
        buf = c.device_alloc(w * h * 3 * 8)
        c.render_into(ptb.params(w, h, 2, mode=1, tile_rows=4, rank=0, world=2, owned_rows_only=1), buf)
        c.device_free(buf)
        c.render(ptb.params(w, h, 2, mode=1, engine=ptb.PT_ENGINE_FP64_ERAND48))
        rays = np.random.default_rng(1).normal(size=(4096, 6))
        rays[:, :3] = rays[:, :3] * 10 + (50, 40, 80)
        rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1, keepdims=True)
        c.intersect(rays, 32)
        c.intersect(rays, 64)
        n += 4
print(f"sanitize_target: {n} device passes ok")
