"""How long pt_readback_owned takes for a one-eighth share of a 4K image, by row-tile height (the strided DMA has one row per tile)."""
import os, sys, time, mmap
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
w, h = 3840, 2160
sc = ptb.builtin_scene("A", w, h)
# a shared-memory mapping like dist.HostImage's, page-locked through the library
buf = mmap.mmap(-1, w * h * 3 * 8)
host = np.frombuffer(buf, dtype=np.float64).reshape(h, w, 3)
with ptb.Context(sc) as c:
    c.host_register(host)
    for tile in (10, 5, 2, 1):
        for oro in (1, 0):
            c.render(ptb.params(w, h, 4, mode=0, tile_rows=tile, rank=3, world=8, owned_rows_only=oro))
            c.readback_owned(host)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter(); c.readback_owned(host); ts.append((time.perf_counter() - t0) * 1e3)
            rows = np.arange(h)[(np.arange(h) // tile) % 8 == 3]
            ok = bool(np.isfinite(host[rows]).all() and host[rows].sum() > 0)
            print(f"tile {tile:2d} owned_rows_only={oro}: readback_owned median {np.median(ts):.3f} ms  min {min(ts):.3f} ms  (25 MB; rows ok: {ok})", flush=True)
    c.host_unregister(host)
