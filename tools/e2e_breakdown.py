"""Where the end-to-end step of bench.py spends its wall time (C2): scene upload, render (wall vs device), read-back."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
w = h = 512
sc = ptb.builtin_scene("A", w, h)
p = ptb.params(w, h, 512, mode=0)
out = np.empty((h, w, 3))
with ptb.Context(sc) as c:
    for _ in range(3):
        c.update_scene(sc); c.render(p); c.readback(out=out)
    rows = []
    for _ in range(10):
        t0 = time.perf_counter(); c.update_scene(sc)
        t1 = time.perf_counter(); c.render(p)
        t2 = time.perf_counter(); _, st = (c.readback_view() if os.environ.get("VIEW") else c.readback(out=out))
        t3 = time.perf_counter()
        rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, st.render_ms, (t3 - t2) * 1e3))
    a = np.array(rows)
    print("median ms: upload %.3f | render wall %.3f (device %.3f, host overhead %.3f) | readback %.3f | total %.3f" %
          (np.median(a[:, 0]), np.median(a[:, 1]), np.median(a[:, 2]), np.median(a[:, 1] - a[:, 2]), np.median(a[:, 3]), np.median(a.sum(1) - a[:, 2])))
    t0 = time.perf_counter()
    for _ in range(20):
        c.stats()
    print("stats() call: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
