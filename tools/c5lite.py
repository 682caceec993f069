"""C5-lite (scene A, 3840x2160, NEE) at reduced spp: one render for ncu / timing.  usage: c5lite.py [spp] [capacity]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = ptb.builtin_scene("A", 3840, 2160)
with ptb.Context(sc) as c:
    for _ in range(2):
        c.render(ptb.params(3840, 2160, spp, mode=0, tile_rows=16, queue_capacity=cap))
        st = c.stats()
        print("C5-lite %d spp cap %d: %.2f ms %.1f Mpaths/s its %d" % (spp, cap, st.render_ms, st.paths / st.render_ms * 1e-3, st.iterations), flush=True)
