import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
spec = int(sys.argv[1]) if len(sys.argv) > 1 else 2
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sc = ptb.builtin_scene("B", 512, 512)
with ptb.Context(sc) as c:
    c.set_specialisation(spec)
    for i in range(3):
        c.render(ptb.params(512, 512, spp, mode=1))
        st = c.stats()
        print("render", i, "ok", st.render_ms, st.iterations, st.max_depth_seen, flush=True)
