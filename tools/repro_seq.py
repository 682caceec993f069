"""repro_seq.py [--peak] scene:spp:mode ... : render the given configurations one context after the other (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
args = sys.argv[1:]
peak = "--peak" in args
args = [a for a in args if a != "--peak"]
for k, a in enumerate(args):
    scene, spp, mode = a.split(":")
    sc = ptb.builtin_scene(scene, 512, 512)
    with ptb.Context(sc) as c:
        c.set_specialisation(2)
        if peak and k == 0:
            c.ffma_peak()
        for i in range(4):
            c.render(ptb.params(512, 512, int(spp), mode=int(mode)))
            st = c.stats()
        print(a, "ok", round(st.render_ms, 2), st.iterations, flush=True)
