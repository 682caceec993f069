"""Sweep bounces-per-launch x waves (queue capacity) on one GPU for the library named by PTB200_LIB (default: in-tree)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
if os.environ.get("PTB200_LIB"):
    ptb.capi.LIB_PATH = os.environ["PTB200_LIB"]
bps = int(os.environ.get("PTB200_BPS", "5"))
wave = 148 * bps * 256
cases = {"c2": ("A", 512, 512, 512, 0), "c1x8": ("A", 512, 512, 128, 1), "c4/8": ("synthetic", 1920, 1080, 32, 1), "c5/16": ("A", 3840, 2160, 256, 0)}
only = sys.argv[1].split(",") if len(sys.argv) > 1 else list(cases)
for name in only:
    scene, w, h, spp, mode = cases[name]
    sc = ptb.builtin_scene(scene, w, h)
    with ptb.Context(sc) as c:
        for waves in [int(x) for x in os.environ.get('PTB200_WAVES', '1,2,4,8').split(',')]:
            row = []
            for iters in [int(x) for x in os.environ.get('PTB200_ITERS', '1,4,8,16,32').split(',')]:
                best = None
                for _ in range(3):
                    c.render(ptb.params(w, h, spp, mode=mode, queue_capacity=waves * wave, bounces_per_launch=iters))
                    st = c.stats()
                    if best is None or st.render_ms < best.render_ms: best = st
                row.append("K%-2d %7.2fms %6.0fMp/s L%-4d" % (iters, best.render_ms, best.paths / best.render_ms * 1e-3, best.iterations))
            print("%-6s waves %d | %s" % (name, waves, " | ".join(row)), flush=True)
