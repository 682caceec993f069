"""Latency breakdown of one bounce for a lone warp (tuning aid): a 4x1-pixel image (one warp), specialised kernel built with
-DPT_SECTION_CLOCKS; prints SM cycles per warp-iteration by section.   python tools/section_clocks.py [scene] [mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PTB200_JIT_OPTS"] = "-DPT_SECTION_CLOCKS"
os.environ["PTB200_CACHE_DIR"] = "off"
from _pkg import ptb
scene = sys.argv[1] if len(sys.argv) > 1 else "A"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
names = ["pop + regeneration", "Philox + camera ray", "closest hit", "material fetch + refine + normal", "roulette + sampling (+ shadow ray)", "accumulation + reconvergence", "-", "push + loop back"]
for w, h, spp in ((4, 1, 512), (512, 512, 32)):
    with ptb.Context(ptb.builtin_scene(scene, w, h)) as c:
        c.set_specialisation(2)
        for _ in range(2):
            c.render(ptb.params(w, h, spp, mode=mode, seed=0))
            st = c.stats()
        hist = list(st.live_at_depth)
        iters = hist[39] or 1
        print(f"scene {scene} mode {mode} {w}x{h}x{spp}: {st.render_ms:.3f} ms, {iters} warp-iterations, {st.shaded_vertices} vertices, total {sum(hist[40:48]) / iters:.0f} cycles per warp-iteration")
        for k in range(8):
            if hist[40 + k]:
                print(f"   {names[k]:38s} {hist[40 + k] / iters:8.1f} cycles")
