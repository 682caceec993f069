"""Queue-capacity / workload sweep on one GPU: prints ms and Mpaths/s from the library's own CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
wave = 148 * 4 * 256
def run(scene, w, h, spp, mode, cap, reps=3):
    sc = ptb.builtin_scene(scene, w, h)
    with ptb.Context(sc) as c:
        best = None
        for _ in range(reps):
            c.render(ptb.params(w, h, spp, mode=mode, queue_capacity=cap))
            st = c.stats()
            if best is None or st.render_ms < best.render_ms: best = st
        return best
for waves in (3, 4, 5, 6, 7, 8, 10, 12, 16):
    st = run("A", 512, 512, 512, 0, waves * wave)
    print("C2 waves %2d cap %8d: %7.2f ms  %8.1f Mpaths/s  its %d" % (waves, waves * wave, st.render_ms, st.paths / st.render_ms * 1e-3, st.iterations), flush=True)
for name, (scene, w, h, spp, mode) in {"C1 A cos 16spp": ("A", 512, 512, 16, 1), "C3 A uni 32spp": ("A", 512, 512, 32, 2), "B cos 64spp": ("B", 512, 512, 64, 1),
                                        "B nee 512spp": ("B", 512, 512, 512, 0), "B cone 64spp": ("B", 512, 512, 64, 3),
                                        "C4 synthetic 1080p 16spp cos": ("synthetic", 1920, 1080, 16, 1), "C5-lite A 4K 64spp nee": ("A", 3840, 2160, 64, 0)}.items():
    st = run(scene, w, h, spp, mode, 0, reps=2)
    print("%-30s %8.2f ms  %8.1f Mpaths/s  %8.1f Mrays/s  rays/path %.2f its %d maxdepth %d" % (name, st.render_ms, st.paths / st.render_ms * 1e-3, st.rays / st.render_ms * 1e-3, st.rays / st.paths, st.iterations, st.max_depth_seen), flush=True)
