"""Utilisation over time inside a render (tuning aid): builds the specialised kernel with -DPT_TIMELINE=<bucket ns> and prints,
per time bucket since the first launch started, the active lanes per warp-iteration and the share of all lane-iterations.
  python tools/timeline.py WORKLOAD [bucket_us]     WORKLOAD: c1 c1b c2 c3 c3b (bench.py's names)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bucket_us = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
os.environ["PTB200_JIT_OPTS"] = f"-DPT_TIMELINE={int(bucket_us * 1000)}"
os.environ["PTB200_CACHE_DIR"] = "off"
from _pkg import ptb
import bench
scene, w, h, spp, mode, desc = bench.WORKLOADS[sys.argv[1]]
with ptb.Context(ptb.builtin_scene(scene, w, h)) as c:
    c.set_specialisation(2)
    for _ in range(3):
        c.render(ptb.params(w, h, spp, mode=mode, seed=0))
        st = c.stats()
    hist = list(st.live_at_depth)
    print(f"{desc}: {st.render_ms:.3f} ms, generating {st.main_kernel_ms:.3f} tail {st.tail_ms:.3f}, launches {st.iterations}, max depth {st.max_depth_seen}")
    tot = sum(hist[:32]) or 1
    for k in range(32):
        if hist[32 + k]:
            print(f"  {k * bucket_us:7.0f} us  lanes/warp-iter {hist[k] / hist[32 + k]:5.1f}  warp-iters {hist[32 + k]:9d}  share of lane-iterations {100.0 * hist[k] / tot:5.1f} %")
