"""Diagnostic: generic vs specialised build on the synthetic scene, cone mode (tests/test_gpu_production.py::test_scene_specialised_kernel_matches_the_generic_one[synthetic-3])."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    from _pkg import ptb
    w, h, spp = 160, 120, 4
    sc = ptb.builtin_scene("synthetic", w, h)
    out = []
    with ptb.Context(sc) as c:
        for spec in (0, 2):
            c.set_specialisation(spec)
            c.render(ptb.params(w, h, spp, mode=3, seed=11))
            mean, st = c.readback()
            out.append((mean.copy(), st.rays, st.shaded_vertices, st.spawned_branches, st.split_refusals, st.miss_events, st.rays_shadow, st.rays_scatter, st.iterations))
    print("generic    :", out[0][1:])
    print("specialised:", out[1][1:])
    print("pixels identical: %.4f %%" % (100 * (out[0][0] == out[1][0]).all(axis=2).mean()))
else:
    for opts in ("", "-DPT_NO_LOCKSTEP", "-DPT_NO_SPLIT", "-DPT_NO_LOCKSTEP -DPT_NO_SPLIT"):
        env = dict(os.environ, PTB200_JIT_OPTS=opts, PTB200_CACHE_DIR="off")
        print("== JIT opts:", opts or "(none)", flush=True)
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout, flush=True)
