"""Time C4/8 with the immediate sphere scan grouped by 1, 2, 4 or 8 chunks (PTB200_JIT_OPTS=-DPT_SPH_GROUP=g)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    from _pkg import ptb
    res = {}
    for name, (scene, w, h, spp, mode) in {"c4/8": ("synthetic", 1920, 1080, 32, 1), "Bcone": ("B", 512, 512, 64, 3)}.items():
        with ptb.Context(ptb.builtin_scene(scene, w, h)) as c:
            c.set_specialisation(2)
            best = 1e9
            for _ in range(5):
                c.render(ptb.params(w, h, spp, mode=mode))
                st = c.stats(); best = min(best, st.render_ms)
            res[name] = "%.2f ms %.0f Mp/s" % (best, st.paths / best * 1e-3)
    print(json.dumps(res))
else:
    for g in (1, 2, 4, 8):
        env = dict(os.environ, PTB200_JIT_OPTS="-DPT_SPH_GROUP=%d" % g, PTB200_CACHE_DIR="off")
        out = subprocess.check_output([sys.executable, __file__, "child"], env=env, text=True).strip().splitlines()[-1]
        print("group", g, out, flush=True)
