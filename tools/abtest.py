"""A/B timing of alternative builds of libptb200.so (expt/*.so) on C2: PTB200_LIB=<path> selects the library."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    from _pkg import ptb
    if os.environ.get("PTB200_LIB"):
        ptb.capi.LIB_PATH = os.environ["PTB200_LIB"]
    res = {}
    for name, (scene, w, h, spp, mode) in {"c2": ("A", 512, 512, 512, 0), "c1x8": ("A", 512, 512, 128, 1), "Bcos": ("B", 512, 512, 128, 1),
                                            "c4/8": ("synthetic", 1920, 1080, 32, 1), "c5/16": ("A", 3840, 2160, 64, 0)}.items():
        sc = ptb.builtin_scene(scene, w, h)
        with ptb.Context(sc) as c:
            best = 1e9
            for _ in range(4):
                c.render(ptb.params(w, h, spp, mode=mode))
                st = c.stats(); best = min(best, st.render_ms)
            res[name] = "%.2f ms %.0f Mp/s" % (best, st.paths / best * 1e-3)
    print(json.dumps(res))
else:
    libs = [""] + sorted(os.path.join(ROOT, "expt", f) for f in os.listdir(os.path.join(ROOT, "expt")) if f.endswith(".so"))
    for lib in libs:
        env = dict(os.environ, PTB200_LIB=lib)
        out = subprocess.check_output([sys.executable, __file__, "child"], env=env, text=True).strip().splitlines()[-1]
        print(os.path.basename(lib) or "default", out, flush=True)
