"""A/B timing of alternative builds of libptb200.so (expt/*.so) on C2: PTB200_LIB=<path> selects the library."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    from _pkg import ptb
    if os.environ.get("PTB200_LIB"):
        ptb.capi.LIB_PATH = os.environ["PTB200_LIB"]
    sc = ptb.builtin_scene("A", 512, 512)
    with ptb.Context(sc) as c:
        best = 1e9
        for _ in range(6):
            c.render(ptb.params(512, 512, 512, mode=0))
            st = c.stats(); best = min(best, st.render_ms)
        print(json.dumps({"ms": best, "mpaths": st.paths / best * 1e-3, "its": st.iterations}))
else:
    libs = [""] + sorted(os.path.join(ROOT, "expt", f) for f in os.listdir(os.path.join(ROOT, "expt")) if f.endswith(".so"))
    for lib in libs:
        env = dict(os.environ, PTB200_LIB=lib)
        out = subprocess.check_output([sys.executable, __file__, "child"], env=env, text=True).strip().splitlines()[-1]
        print(os.path.basename(lib) or "default", out, flush=True)
