"""A/B timing on one GPU: the in-tree library with the generic and the scene-specialised kernel, plus any alternative
builds under expt/*.so (PTB200_LIB=<path> selects the library in the child process)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"c2": ("A", 512, 512, 512, 0), "c1x8": ("A", 512, 512, 128, 1), "Bcos": ("B", 512, 512, 128, 1), "Bcone": ("B", 512, 512, 64, 3),
         "c4/8": ("synthetic", 1920, 1080, 32, 1), "c5/16": ("A", 3840, 2160, 64, 0)}
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    from _pkg import ptb
    if os.environ.get("PTB200_LIB"):
        ptb.capi.LIB_PATH = os.environ["PTB200_LIB"]
    spec = int(os.environ.get("PTB200_SPEC", "1"))
    res = {}
    for name, (scene, w, h, spp, mode) in CASES.items():
        sc = ptb.builtin_scene(scene, w, h)
        with ptb.Context(sc) as c:
            c.set_specialisation(spec)
            best = 1e9
            for _ in range(4):
                c.render(ptb.params(w, h, spp, mode=mode))
                st = c.stats(); best = min(best, st.render_ms)
            res[name] = "%.2f ms %.0f Mp/s%s" % (best, st.paths / best * 1e-3, " [spec]" if st.specialised else "")
    print(json.dumps(res))
else:
    libs = [("", "0"), ("", "2")] + [(os.path.join(ROOT, "expt", f), os.environ.get("ABTEST_EXPT_SPEC", "2")) for f in (sorted(os.listdir(os.path.join(ROOT, "expt"))) if os.path.isdir(os.path.join(ROOT, "expt")) else []) if f.endswith(".so")]
    for lib, spec in libs:
        env = dict(os.environ, PTB200_LIB=lib, PTB200_SPEC=spec)
        out = subprocess.check_output([sys.executable, __file__, "child"], env=env, text=True).strip().splitlines()[-1]
        print((os.path.basename(lib) or "default") + " spec=" + spec, out, flush=True)
