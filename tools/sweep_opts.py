"""Time a few renders under alternative NVRTC options: sweep_opts.py "-DX=1" "-DX=2 -DY=3" ...  ("" = defaults)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"Bcone": ("B", 512, 512, 128, 3), "Bcos": ("B", 512, 512, 128, 1), "Buni": ("B", 512, 512, 128, 2), "c2/4": ("A", 512, 512, 128, 0),
         "c4/16": ("synthetic", 1920, 1080, 16, 1), "c4cone": ("synthetic", 960, 540, 16, 3), "c4nee": ("synthetic", 960, 540, 16, 0)}
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    from _pkg import ptb
    res = {}
    for name, (scene, w, h, spp, mode) in CASES.items():
        with ptb.Context(ptb.builtin_scene(scene, w, h)) as c:
            c.set_specialisation(2)
            best = 1e9
            for _ in range(5):
                c.render(ptb.params(w, h, spp, mode=mode))
                st = c.stats(); best = min(best, st.render_ms)
            res[name] = "%.2f ms %.0f Mp/s" % (best, st.paths / best * 1e-3)
    print(json.dumps(res))
else:
    for opts in sys.argv[1:] or [""]:
        env = dict(os.environ, PTB200_CACHE_DIR="off")
        if opts:
            env["PTB200_JIT_OPTS"] = opts
        out = subprocess.check_output([sys.executable, __file__, "child"], env=env, text=True).strip().splitlines()[-1]
        print("[%s]" % opts, out, flush=True)
