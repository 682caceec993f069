import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"c2": ("A", 512, 512, 512, 0), "c1": ("A", 512, 512, 16, 1), "c3": ("A", 512, 512, 32, 2), "c1x8": ("A", 512, 512, 128, 1), "Bcos": ("B", 512, 512, 128, 1), "c5/16": ("A", 3840, 2160, 64, 0), "c5/4": ("A", 3840, 2160, 256, 0)}
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
            res[name] = "%.3fms" % best
    print(json.dumps(res))
else:
    for delta in ("-40", "-4", "-3", "-2", "-1", "0"):
        env = dict(os.environ, PTB200_FAIR_DELTA=delta)
        out = subprocess.check_output([sys.executable, __file__, "child"], env=env, text=True).strip().splitlines()[-1]
        print("fair delta %3s | %s" % (delta, out), flush=True)
