"""Tail tuning: bounces per launch once generation is exhausted (PTB200_ITERS_TAIL), in the drain phase (PTB200_ITERS_DRAIN),
the drain threshold (PTB200_DRAIN_BELOW) and the host batch (PTB200_BATCH).  Each setting runs in a child process."""
import os, sys, subprocess, json, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"c2": ("A", 512, 512, 512, 0), "c1": ("A", 512, 512, 16, 1), "c3": ("A", 512, 512, 32, 2), "c1x8": ("A", 512, 512, 128, 1), "Bcos": ("B", 512, 512, 128, 1)}
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
            res[name] = "%.2fms L%d" % (best, st.iterations)
    print(json.dumps(res))
else:
    for tail, drain, below, batch in [(2, 16, 37888, 2), (2, 16, 37888, 1), (2, 16, 37888, 4), (4, 16, 37888, 2), (8, 16, 37888, 2), (1, 16, 37888, 2),
                                      (2, 32, 37888, 2), (2, 64, 37888, 2), (2, 16, 151552, 2), (4, 32, 151552, 2), (2, 16, 9472, 2), (4, 64, 303104, 2)]:
        env = dict(os.environ, PTB200_ITERS_TAIL=str(tail), PTB200_ITERS_DRAIN=str(drain), PTB200_DRAIN_BELOW=str(below), PTB200_BATCH=str(batch))
        out = subprocess.check_output([sys.executable, __file__, "child"], env=env, text=True).strip().splitlines()[-1]
        print("tail %2d drain %2d below %6d batch %d | %s" % (tail, drain, below, batch, out), flush=True)
