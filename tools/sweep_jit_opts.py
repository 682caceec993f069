"""Occupancy / block-shape experiments on the scene-specialised kernel: PTB200_JIT_OPTS passes -D options to NVRTC.
The generic build's PT_BLOCK must match (the host launches with its own PT_BLOCK), so only PT_BLOCKS_PER_SM and
code-generation options are varied here."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"c2": ("A", 512, 512, 512, 0), "c1x8": ("A", 512, 512, 128, 1), "Bcos": ("B", 512, 512, 128, 1), "c4/8": ("synthetic", 1920, 1080, 32, 1), "c5/16": ("A", 3840, 2160, 64, 0)}
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    from _pkg import ptb
    res = {}
    for name, (scene, w, h, spp, mode) in CASES.items():
        with ptb.Context(ptb.builtin_scene(scene, w, h)) as c:
            c.set_specialisation(2)
            best = 1e9
            for _ in range(4):
                c.render(ptb.params(w, h, spp, mode=mode, queue_capacity=int(os.environ.get("PTB200_CAP", "0"))))
                st = c.stats(); best = min(best, st.render_ms)
            res[name] = "%.2fms %.0fMp/s" % (best, st.paths / best * 1e-3)
    print(json.dumps(res))
else:
    for opts, cap in [("", 0), ("-DPT_BLOCKS_PER_SM=3", 0), ("", 0), ("-DPT_BLOCKS_PER_SM=3", 0)]:
        env = dict(os.environ, PTB200_JIT_OPTS=opts, PTB200_CAP=str(cap))
        try:
            out = subprocess.check_output([sys.executable, __file__, "child"], env=env, text=True, stderr=subprocess.STDOUT).strip().splitlines()[-1]
        except subprocess.CalledProcessError as e:
            out = "FAILED " + e.output[-300:]
        print("%-44s | %s" % (opts or "(default: 4 blocks/SM)", out), flush=True)
