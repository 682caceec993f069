"""Host-only: build the scene-specialised kernel of a built-in scene offline and report its resources.
  python tools/jit_offline.py SCENE MODE [OUT_PREFIX]   -> OUT_PREFIX.cu (the NVRTC translation unit), OUT_PREFIX.cubin, registers / spills / SASS size
No GPU needed (NVRTC cross-compiles sm_100a)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
scene, mode = sys.argv[1], int(sys.argv[2])
prefix = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "expt", f"jit_{scene}_{mode}")
os.makedirs(os.path.dirname(prefix), exist_ok=True)
os.environ["PTB200_JIT_KEEP_SRC"] = prefix + ".cu"
os.environ["PTB200_JIT_DUMP"] = prefix + ".cubin"
os.environ["PTB200_CACHE_DIR"] = "off"
from _pkg import ptb
w, h = (1920, 1080) if scene == "synthetic" else (512, 512)
spec, nbytes, secs = ptb.specialise(ptb.builtin_scene(scene, w, h), mode)
print(f"{scene} mode {mode}: cubin {nbytes} B, NVRTC {secs:.2f} s")
res = subprocess.run(["cuobjdump", "-res-usage", prefix + ".cubin"], capture_output=True, text=True).stdout
print("\n".join(l for l in res.splitlines() if "REG" in l or "Function" in l))
sass = subprocess.run(["cuobjdump", "-sass", "-fun", "k_bounce_jit", prefix + ".cubin"], capture_output=True, text=True).stdout
ins = [l for l in sass.splitlines() if "/*" in l and ";" in l]
print("k_bounce_jit SASS instructions:", len(ins), " FFMA2:", sum("FFMA2" in l for l in ins), " STL/LDL:", sum((" STL" in l or " LDL" in l) for l in ins))
open(prefix + ".sass", "w").write(sass)
