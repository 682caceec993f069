"""One render of a built-in scene for ncu: prof_scene.py SCENE SPP MODE (specialised build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
scene, spp, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
with ptb.Context(ptb.builtin_scene(scene, 512, 512)) as c:
    c.set_specialisation(2)
    for _ in range(2):
        c.render(ptb.params(512, 512, spp, mode=mode))
    print(c.stats().render_ms)
