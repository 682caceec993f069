import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
sc = ptb.builtin_scene("A", 512, 512)
p = ptb.params(512, 512, 512, mode=0)
for rep in range(4):
    t0 = time.perf_counter(); c = ptb.Context(sc); t1 = time.perf_counter()
    c.render(p); t2 = time.perf_counter()
    m, st = c.readback(); t3 = time.perf_counter()
    c.render(p); t4 = time.perf_counter()
    c.close(); t5 = time.perf_counter()
    print("upload %.2f ms | render(first) %.2f (gpu %.2f) | readback %.2f | render(second) %.2f | close %.2f" %
          ((t1-t0)*1e3, (t2-t1)*1e3, st.render_ms, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3))
