"""All five BASELINE.json configurations on ONE GPU (C5 at reduced spp unless --full), both scenes where the config says so:
Mpaths/s, Mrays/s, rays/path, fraction of the FP32 roofline (SURVEY 8d accounting).  Prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import ptb
FULL = "--full" in sys.argv
F_SHADE = {0: 150.0, 1: 110.0, 2: 110.0, 3: 150.0}
CONFIGS = [("C1 scene A 512x512 16 spp cosine", "A", 512, 512, 16, 1), ("C1 scene B 512x512 16 spp cosine", "B", 512, 512, 16, 1),
           ("C2 scene A 512x512 512 spp NEE (reference rect light)", "A", 512, 512, 512, 0), ("C2 scene B 512x512 512 spp NEE (cone, sphere light)", "B", 512, 512, 512, 3),
           ("C3 scene A 512x512 32 spp uniform", "A", 512, 512, 32, 2), ("C3 scene B 512x512 32 spp uniform", "B", 512, 512, 32, 2),
           ("C4 synthetic 256 spheres 1920x1080 256 spp cosine", "synthetic", 1920, 1080, 256, 1),
           ("C5 scene A 3840x2160 %d spp NEE (1 GPU)" % (1024 if FULL else 128), "A", 3840, 2160, 1024 if FULL else 128, 0)]
out = []
peak = None
for desc, scene, w, h, spp, mode in CONFIGS:
    sc = ptb.builtin_scene(scene, w, h)
    with ptb.Context(sc) as c:
        c.set_specialisation(2)
        if peak is None:
            peak = c.ffma_peak()[0]
        best = None
        for _ in range(4):
            c.render(ptb.params(w, h, spp, mode=mode))
            st = c.stats()
            if best is None or st.render_ms < best.render_ms:
                best = st
        flops = float(best.rays) * sc.flops_per_ray() + float(best.shaded_vertices) * F_SHADE[mode]
        tf = flops / (best.render_ms * 1e-3) / 1e12
        out.append({"config": desc, "ms": round(best.render_ms, 3), "mpaths_per_s": round(best.paths / best.render_ms * 1e-3, 1),
                    "mrays_per_s": round(best.rays / best.render_ms * 1e-3, 1), "rays_per_path": round(best.rays / best.paths, 3),
                    "launches": int(best.iterations), "max_depth": int(best.max_depth_seen), "tflops_algorithmic": round(tf, 2),
                    "frac_fp32_peak": round(tf / peak, 4), "specialised": bool(best.specialised)})
        print("%-58s %9.2f ms %9.1f Mpaths/s %9.1f Mrays/s  %5.1f %% of FP32 peak" % (desc, best.render_ms, out[-1]["mpaths_per_s"], out[-1]["mrays_per_s"], 100 * tf / peak), file=sys.stderr, flush=True)
print(json.dumps({"ffma_peak_tflops": peak, "configs": out}))
