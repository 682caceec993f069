"""First-light diagnostics on a GPU box: KATs, FP64 parity, FP32 intersect parity, FP32 statistics, timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _pkg import ptb
from oracle import pyoracle as orc

def main():
    w = h = 128
    scA = ptb.builtin_scene("A", w, h)
    ctx = ptb.Context(scA)
    print("version", ptb.lib().pt_version())
    # KATs
    e = ctx.erand48([[0, 0, 125]], 4)
    print("erand48", e.tolist())
    ph = ctx.philox([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]],
                    [[0, 0], [0xffffffff] * 2, [0xa4093822, 0x299f31d0]])
    print("philox", [[hex(x) for x in r] for r in ph.tolist()])
    tf, mhz = ctx.ffma_peak()
    print("ffma peak TFLOP/s", tf, "clock MHz", mhz)
    for scn in "ABC":
        sc = ptb.builtin_scene(scn, w, h)
        c = ptb.Context(sc)
        # intersect parity
        rng = np.random.default_rng(1)
        n = 200000
        o = np.stack([rng.uniform(2, 98, n), rng.uniform(1, 80, n), rng.uniform(1, 169, n)], 1).astype(np.float32).astype(np.float64)
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        d = d.astype(np.float32).astype(np.float64)
        rays = np.concatenate([o, d], 1)
        t_o, id_o = orc.oracle_intersect(sc, rays)
        t64, id64 = c.intersect(rays, 64)
        t32, id32 = c.intersect(rays, 32)
        print(scn, "isect fp64: id equal", np.array_equal(id_o, id64), "t bit-equal", np.array_equal(t_o, t64))
        hit = id_o >= 0
        same = id32 == id_o
        rel = np.abs(t32 - t_o)[same & hit] / t_o[same & hit]
        print(scn, "isect fp32: id mismatch %d / %d, max rel t err %.3g, >1e-6: %d" % ((~same).sum(), n, rel.max(), (rel > 1e-6).sum()))
        for mode in (0, 1, 2):
            for sincos in (0, 1):
                p = ptb.params(w, h, 8, mode=mode, engine=1, sincos=sincos)
                t0 = time.time(); c.render(p); mean, st = c.readback(); dt = time.time() - t0
                cl, omean, osq, ost = orc.oracle_render(sc, p)
                rel = np.abs(mean - omean) / np.maximum(np.abs(omean), 1e-300)
                ok = (np.abs(mean - omean) <= 1e-9 * np.abs(omean)).all(axis=2)
                print(scn, "fp64 mode", mode, "sincos", sincos, "match %.4f%% pixels, rows intact %d/%d, gpu ms %.1f" %
                      (100 * ok.mean(), ok.all(axis=1).sum(), h, st.render_ms),
                      "stats eq", (st.paths, st.rays_camera, st.rays_scatter, st.rays_shadow, st.shaded_vertices, st.miss_events, st.max_depth_seen) ==
                      (ost.paths, ost.rays_camera, ost.rays_scatter, ost.rays_shadow, ost.shaded_vertices, ost.miss_events, ost.max_depth_seen))
        for mode in (0, 1, 2):
            p = ptb.params(w, h, 256, mode=mode, engine=0, collect_stats=1)
            c.render(p); mean, sq, st = c.readback(True)
            po = ptb.params(w, h, 256, mode=mode, engine=1)
            cl, omean, osq, ost = orc.oracle_render(sc, po)
            n_s = 256
            var_g = np.maximum(sq / n_s - mean ** 2, 0) / n_s
            var_o = np.maximum(osq / n_s - omean ** 2, 0) / n_s
            z = (mean - omean) / np.sqrt(var_g + var_o + 1e-30)
            print(scn, "fp32 mode", mode, "mean gpu %.5f oracle %.5f | |z|>3: %.3f%% | rays/path gpu %.3f oracle %.3f | miss/path gpu %.4f oracle %.4f | maxd %d/%d | ms %.2f Mpaths/s %.1f its %d" %
                  (mean.mean(), omean.mean(), 100 * (np.abs(z) > 3).mean(), st.rays / st.paths, ost.rays / ost.paths,
                   st.miss_events / st.paths, ost.miss_events / ost.paths, st.max_depth_seen, ost.max_depth_seen,
                   st.render_ms, st.paths / st.render_ms * 1e-3, st.iterations))
        c.close()
    # throughput: C2
    sc = ptb.builtin_scene("A", 512, 512)
    c = ptb.Context(sc)
    for spp in (16, 512):
        for rep in range(3):
            p = ptb.params(512, 512, spp, mode=0, engine=0)
            c.render(p); mean, st = c.readback()
            print("C2-like spp", spp, "ms %.2f Mpaths/s %.1f Mrays/s %.1f iterations %d launches %d" %
                  (st.render_ms, st.paths / st.render_ms * 1e-3, st.rays / st.render_ms * 1e-3, st.iterations, st.kernel_launches))
    p = ptb.params(512, 512, 16, mode=0, engine=1)
    c.render(p); mean, st = c.readback()
    print("FP64 validate 512x512x16 NEE ms %.1f" % st.render_ms)

if __name__ == "__main__":
    main()
