"""GPU: the SPEC / REFR arms of radiance() (reference src/smallpt.cpp:481-495; live in the sphere-era binary).

The reference reflects off mirrors, and at a dielectric follows BOTH arms while depth <= 2 (:494-495), one arm chosen
with P = .25 + .5 Re afterwards (:492-493).  Checked here:
  * the FP64 erand48 engine against the CPU oracle per pixel (1e-9) on scene G (the sphere-era box with Beason's
    mirror and glass spheres, ids 7 and 8);
  * the FP32 engine against the oracle's converged image (3 sigma), with the split (production) and without it
    (collect_stats renders take one arm at every depth: same expectation);
  * Fresnel / total-internal-reflection known answers on single rays (the split makes them noise-free);
  * a white furnace (every surface emits E with albedo rho: radiance = E / (1 - rho) everywhere);
  * the split's variance on the glass sphere against the oracle's."""
import os

import numpy as np
import pytest

from conftest import ptb, orc, GOLDEN

pytestmark = pytest.mark.gpu


def _z(mean_a, var_a_of_mean, mean_b, var_b_of_mean):
    se = np.sqrt(var_a_of_mean + var_b_of_mean)
    return (mean_a - mean_b) / np.maximum(se, 1e-12), se


def test_fp64_engine_matches_oracle_on_mirror_and_glass():
    w, h, spp = 64, 48, 8
    sc = ptb.builtin_scene("G", w, h)
    p = ptb.params(w, h, spp, mode=ptb.PT_MODE_COS, engine=ptb.PT_ENGINE_FP64_ERAND48, sincos=ptb.PT_SINCOS_DET)
    with ptb.Context(sc) as c:
        c.render(p)
        mean, st = c.readback()
    cl, omean, osq, ost = orc.oracle_render(sc, p)
    ok = (np.abs(mean - omean) <= 1e-9 * np.abs(omean)).all(axis=2)
    assert ok.mean() >= 0.999, ok.mean()
    assert st.rays == ost.rays and st.shaded_vertices == ost.shaded_vertices      # a branch flip would change the counters


def test_gate2_three_sigma_on_mirror_and_glass():
    ref = np.load(os.path.join(GOLDEN, "converged_G_cos.npz"))
    omean, osq, n_o = ref["mean"].astype(np.float64), ref["sumsq"].astype(np.float64), int(ref["spp"])
    h, w, _ = omean.shape
    spp = 4096
    sc = ptb.builtin_scene("G", w, h)
    var_o = np.maximum(osq / n_o - omean ** 2, 0)
    with ptb.Context(sc) as c:
        # one arm at every depth (collect_stats): per-pixel variance of its own
        c.render(ptb.params(w, h, spp, mode=1, seed=77, collect_stats=1))
        m1, sq1, st1 = c.readback(True)
        assert st1.spawned_branches == 0
        var1 = np.maximum(sq1 / spp - m1 ** 2, 0)
        # production: both arms while depth <= 2; its per-sample variance is the oracle's (same estimator)
        c.render(ptb.params(w, h, spp, mode=1, seed=78))
        m2, st2 = c.readback()
        assert st2.spawned_branches > 0 and st2.truncated == 0 and st2.split_refusals == 0
    for name, m, var in (("one arm", m1, var1), ("split", m2, var_o)):
        z, se = _z(m, var / spp, omean, var_o / n_o)
        informative = se > 1e-9
        frac = ((np.abs(z) > 3) & informative).sum() / informative.sum()
        assert frac <= 0.012, (name, frac)
        tot_diff, tot_se = (m - omean).sum(), np.sqrt((se ** 2).sum())
        assert abs(tot_diff) < 0.01 * omean.sum() + 4 * tot_se, (name, tot_diff / omean.sum())
    # the split traces more rays per camera path than it has vertices on a single arm: the oracle's count
    assert abs(st2.rays / st2.paths - float(ref["rays_per_path"])) < 0.03 * float(ref["rays_per_path"])


def _slab_scene(glass_c, e_up, e_down, lookfrom, lookat):
    """A REFR rectangle in the plane z = 0 between two black-bodied emitters at z = +100 (e_up) and z = -100 (e_down);
    a 2x2-pixel camera with a 0.002-degree field of view: pixel (row 0, column 1) is centred on the look direction
    (the reference's footprint is [x - 0.5, x + 0.5) / w, :533-534)."""
    planes = [ptb.rect(ptb.PT_PLANE_XY, -1000, 1000, -1000, 1000, 0.0, c=glass_c, refl=ptb.PT_REFR),
              ptb.rect(ptb.PT_PLANE_XY, -1e4, 1e4, -1e4, 1e4, 100.0, e=e_up, c=(0, 0, 0)),
              ptb.rect(ptb.PT_PLANE_XY, -1e4, 1e4, -1e4, 1e4, -100.0, e=e_down, c=(0, 0, 0))]
    base = ptb.builtin_scene("A", 2, 2)
    cam = ptb.make_camera(lookfrom, lookat, (0, 1, 0), 0.002, 1.0)
    light = ptb.Light()
    light.id = -1
    return ptb.Scene([], planes, [0, 1, 2], light, camera=cam)


def _fresnel(theta, into):
    """Re, Tr of :488-491 for incidence angle theta (nc = 1, nt = 1.5); None = total internal reflection."""
    nnt = 1 / 1.5 if into else 1.5
    cos2t = 1 - nnt * nnt * np.sin(theta) ** 2
    if cos2t < 0:
        return None
    c = 1 - (np.cos(theta) if into else np.sqrt(cos2t))
    re = 0.04 + 0.96 * c ** 5
    return re, 1 - re


@pytest.mark.parametrize("deg,into", [(0.0, True), (30.0, True), (60.0, True), (80.0, True), (20.0, False), (40.0, False), (45.0, False), (70.0, False)])
def test_fresnel_and_total_internal_reflection_known_answers(deg, into):
    th = np.radians(deg)
    glass = (0.9, 0.8, 0.7)
    e_up, e_down = np.array([2.0, 2.0, 2.0]), np.array([5.0, 3.0, 1.0])
    if into:     # from above, travelling down: reflection returns to z = +100, refraction reaches z = -100
        lookfrom, lookat = (0.0, 0.0, 10.0), (10 * np.tan(th), 0.0, 0.0)
        e_refl, e_trans = e_up, e_down
    else:        # from below (the "inside"): n.nl < 0, nnt = 1.5, TIR beyond 41.8 degrees
        lookfrom, lookat = (0.0, 0.0, -10.0), (10 * np.tan(th), 0.0, 0.0)
        e_refl, e_trans = e_down, e_up
    fr = _fresnel(th, into)
    want = np.array(glass) * (e_refl if fr is None else fr[0] * e_refl + fr[1] * e_trans)
    sc = _slab_scene(glass, tuple(e_up), tuple(e_down), lookfrom, lookat)
    with ptb.Context(sc) as c:
        c.render(ptb.params(2, 2, 64, mode=1, seed=1))                           # FP32, split: no noise at all
        got32, st = c.readback()
        assert st.spawned_branches == (0 if fr is None else 2 * 2 * 64)
        c.render(ptb.params(2, 2, 32768, mode=1, seed=2, collect_stats=1))       # FP32, one arm chosen with P: converges to the same
        got_st, sq_st, _ = c.readback(True)
        c.render(ptb.params(2, 2, 16, mode=1, engine=ptb.PT_ENGINE_FP64_ERAND48))  # FP64 replay engine (splits like the reference)
        got64, _ = c.readback()
    assert np.allclose(got64[0, 1], want, rtol=2e-5), (got64[0, 1], want)
    assert np.allclose(got32[0, 1], want, rtol=1e-3), (got32[0, 1], want)
    se = np.sqrt(np.maximum(sq_st[0, 1] / 32768 - got_st[0, 1] ** 2, 0) / 32768)
    assert (np.abs(got_st[0, 1] - want) <= 4 * se + 1e-3 * want).all(), (got_st[0, 1], want, se)


def _furnace_scene(w, h, rho=0.75, emit=0.25):
    base = ptb.builtin_scene("G", w, h)
    spheres = []
    for i in range(10):
        s = base.spheres[i]
        spheres.append(ptb.sphere(s.rad, s.p.tup(), e=(emit,) * 3, c=(rho,) * 3, refl=s.refl))
    return ptb.Scene(spheres, [], [~i for i in range(10)], base.light, base.camera)


def test_white_furnace_energy_balance():
    # every surface (diffuse walls, the mirror, the glass) emits E and has albedo rho: L = E + rho L = E / (1 - rho) = 1
    # wherever the camera looks.  Checks the weights of every arm: cosine sampling, mirror, Re / Tr split, Re/P and Tr/(1-P).
    w = h = 48
    sc = _furnace_scene(w, h)
    with ptb.Context(sc) as c:
        c.render(ptb.params(w, h, 2048, mode=1, seed=5, collect_stats=1))
        m1, sq1, _ = c.readback(True)
        c.render(ptb.params(w, h, 2048, mode=1, seed=6))
        m2, st2 = c.readback()
        c.render(ptb.params(w, h, 16, mode=1, engine=ptb.PT_ENGINE_FP64_ERAND48))
        m64, _ = c.readback()
    assert st2.spawned_branches > 0
    se = np.sqrt(np.maximum(sq1 / 2048 - m1 ** 2, 0) / 2048)
    z = (m1 - 1.0) / np.maximum(se, 1e-12)
    assert (np.abs(z) > 3).mean() <= 0.012
    assert abs(m1.mean() - 1.0) < 2e-3 and abs(m2.mean() - 1.0) < 2e-3, (m1.mean(), m2.mean())
    assert abs(m64.mean() - 1.0) < 0.02
    assert np.abs(m2 - 1.0).max() < 6 * se.max() + 0.02


def test_split_variance_on_the_glass_sphere_matches_the_oracle():
    # The split is a variance-reduction device (:494-495): an engine that always took one arm would be right on average
    # and noisier.  Per-sample variance over the pixels whose primary ray hits the glass sphere (id 8): FP32 engine
    # (empirical, across independently seeded renders) vs the oracle's per-pixel sums of squares.
    ref = np.load(os.path.join(GOLDEN, "converged_G_cos.npz"))
    omean, osq, n_o = ref["mean"].astype(np.float64), ref["sumsq"].astype(np.float64), int(ref["spp"])
    h, w, _ = omean.shape
    sc = ptb.builtin_scene("G", w, h)
    cam = sc.camera
    ys, xs = np.mgrid[0:h, 0:w]
    u, v = (xs / w).ravel(), ((h - ys - 1) / h).ravel()
    o = np.array(cam.origin.tup())
    d = (np.array(cam.lower_left_corner.tup()) + u[:, None] * np.array(cam.horizontal.tup()) + v[:, None] * np.array(cam.vertical.tup())) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([np.broadcast_to(o, d.shape), d], 1)
    n_rend, spp = 64, 16
    with ptb.Context(sc) as c:
        ids = c.intersect(rays, 64)[1].reshape(h, w)
        imgs = []
        for k in range(n_rend):
            c.render(ptb.params(w, h, spp, mode=1, seed=1000 + k))
            imgs.append(c.readback()[0].copy())
        c.render(ptb.params(w, h, n_rend * spp, mode=1, seed=5, collect_stats=1))
        m1, sq1, _ = c.readback(True)
    glass = ids == 8
    assert glass.sum() > 200
    var_split = np.var(np.stack(imgs), axis=0, ddof=1) * spp                  # per-sample variance, production engine
    var_oracle = np.maximum(osq / n_o - omean ** 2, 0)
    var_one_arm = np.maximum(sq1 / (n_rend * spp) - m1 ** 2, 0)
    a, b, one = var_split[glass].mean(), var_oracle[glass].mean(), var_one_arm[glass].mean()
    assert abs(a / b - 1.0) < 0.10, (a, b)
    assert one > 1.05 * a, (one, a)
