"""GPU: the opt-in acceleration structure of the FP32 engine - a uniform grid over the small spheres.

Brute force (reference src/smallpt.cpp:323-335: every primitive, every ray) stays the measured contract; the grid is for
scenes beyond it and is checked against brute force: hit ids on >= 1 M rays against the FP64 engine's literal loop, and - on
a scene that fits both - bit-identical t and ids on 2^20 rays, and the same renders (to one grazing ray in a million), against
the FP32 engine's own brute-force scan."""
import numpy as np
import pytest

from conftest import ptb, orc, room_rays

pytestmark = pytest.mark.gpu


def _sphere_cloud(n, seed, w, h, rmin=0.6, rmax=2.0, glossy=True):
    """The built-in room (six walls + the rectangular light of scene A) filled with n random spheres."""
    rng = np.random.default_rng(seed)
    base = ptb.builtin_scene("A", w, h)
    planes = [base.planes[i] for i in range(7)]
    spheres = []
    for i in range(n):
        p = (rng.uniform(3, 97), rng.uniform(2, 79), rng.uniform(5, 165))
        mt = rng.uniform()
        refl = ptb.PT_DIFF if (mt < .8 or not glossy) else ptb.PT_SPEC if mt < .9 else ptb.PT_REFR
        e = (8.0 * rng.uniform(),) * 3 if rng.uniform() < .05 else (0, 0, 0)
        spheres.append(ptb.sphere(rng.uniform(rmin, rmax), p, e=e, c=tuple(rng.uniform(.2, .9, 3)), refl=refl))
    order = list(range(7)) + [~i for i in range(n)]
    return ptb.Scene(spheres, planes, order, base.light, base.camera)


def test_grid_equals_brute_force_on_a_scene_that_fits_both():
    # config C4's scene (256 spheres + tilted planes): pt_set_acceleration 2 (always the grid) against 0 (the scan)
    w, h, spp = 160, 90, 32
    sc = ptb.builtin_scene("synthetic", w, h)
    rays = room_rays(1 << 20, 11, f32_exact=True)
    out = {}
    with ptb.Context(sc) as c:
        c.set_specialisation(0)                       # the grid runs the ahead-of-time build; compare like with like
        for accel in (0, 2):
            c.set_acceleration(accel)
            t, ids = c.intersect(rays, 32)
            imgs = []
            for mode in (1, 3):
                c.render(ptb.params(w, h, spp, mode=mode, seed=4))
                img, st = c.readback()
                assert st.accel_structure == (1 if accel else 0)
                imgs.append((img.copy(), st.rays, st.shaded_vertices, st.miss_events))
            out[accel] = (t, ids, imgs)
    assert np.array_equal(out[0][1], out[2][1]) and np.array_equal(out[0][0], out[2][0])      # ids and t, 1 M rays
    assert (out[0][1] >= 15).mean() > 0.15                                                   # ... a good part of them on spheres
    # whole renders: the same paths - except where a ray grazes a sphere within FP32 rounding (the scan's conservative
    # bound and the grid's padded cells then decide differently for about one ray in a million): counters to 1e-5, and
    # no more than a handful of pixels touched
    for a, b in zip(out[0][2], out[2][2]):
        assert all(abs(x - y) <= 1e-5 * x for x, y in zip(a[1:], b[1:])), (a[1:], b[1:])
        differing = (a[0] != b[0]).any(axis=2).mean()
        assert differing <= 2e-3, differing
        assert abs(a[0].mean() - b[0].mean()) <= 1e-4 * a[0].mean()


@pytest.mark.parametrize("n", [1024, 4096])
def test_grid_hit_ids_against_fp64_brute_force(n):
    sc = _sphere_cloud(n, 5, 64, 48)
    rays = room_rays(1 << 20, 12, f32_exact=True)
    with ptb.Context(sc) as c:
        t64, id64 = c.intersect(rays, 64)             # the literal loop of :323-335 over all n + 7 objects
        t32, id32 = c.intersect(rays, 32)             # FP32 engine: rectangles one by one, spheres through the grid
        c.render(ptb.params(64, 48, 4, mode=1))
        assert c.stats().accel_structure == 1
    same = id32 == id64
    assert same.mean() >= 0.9995, same.mean()
    hit = same & (id64 >= 0)
    rel = np.abs(t32[hit] - t64[hit]) / t64[hit]
    assert np.median(rel) < 1e-6 and np.quantile(rel, 0.99) < 5e-5 and (id64 >= 7).mean() > 0.15
    # and against the CPU oracle on a sample (the FP64 engine is itself pinned to it, tests/test_gpu_units.py)
    t_o, id_o = orc.oracle_intersect(sc, rays[:20000])
    assert np.array_equal(id_o, id64[:20000])


def test_grid_render_matches_the_oracle_statistically():
    # 1024 spheres, cosine mode: FP32 engine through the grid (one REFR arm: collect_stats) vs the CPU oracle's brute force
    w, h = 40, 30
    sc = _sphere_cloud(1024, 7, w, h)
    with ptb.Context(sc) as c:
        c.render(ptb.params(w, h, 1024, mode=1, seed=3, collect_stats=1))
        mean, sq, st = c.readback(True)
        c.render(ptb.params(w, h, 1024, mode=1, seed=4))
        mean2, st2 = c.readback()
    assert st.accel_structure == 1 and st2.spawned_branches > 0 and st2.split_refusals == 0
    cl, omean, osq, ost = orc.oracle_render(sc, ptb.params(w, h, 512, mode=1, engine=1))
    var = np.maximum(sq / 1024 - mean ** 2, 0) / 1024 + np.maximum(osq / 512 - omean ** 2, 0) / 512
    se = np.sqrt(var)
    informative = se > 1e-9
    for m in (mean, mean2):
        z = (m - omean) / np.maximum(se, 1e-12)
        assert ((np.abs(z) > 3) & informative).sum() / informative.sum() <= 0.015
        assert abs(m.mean() - omean.mean()) < 0.02 * omean.mean()
    assert abs(st2.rays / st2.paths - ost.rays / ost.paths) < 0.03 * ost.rays / ost.paths


def test_grid_edge_cases():
    # one sphere; spheres outside the room; rays that start inside a sphere, on a sphere, outside the grid's box, axis-parallel
    w, h = 16, 12
    base = ptb.builtin_scene("A", w, h)
    planes = [base.planes[i] for i in range(7)]
    spheres = [ptb.sphere(10.0, (50, 40, 80)), ptb.sphere(3.0, (50, 40, 80), c=(.3, .3, .9)), ptb.sphere(5.0, (200, 40, 80)), ptb.sphere(0.25, (20, 10, 30))]
    sc = ptb.Scene(spheres, planes, list(range(7)) + [~i for i in range(4)], base.light, base.camera)
    o = np.array([[50, 40, 80], [50, 40, 86], [50, 40, 168], [2, 2, 2], [50, 40, 80], [20, 10, 40], [98, 80, 169], [50, 40, 70.0]], dtype=np.float64)
    d = np.array([[0, 0, 1], [0, 0, -1], [0, 0, -1], [1, 0, 0], [1, 0, 0], [0, 0, -1], [-.577350269, -.577350269, -.577350269], [0, 1, 0]], dtype=np.float64)
    rays = np.ascontiguousarray(np.concatenate([o, d], 1).astype(np.float32).astype(np.float64))
    with ptb.Context(sc) as c:
        c.set_acceleration(2)
        t32, id32 = c.intersect(rays, 32)
        t64, id64 = c.intersect(rays, 64)
        c.render(ptb.params(w, h, 8, mode=1))
        img, st = c.readback()
    assert np.array_equal(id32, id64), (id32, id64)
    assert np.allclose(t32, t64, rtol=2e-5)
    assert st.accel_structure == 1 and np.isfinite(img).all()
