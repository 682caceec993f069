"""CPU: the oracle's building blocks against the UNMODIFIED reference functions (golden vectors produced by
oracle/_ref/librefharness.so, see tests/golden/make_golden.py) and against published known answers."""
import ctypes as C

import numpy as np
import pytest

from conftest import ptb, orc


def test_erand48_kat_survey_appendix_f():
    # seed Xi={0,0,125} (row y=5)
    L = orc.load_oracle()
    xi = (C.c_uint16 * 3)(0, 0, 125)
    got = [L.oracle_erand48(xi) for _ in range(4)]
    assert got == [0.51258850097660158, 0.084069501119962808, 0.089986104133675582, 0.59578631930072845]


def test_erand48_matches_reference(golden_units):
    L = orc.load_oracle()
    xi = (C.c_uint16 * 3)(0, 0, 125)
    got = np.array([L.oracle_erand48(xi) for _ in range(64)])
    assert np.array_equal(got, golden_units["erand48_0_0_125"])


def test_philox_kat():
    # Random123 known answers for philox4x32-10
    L = orc.load_oracle()
    cases = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
             ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
             ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
              [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in cases:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        L.oracle_philox4x32_10(c, k, o)
        assert list(o) == want


def test_det_sincos_accuracy_and_quadrants():
    L = orc.load_oracle()
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.uniform(0, 2 * np.pi, 20000), [0.0, np.pi / 2, np.pi, 1.5 * np.pi, 2 * np.pi, np.pi / 4]])
    s, c = C.c_double(), C.c_double()
    worst = 0.0
    for x in a:
        L.oracle_det_sincos(float(x), C.byref(s), C.byref(c))
        for got, want in ((s.value, np.sin(x)), (c.value, np.cos(x))):
            ulp = np.spacing(abs(want)) if want != 0 else 1e-300
            worst = max(worst, abs(got - want) / max(ulp, 2.3e-16))
        assert abs(s.value ** 2 + c.value ** 2 - 1) < 1e-15
    assert worst <= 2.0   # <= 2 ulp (measured: 1)


def test_camera_matches_reference(golden_units):
    L = orc.load_oracle()
    for args, want in zip(golden_units["cam_args"], golden_units["cams"]):
        lf, la, vu = (ptb.Vec3(*args[0:3]), ptb.Vec3(*args[3:6]), ptb.Vec3(*args[6:9]))
        cam = ptb.Camera()
        L.oracle_camera(C.byref(lf), C.byref(la), C.byref(vu), C.c_float(args[9]), C.c_float(args[10]), C.byref(cam))
        got = np.array(cam.origin.tup() + cam.lower_left_corner.tup() + cam.horizontal.tup() + cam.vertical.tup())
        assert np.array_equal(got, want)


def test_host_camera_matches_reference(golden_units):
    # the C++ host Camera (host/smallpt_b200.hpp) — the one the product uses
    for args, want in zip(golden_units["cam_args"], golden_units["cams"]):
        cam = ptb.make_camera(args[0:3], args[3:6], args[6:9], float(args[9]), float(args[10]))
        got = np.array(cam.origin.tup() + cam.lower_left_corner.tup() + cam.horizontal.tup() + cam.vertical.tup())
        assert np.array_equal(got, want)


def test_scene_intersect_matches_reference(golden_units):
    sc = ptb.builtin_scene("A")
    t, ids = orc.oracle_intersect(sc, golden_units["rays"])
    assert np.array_equal(ids, golden_units["scene_id"])
    assert np.array_equal(t, golden_units["scene_t"])     # bit-exact, incl. 1e20 on misses


def test_each_rectangle_matches_reference(golden_units):
    # single-object scenes: isolates Rectangle_*::intersect incl. the no-epsilon / NaN / float-rounding quirks
    full = ptb.builtin_scene("A")
    rays = golden_units["rays"]
    for i in range(full.n_objects):
        one = ptb.Scene([], [full.planes[full.order[i]]], [0], full.light, full.camera)
        t, ids = orc.oracle_intersect(one, rays)
        want = golden_units["rect_t"][i]
        # reference returns NaN / inf / 0 for non-hits; intersect() (:328) keeps only finite 0 < d < 1e20
        hit = np.nan_to_num(want, nan=0.0, posinf=0.0) != 0
        hit &= np.nan_to_num(want, nan=1e30, posinf=1e30) < 1e20
        assert np.array_equal(ids >= 0, hit), i
        assert np.array_equal(t[hit], want[hit]), i


def test_spheres_match_reference(golden_units):
    rays = golden_units["rays"]
    light = ptb.Light()
    cam = ptb.builtin_scene("A").camera
    for s, want in zip(golden_units["spheres"], golden_units["sphere_t"]):
        one = ptb.Scene([ptb.sphere(s[0], s[1:4])], [], [~0], light, cam)
        t, ids = orc.oracle_intersect(one, rays)
        hit = want != 0
        assert np.array_equal(ids >= 0, hit)
        assert np.array_equal(t[hit], want[hit])


def test_toint_and_clamp(golden_units):
    L = orc.load_oracle()
    got = np.array([L.oracle_toInt(float(x)) for x in golden_units["toint_x"]])
    assert np.array_equal(got, golden_units["toint_y"])
    got_host = np.array([ptb.to_int(float(x)) for x in golden_units["toint_x"]])
    assert np.array_equal(got_host, golden_units["toint_y"])


def test_tilted_plane_reduces_to_axis_rectangle():
    # parity unpinned (no tilted plane in the reference source): an axis-aligned instance of the tilted
    # primitive must agree with Rectangle_xz away from its edges and epsilon band.
    from conftest import room_rays
    base = ptb.builtin_scene("A")
    rect = ptb.rect(ptb.PT_PLANE_XZ, 10, 60, 20, 90, 40.0)
    tilt = ptb.tilted_plane((35, 40, 55), (0, 1, 0), (1, 0, 0), 25, 35)
    rays = room_rays(20000, 5, f32_exact=False)
    ta, ia = orc.oracle_intersect(ptb.Scene([], [rect], [0], base.light, base.camera), rays)
    tb, ib = orc.oracle_intersect(ptb.Scene([], [tilt], [0], base.light, base.camera), rays)
    hp = rays[:, :3] + rays[:, 3:] * np.where(ia >= 0, ta, 0)[:, None]
    interior = (ia >= 0) & (ta > 1e-3) & (np.abs(hp[:, 0] - 35) < 24.99) & (np.abs(hp[:, 2] - 55) < 34.99)
    assert interior.sum() > 1000
    assert np.array_equal(ib[interior], ia[interior])
    assert np.allclose(tb[interior], ta[interior], rtol=1e-12, atol=0)
    outside = (ia < 0) & (ib >= 0)
    assert outside.sum() <= 5    # only float-rounded edge cases may differ


def test_random_scattering_matches_reference(golden_units):
    L = orc.load_oracle()
    dp = C.POINTER(C.c_double)
    L.oracle_random_scattering.argtypes = [dp, C.POINTER(C.c_uint16), C.c_int, C.c_int, dp]
    L.oracle_random_scattering.restype = None
    for i, (nl, want) in enumerate(zip(golden_units["scatter_normals"], golden_units["scatter_dirs"])):
        xi = (C.c_uint16 * 3)(1, 2, 3 + i)
        nl = np.ascontiguousarray(nl)
        for k in range(32):
            out = np.zeros(3)
            L.oracle_random_scattering(nl.ctypes.data_as(dp), xi, ptb.PT_MODE_COS, ptb.PT_SINCOS_LIBM, out.ctypes.data_as(dp))
            assert np.array_equal(out, want[k]), (i, k)
