"""CPU: the host-side mirror of the reference's source-level surface (host/smallpt_b200.hpp)."""
import numpy as np

from conftest import ptb


def test_scene_A_is_the_head_scene():
    sc = ptb.builtin_scene("A")
    assert (sc.n_objects, sc.n_spheres, sc.n_planes) == (17, 0, 17)          # NUMBER_OBJ, src/smallpt.cpp:19
    light = sc.object(6)                                                      # :294, hard-coded id at :467
    assert (light.kind, light.a1, light.a2, light.b1, light.b2, light.k) == (ptb.PT_PLANE_XZ, 32, 68, 63, 96, 81.5)
    assert light.e.tup() == (12, 12, 12) and light.c.tup() == (0, 0, 0)
    kinds = [sc.object(i).kind for i in range(17)]
    assert kinds == [1, 1, 2, 2, 0, 0, 0, 1, 1, 2, 2, 0, 1, 1, 2, 2, 0]      # xy xy yz yz xz xz xz | boxes
    assert sc.flops_per_ray() == 102
    L = sc.light
    assert (L.id, L.x0, L.xw, L.z0, L.zw, L.y, L.area) == (6, 32, 36, 63, 36, 81.6, 1296)   # :365-367,:467,:471


def test_scene_B_is_the_sphere_era_scene():
    sc = ptb.builtin_scene("B")
    assert (sc.n_objects, sc.n_spheres) == (10, 10)
    assert sc.object(0).p.tup() == sc.object(1).p.tup() == (1e5 + 1, 40.8, 81.6)   # the duplicate in src/a.exe
    assert sc.object(9).rad == 600 and sc.object(9).e.tup() == (12, 12, 12)
    assert sc.object(7).c.tup() == (.999, .999, .999)
    assert sc.flops_per_ray() == 200 and sc.light.id == 9


def test_scene_C_and_synthetic():
    sc = ptb.builtin_scene("C")
    assert (sc.n_objects, sc.n_spheres, sc.n_planes) == (9, 2, 7) and sc.flops_per_ray() == 82
    assert [sc.order[i] < 0 for i in range(9)] == [False] * 7 + [True] * 2
    syn = ptb.builtin_scene("synthetic")
    assert syn.n_spheres == 256 and syn.n_planes == 7 + 8
    assert syn.flops_per_ray() == 20 * 256 + 6 * 7 + 31 * 8
    syn2 = ptb.builtin_scene("synthetic")                                      # deterministic generator
    assert all(syn.spheres[i].p.tup() == syn2.spheres[i].p.tup() for i in range(256))
    refl = [syn.spheres[i].refl for i in range(256)]
    assert 0 < refl.count(1) < 60 and 0 < refl.count(2) < 60
    for i in range(7, 15):                                                     # tilted planes: orthonormal frames
        p = syn.planes[i]
        n, s, t = np.array(p.n.tup()), np.array(p.s.tup()), np.array(p.t.tup())
        assert p.kind == ptb.PT_PLANE_TILTED
        assert abs(n @ n - 1) < 1e-12 and abs(s @ s - 1) < 1e-12 and abs(n @ s) < 1e-12 and np.allclose(np.cross(n, s), t)


def test_builtin_camera_values():
    # src/smallpt.cpp:521 with 512x512: u=(1,0,0), v=(0,1,0), w=(0,0,-1), half height tanf(32.5 deg)
    cam = ptb.builtin_scene("A", 512, 512).camera
    hh = float(np.tan(np.float32(np.float32(65 * np.pi / 180) / np.float32(2)), dtype=np.float32))
    assert cam.origin.tup() == (50, 40, 168)
    assert abs(cam.vertical.y - 2 * hh) < 1e-6 and cam.horizontal.y == 0
    assert abs(cam.lower_left_corner.z - 167) < 1e-12
