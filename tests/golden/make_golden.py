#!/usr/bin/env python3
"""Generate the committed golden fixtures from the REFERENCE ITSELF (run in the build container, where
/root/reference exists):

  ref_render.npz   per-pixel FP64 outputs of oracle/_ref/smallpt_ref (the reference's src/smallpt.cpp +
                   patches P0-P6) for scenes A/B/C x modes NEE/COS/UNI x sincos libm/det, 32x24 @ 4 spp.
  ref_units.npz    outputs of the UNMODIFIED reference functions via oracle/_ref/librefharness.so:
                   Camera, rect[i]->intersect, intersect(), hittingPoint, Sphere::intersect/normal,
                   random_scattering, erand48, clamp/toInt.
  converged_*.npz  4096-spp images (mean + per-pixel sum of squares, 128x128) of the C oracle, which
                   tests/test_oracle_golden.py pins bit-for-bit to smallpt_ref; used by the 3-sigma gate.

Usage: python tests/golden/make_golden.py [--converged]
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "smallpt_ref")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "librefharness.so")


def ref_render():
    out = {}
    w, h, spp = 32, 24, 4
    with tempfile.TemporaryDirectory() as tmp:
        for scene in "ABC":
            for mode in (0, 1, 2):
                for det in (0, 1):
                    prefix = os.path.join(tmp, f"{scene}{mode}{det}")
                    subprocess.check_call([REF_BIN, str(spp), str(mode), scene, str(w), str(h), prefix, str(det), "1"],
                                          stdout=subprocess.DEVNULL)
                    for kind in ("clamped", "mean", "sumsq"):
                        out[f"{scene}_{mode}_{det}_{kind}"] = np.fromfile(f"{prefix}.{kind}.f64").reshape(h, w, 3)
                    if scene == "A" and mode == 0 and det == 0:
                        with open(prefix + ".ppm", "rb") as f:
                            out["A_0_0_ppm"] = np.frombuffer(f.read(), dtype=np.uint8)
    out["meta_whs"] = np.array([w, h, spp])
    np.savez_compressed(os.path.join(HERE, "ref_render.npz"), **out)
    print("wrote ref_render.npz", len(out), "arrays")


def ref_units():
    L = C.CDLL(HARNESS)
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(20191)
    out = {}
    # rays: origins inside the room, on surfaces, and a few degenerate directions
    n = 1500
    o = np.stack([rng.uniform(1, 99, n), rng.uniform(0, 81.6, n), rng.uniform(0, 170, n)], 1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # origins exactly on planes (self-hit cases), axis-parallel directions (division by zero -> inf/NaN)
    o[:50, 1] = 0.0
    o[50:100, 1] = 81.6
    o[100:150, 0] = 1.0
    o[150:200, 2] = 0.0
    d[200:210] = [0, 1, 0]
    d[210:220] = [1, 0, 0]
    d[220:230] = [0, 0, -1]
    d[230:240, 1] = 0.0
    rays = np.ascontiguousarray(np.concatenate([o, d], 1))
    out["rays"] = rays
    nobj = L.ref_number_obj()
    per_obj = np.zeros((nobj, n))
    L.ref_object_intersect.argtypes = [C.c_int, dp, C.c_int, dp]
    for i in range(nobj):
        L.ref_object_intersect(i, rays.ctypes.data_as(dp), n, per_obj[i].ctypes.data_as(dp))
    out["rect_t"] = per_obj
    t = np.zeros(n)
    ids = np.zeros(n, dtype=np.int32)
    L.ref_scene_intersect.argtypes = [dp, C.c_int, dp, C.POINTER(C.c_int)]
    L.ref_scene_intersect(rays.ctypes.data_as(dp), n, t.ctypes.data_as(dp), ids.ctypes.data_as(C.POINTER(C.c_int)))
    out["scene_t"], out["scene_id"] = t, ids
    # hittingPoint
    hp = np.zeros((n, 3))
    hid = np.zeros(n, dtype=np.int32)
    L.ref_hitting_point.argtypes = [dp, dp, C.POINTER(C.c_int)]
    for k in range(n):
        idc = C.c_int()
        L.ref_hitting_point(rays[k].ctypes.data_as(dp), hp[k].ctypes.data_as(dp), C.byref(idc))
        hid[k] = idc.value
    out["hit_point"], out["hit_id"] = hp, hid
    # normals of every object for the first 64 rays
    nrm = np.zeros((nobj, 64, 10))
    L.ref_object_normal.argtypes = [C.c_int, dp, dp, dp]
    for i in range(nobj):
        for k in range(64):
            L.ref_object_normal(i, rays[k].ctypes.data_as(dp), hp[k].ctypes.data_as(dp), nrm[i, k].ctypes.data_as(dp))
    out["normals"] = nrm
    # spheres: the two of :297-298, a huge wall sphere and the sphere-era light
    spheres = np.array([[16.5, 27, 16.5, 47], [16.5, 73, 16.5, 78], [1e5, 1e5 + 1, 40.8, 81.6], [600, 50, 681.6 - .27, 81.6]])
    st = np.zeros((len(spheres), n))
    sn = np.zeros((len(spheres), 64, 3))
    L.ref_sphere_intersect.argtypes = [C.c_double, dp, dp, C.c_int, dp]
    L.ref_sphere_normal.argtypes = [C.c_double, dp, dp, dp, dp]
    for i, s in enumerate(spheres):
        p = np.ascontiguousarray(s[1:])
        L.ref_sphere_intersect(s[0], p.ctypes.data_as(dp), rays.ctypes.data_as(dp), n, st[i].ctypes.data_as(dp))
        for k in range(64):
            x = np.ascontiguousarray(rays[k, :3] + rays[k, 3:] * 10.0)
            L.ref_sphere_normal(s[0], p.ctypes.data_as(dp), rays[k].ctypes.data_as(dp), x.ctypes.data_as(dp), sn[i, k].ctypes.data_as(dp))
    out["spheres"], out["sphere_t"], out["sphere_n"] = spheres, st, sn
    # camera
    cams = []
    cam_args = [((50, 40, 168), (50, 40, 5), (0, 1, 0), 65.0, 1.0), ((50, 40, 168), (50, 40, 5), (0, 1, 0), 65.0, 3840 / 2160),
                ((10, 20, 30), (-3, 7, 1), (0.1, 1, 0.2), 40.0, 4 / 3)]
    L.ref_camera.argtypes = [dp, dp, dp, C.c_float, C.c_float, dp]
    for lf, la, vu, fov, asp in cam_args:
        o12 = np.zeros(12)
        L.ref_camera(np.array(lf, float).ctypes.data_as(dp), np.array(la, float).ctypes.data_as(dp),
                     np.array(vu, float).ctypes.data_as(dp), fov, np.float32(asp), o12.ctypes.data_as(dp))
        cams.append(o12)
    out["cam_args"] = np.array([list(a[0]) + list(a[1]) + list(a[2]) + [a[3], np.float32(a[4])] for a in cam_args])
    out["cams"] = np.array(cams)
    # random_scattering + erand48
    L.ref_erand48.restype = C.c_double
    xi = (C.c_uint16 * 3)(0, 0, 125)
    out["erand48_0_0_125"] = np.array([L.ref_erand48(xi) for _ in range(64)])
    L.ref_random_scattering.argtypes = [dp, C.POINTER(C.c_uint16), dp]
    normals = np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [0, 0, -1], [0.6, 0.0, 0.8], [0.05, 0.99, 0.13]], float)
    normals[5] /= np.linalg.norm(normals[5])
    rs = np.zeros((len(normals), 32, 3))
    for i, nl in enumerate(normals):
        xi = (C.c_uint16 * 3)(1, 2, 3 + i)
        for k in range(32):
            L.ref_random_scattering(nl.ctypes.data_as(dp), xi, rs[i, k].ctypes.data_as(dp))
    out["scatter_normals"], out["scatter_dirs"] = normals, rs
    L.ref_toInt.argtypes = [C.c_double]
    xs = np.concatenate([np.linspace(-0.5, 1.5, 201), rng.uniform(0, 1, 200)])
    out["toint_x"] = xs
    out["toint_y"] = np.array([L.ref_toInt(float(x)) for x in xs], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_units.npz"), **out)
    print("wrote ref_units.npz")


def converged():
    from _pkg import ptb
    from oracle import pyoracle as orc
    w = h = 128
    spp = 4096
    for scene in "AB":
        sc = ptb.builtin_scene(scene, w, h)
        for mode, name in ((0, "nee"), (1, "cos"), (2, "uni")):
            p = ptb.params(w, h, spp, mode=mode, engine=1)
            cl, mean, sq, st = orc.oracle_render(sc, p)
            np.savez_compressed(os.path.join(HERE, f"converged_{scene}_{name}.npz"), mean=mean.astype(np.float32),
                                sumsq=sq.astype(np.float32), spp=np.array(spp), rays_per_path=np.array(st.rays / st.paths),
                                miss_per_path=np.array(st.miss_events / st.paths))
            print("wrote converged", scene, name, "%.1f s" % (st.render_ms / 1e3), flush=True)


def converged_glossy():
    """Scene G (mirror + glass spheres, the SPEC / REFR arms of src/smallpt.cpp:481-495, with the depth <= 2 split) from the
    C oracle: converged cosine-mode image + per-pixel sums of squares, for the FP32 engine's 3-sigma and variance gates."""
    from _pkg import ptb
    from oracle import pyoracle as orc
    w = h = 128
    spp = 4096
    sc = ptb.builtin_scene("G", w, h)
    cl, mean, sq, st = orc.oracle_render(sc, ptb.params(w, h, spp, mode=1, engine=1))
    np.savez_compressed(os.path.join(HERE, "converged_G_cos.npz"), mean=mean.astype(np.float32), sumsq=sq.astype(np.float32),
                        spp=np.array(spp), rays_per_path=np.array(st.rays / st.paths), miss_per_path=np.array(st.miss_events / st.paths))
    print("wrote converged G cos", "%.1f s" % (st.render_ms / 1e3), flush=True)


if __name__ == "__main__":
    if "--glossy-only" in sys.argv:
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        converged_glossy()
        sys.exit(0)
    if not (os.path.exists(REF_BIN) and os.path.exists(HARNESS)):
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run `make -C oracle`")
    ref_render()
    ref_units()
    if "--converged" in sys.argv:
        converged()
        converged_glossy()
