"""pyoracle.py — ctypes loader of the CPU checker (oracle/liboracle.so, oracle/liboracle_fp32.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (small-pathtracer_b200/) never does; it has no CPU path.
The struct mirrors (SceneDesc, RenderParams, ...) are the product's own ctypes definitions of include/ptb200.h.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from _pkg import ptb  # noqa: E402

SceneDesc, RenderParams, Vec3, Camera, PtError = ptb.SceneDesc, ptb.RenderParams, ptb.Vec3, ptb.Camera, ptb.PtError


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays_camera", C.c_uint64), ("rays_scatter", C.c_uint64),
                ("rays_shadow", C.c_uint64), ("shaded_vertices", C.c_uint64), ("miss_events", C.c_uint64),
                ("truncated", C.c_uint64), ("max_depth_seen", C.c_uint32), ("threads", C.c_uint32),
                ("render_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}

    @property
    def rays(self):
        return self.rays_camera + self.rays_scatter + self.rays_shadow



_oracle = {}


def load_oracle(fp32=False):
    """CPU oracle (oracle/liboracle.so)."""
    key = "fp32" if fp32 else "fp64"
    if key not in _oracle:
        path = os.path.join(ROOT, "oracle", "liboracle_fp32.so" if fp32 else "liboracle.so")
        if not os.path.exists(path):
            raise PtError(f"{path} missing — run `make -C oracle`")
        L = C.CDLL(path)
        L.oracle_render.argtypes = [C.POINTER(SceneDesc), C.POINTER(RenderParams)] + [C.POINTER(C.c_double)] * 3 + \
            [C.POINTER(OracleStats)]
        L.oracle_intersect.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.oracle_erand48.argtypes = [C.POINTER(C.c_uint16)]
        L.oracle_erand48.restype = C.c_double
        L.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.oracle_philox4x32_10.restype = None
        L.oracle_camera.argtypes = [C.POINTER(Vec3)] * 3 + [C.c_float, C.c_float, C.POINTER(Camera)]
        L.oracle_camera.restype = None
        L.oracle_toInt.argtypes = [C.c_double]
        L.oracle_write_ppm.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int]
        L.oracle_det_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_det_sincos.restype = None
        _oracle[key] = L
    return _oracle[key]


def oracle_render(scene, p, fp32=False):
    """Returns (clamped c[], unclamped mean, sumsq, OracleStats) as (h, w, 3) arrays."""
    L = load_oracle(fp32)
    n = p.width * p.height * 3
    cl, mean, sq = (np.zeros(n, dtype=np.float64) for _ in range(3))
    st = OracleStats()
    d = scene.desc()
    rc = L.oracle_render(C.byref(d), C.byref(p), _dp(cl), _dp(mean), _dp(sq), C.byref(st))
    if rc:
        raise PtError(f"oracle_render failed ({rc})")
    shp = (p.height, p.width, 3)
    return cl.reshape(shp), mean.reshape(shp), sq.reshape(shp), st


def oracle_intersect(scene, rays_od, fp32=False):
    L = load_oracle(fp32)
    r = np.ascontiguousarray(rays_od, dtype=np.float64).reshape(-1, 6)
    t = np.empty(r.shape[0], dtype=np.float64)
    ids = np.empty(r.shape[0], dtype=np.int32)
    d = scene.desc()
    L.oracle_intersect(C.byref(d), _dp(r), r.shape[0], _dp(t), ids.ctypes.data_as(C.POINTER(C.c_int)))
    return t, ids
