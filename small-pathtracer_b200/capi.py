"""ctypes mirror of include/ptb200.h — Python is plumbing only (tests, bench, torch.distributed).

Two shared libraries are bound here:
  libptb200.so        the product: CUDA kernels + C ABI (csrc/).  Missing => every compute call raises.
  libsmallpt_host.so  the C++ host surface (scene tables, Camera, toInt, P3 writer) (host/).
The CPU checker (oracle/liboracle.so) is NOT bound here: its loader lives in oracle/pyoracle.py (test infrastructure).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

PT_DIFF, PT_SPEC, PT_REFR = 0, 1, 2
PT_PLANE_XZ, PT_PLANE_XY, PT_PLANE_YZ, PT_PLANE_TILTED = 0, 1, 2, 3
PT_MODE_NEE_REF_RECT, PT_MODE_COS, PT_MODE_UNI, PT_MODE_NEE_CONE_SPHERE = 0, 1, 2, 3
PT_ENGINE_FP32_PHILOX, PT_ENGINE_FP64_ERAND48 = 0, 1
PT_SINCOS_LIBM, PT_SINCOS_DET = 0, 1
MODES = {"nee": 0, "nee-ref": 0, "cos": 1, "uni": 2, "nee-cone": 3}


class Vec3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]

    def tup(self):
        return (self.x, self.y, self.z)


class Sphere(C.Structure):
    _fields_ = [("rad", C.c_double), ("p", Vec3), ("e", Vec3), ("c", Vec3), ("refl", C.c_int), ("_pad", C.c_int)]


class Plane(C.Structure):
    _fields_ = [("kind", C.c_int), ("refl", C.c_int),
                ("a1", C.c_double), ("a2", C.c_double), ("b1", C.c_double), ("b2", C.c_double), ("k", C.c_double),
                ("p0", Vec3), ("n", Vec3), ("s", Vec3), ("t", Vec3), ("hs", C.c_double), ("ht", C.c_double),
                ("e", Vec3), ("c", Vec3)]


class Camera(C.Structure):
    _fields_ = [("origin", Vec3), ("lower_left_corner", Vec3), ("horizontal", Vec3), ("vertical", Vec3)]


class Light(C.Structure):
    _fields_ = [("id", C.c_int), ("_pad", C.c_int), ("x0", C.c_double), ("xw", C.c_double),
                ("z0", C.c_double), ("zw", C.c_double), ("y", C.c_double), ("area", C.c_double)]


class SceneDesc(C.Structure):
    _fields_ = [("spheres", C.POINTER(Sphere)), ("n_spheres", C.c_int),
                ("planes", C.POINTER(Plane)), ("n_planes", C.c_int),
                ("order", C.POINTER(C.c_int)), ("camera", Camera), ("light", Light)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("spp", C.c_int), ("mode", C.c_int),
                ("engine", C.c_int), ("sincos", C.c_int), ("seed", C.c_uint64),
                ("tile_rows", C.c_int), ("rank", C.c_int), ("world", C.c_int), ("max_depth", C.c_int),
                ("queue_capacity", C.c_int), ("collect_stats", C.c_int), ("bounces_per_launch", C.c_int),
                ("sample_offset", C.c_int), ("accumulate", C.c_int), ("owned_rows_only", C.c_int), ("robust_eps", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays_camera", C.c_uint64), ("rays_scatter", C.c_uint64),
                ("rays_shadow", C.c_uint64), ("shaded_vertices", C.c_uint64), ("miss_events", C.c_uint64),
                ("truncated", C.c_uint64), ("kernel_launches", C.c_uint64), ("iterations", C.c_uint64),
                ("max_depth_seen", C.c_uint32), ("specialised", C.c_uint32),
                ("render_ms", C.c_double), ("main_kernel_ms", C.c_double), ("queue_slots_io", C.c_uint64),
                ("tail_ms", C.c_double), ("resolve_ms", C.c_double), ("tail_launches", C.c_uint64),
                ("term_roulette", C.c_uint64), ("term_emitter", C.c_uint64), ("term_light_sample", C.c_uint64),
                ("dropped_contributions", C.c_uint64), ("spawned_branches", C.c_uint64), ("live_at_depth", C.c_uint64 * 64),
                ("split_refusals", C.c_uint64), ("accel_structure", C.c_uint64)]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "live_at_depth" else getattr(self, n)) for n, _ in self._fields_ if not n.startswith("_")}

    @property
    def rays(self):
        return self.rays_camera + self.rays_scatter + self.rays_shadow


class PtError(RuntimeError):
    pass


# ------------------------------------------------------------------------------ host surface
_host = None


def host_lib():
    global _host
    if _host is None:
        path = os.path.join(HERE, "libsmallpt_host.so")
        if not os.path.exists(path):
            raise PtError(f"{path} missing — run `python -c 'import __graft_entry__ as g; g.build()'` or `make`")
        L = C.CDLL(path)
        L.spt_scene_counts.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.spt_scene_fill.argtypes = [C.c_char_p, C.POINTER(Sphere), C.POINTER(Plane), C.POINTER(C.c_int), C.POINTER(Light)]
        L.spt_camera.argtypes = [C.POINTER(C.c_double)] * 3 + [C.c_float, C.c_float, C.POINTER(Camera)]
        L.spt_camera.restype = None
        L.spt_builtin_camera.argtypes = [C.c_int, C.c_int, C.POINTER(Camera)]
        L.spt_builtin_camera.restype = None
        L.spt_plane_tilted.argtypes = [C.POINTER(C.c_double)] * 3 + [C.c_double, C.c_double] + \
            [C.POINTER(C.c_double)] * 2 + [C.c_int, C.POINTER(Plane)]
        L.spt_plane_tilted.restype = None
        L.spt_clamp.argtypes = [C.c_double]
        L.spt_clamp.restype = C.c_double
        L.spt_toInt.argtypes = [C.c_double]
        L.spt_toInt.restype = C.c_int
        L.spt_write_ppm.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int]
        L.spt_write_ppm_binary.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int]
        L.spt_write_pfm.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int]
        L.spt_write_raw64.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.spt_scene_text.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.spt_parse_scene.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.spt_parse_error.restype = C.c_char_p
        L.spt_parsed_fill.argtypes = [C.POINTER(Sphere), C.POINTER(Plane), C.POINTER(C.c_int), C.POINTER(Light), C.c_int, C.c_int, C.POINTER(Camera)]
        L.spt_parsed_fill.restype = None
        _host = L
    return _host


def _dbl3(v):
    return (C.c_double * 3)(*v)


class Scene:
    """A scene table + camera in C-ABI form (keeps the ctypes arrays alive)."""

    def __init__(self, spheres, planes, order, light, camera=None, name="custom"):
        self.name = name
        self.spheres = (Sphere * max(1, len(spheres)))(*spheres)
        self.planes = (Plane * max(1, len(planes)))(*planes)
        self.n_spheres, self.n_planes = len(spheres), len(planes)
        self.order = (C.c_int * len(order))(*order)
        self.light = light
        self.camera = camera

    @property
    def n_objects(self):
        return self.n_spheres + self.n_planes

    def with_camera(self, w, h):
        cam = Camera()
        host_lib().spt_builtin_camera(w, h, C.byref(cam))
        self.camera = cam
        return self

    def desc(self):
        if self.camera is None:
            raise PtError("scene has no camera; call with_camera(w, h)")
        d = SceneDesc()
        d.spheres = C.cast(self.spheres, C.POINTER(Sphere))
        d.n_spheres = self.n_spheres
        d.planes = C.cast(self.planes, C.POINTER(Plane))
        d.n_planes = self.n_planes
        d.order = C.cast(self.order, C.POINTER(C.c_int))
        d.camera = self.camera
        d.light = self.light
        return d

    def object(self, i):
        ref = self.order[i]
        return self.spheres[~ref] if ref < 0 else self.planes[ref]

    # algorithmic intersection FLOPs per ray, SURVEY 8(d): sphere 20, axis rectangle 6, tilted plane 31
    def flops_per_ray(self):
        n_tilt = sum(1 for i in range(self.n_planes) if self.planes[i].kind == PT_PLANE_TILTED)
        return 20 * self.n_spheres + 6 * (self.n_planes - n_tilt) + 31 * n_tilt


def builtin_scene(name, w=512, h=512):
    L = host_lib()
    ns, npl = C.c_int(), C.c_int()
    if L.spt_scene_counts(name.encode(), C.byref(ns), C.byref(npl)):
        raise PtError(f"unknown scene {name!r}")
    sph = (Sphere * max(1, ns.value))()
    pl = (Plane * max(1, npl.value))()
    order = (C.c_int * (ns.value + npl.value))()
    light = Light()
    L.spt_scene_fill(name.encode(), sph, pl, order, C.byref(light))
    sc = Scene(list(sph)[:ns.value], list(pl)[:npl.value], list(order), light, name=name)
    return sc.with_camera(w, h)


def scene_text(name):
    """A built-in scene in the text scene format (host/scene_io.hpp)."""
    L = host_lib()
    n = L.spt_scene_text(name.encode(), None, 0)
    if n < 0:
        raise PtError(f"unknown scene {name!r}")
    buf = C.create_string_buffer(n + 1)
    L.spt_scene_text(name.encode(), buf, n + 1)
    return buf.value.decode()


def parse_scene(text, w=512, h=512):
    """Text scene format -> Scene (camera from the file's `camera` statement, else the reference's, :521)."""
    L = host_lib()
    ns, npl, has_cam = C.c_int(), C.c_int(), C.c_int()
    n = L.spt_parse_scene(text.encode(), C.byref(ns), C.byref(npl), C.byref(has_cam))
    if n < 0:
        raise PtError("scene file: " + L.spt_parse_error().decode())
    sph = (Sphere * max(1, ns.value))()
    pl = (Plane * max(1, npl.value))()
    order = (C.c_int * n)()
    light, cam = Light(), Camera()
    L.spt_parsed_fill(sph, pl, order, C.byref(light), w, h, C.byref(cam))
    return Scene(list(sph)[:ns.value], list(pl)[:npl.value], list(order), light, camera=cam, name="file")


def write_image(path, mean, fmt="ppm"):
    """fmt: 'ppm' (P3, the reference's writer :548-551), 'ppm6' (binary), 'pfm' (float32 linear) or 'raw64'."""
    a = np.ascontiguousarray(mean, dtype=np.float64)
    h, w = a.shape[0], a.shape[1]
    L = host_lib()
    dp = a.ctypes.data_as(C.POINTER(C.c_double))
    rc = {"ppm": lambda: L.spt_write_ppm(path.encode(), dp, w, h), "ppm6": lambda: L.spt_write_ppm_binary(path.encode(), dp, w, h),
          "pfm": lambda: L.spt_write_pfm(path.encode(), dp, w, h),
          "raw64": lambda: L.spt_write_raw64(path.encode(), dp, w, h, 0, b"mean")}[fmt]()
    if rc:
        raise PtError(f"cannot write {path}")


def make_camera(lookfrom, lookat, vup, vfov, aspect):
    cam = Camera()
    host_lib().spt_camera(_dbl3(lookfrom), _dbl3(lookat), _dbl3(vup), vfov, aspect, C.byref(cam))
    return cam


def tilted_plane(p0, n, along, hs, ht, e=(0, 0, 0), c=(.75, .75, .75), refl=PT_DIFF):
    p = Plane()
    host_lib().spt_plane_tilted(_dbl3(p0), _dbl3(n), _dbl3(along), hs, ht, _dbl3(e), _dbl3(c), refl, C.byref(p))
    return p


def rect(kind, a1, a2, b1, b2, k, e=(0, 0, 0), c=(.75, .75, .75), refl=PT_DIFF):
    p = Plane()
    p.kind, p.refl, p.a1, p.a2, p.b1, p.b2, p.k = kind, refl, a1, a2, b1, b2, k
    p.e, p.c = Vec3(*e), Vec3(*c)
    return p


def sphere(rad, p, e=(0, 0, 0), c=(.75, .75, .75), refl=PT_DIFF):
    s = Sphere()
    s.rad, s.p, s.e, s.c, s.refl = rad, Vec3(*p), Vec3(*e), Vec3(*c), refl
    return s


def to_int(x):
    return host_lib().spt_toInt(float(x))


def write_ppm(path, rgb_mean, w, h):
    a = np.ascontiguousarray(rgb_mean, dtype=np.float64)
    if host_lib().spt_write_ppm(path.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), w, h):
        raise PtError(f"cannot write {path}")


def params(w, h, spp, mode=PT_MODE_NEE_REF_RECT, engine=PT_ENGINE_FP32_PHILOX, sincos=PT_SINCOS_LIBM, seed=0,
           tile_rows=0, rank=0, world=1, max_depth=0, queue_capacity=0, collect_stats=0, bounces_per_launch=0,
           sample_offset=0, accumulate=0, owned_rows_only=0, robust_eps=0):
    p = RenderParams()
    p.width, p.height, p.spp, p.mode, p.engine, p.sincos, p.seed = w, h, spp, mode, engine, sincos, seed
    p.tile_rows, p.rank, p.world, p.max_depth = tile_rows, rank, world, max_depth
    p.queue_capacity, p.collect_stats, p.bounces_per_launch = queue_capacity, collect_stats, bounces_per_launch
    p.sample_offset, p.accumulate, p.owned_rows_only, p.robust_eps = sample_offset, accumulate, owned_rows_only, robust_eps
    return p


# ------------------------------------------------------------------------------ product library
_lib = None
LIB_PATH = os.path.join(HERE, "libptb200.so")
EXPORTS = ["pt_scene_upload", "pt_render", "pt_render_multi", "pt_render_into", "pt_readback", "pt_readback_view", "pt_readback_owned", "pt_host_register", "pt_host_unregister", "pt_accum_device_ptr",
           "pt_debug_intersect", "pt_debug_erand48", "pt_debug_philox", "pt_debug_ffma_peak",
           "pt_set_specialisation", "pt_set_acceleration", "pt_debug_specialise", "pt_debug_plan", "pt_debug_stats", "pt_accum_upload", "pt_accum_download", "pt_device_alloc", "pt_device_free", "pt_ipc_export", "pt_ipc_open", "pt_ipc_close", "pt_destroy", "pt_last_error", "pt_version"]


def lib():
    """The CUDA product library.  No fallback: a missing .so is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PtError(f"{LIB_PATH} missing — the CUDA extension is not built; there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.pt_scene_upload.argtypes = [C.POINTER(vp), C.POINTER(SceneDesc), C.c_int]
        L.pt_render.argtypes = [vp, C.POINTER(RenderParams)]
        L.pt_render_multi.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(RenderParams)]
        L.pt_render_into.argtypes = [vp, C.POINTER(RenderParams), vp, vp]
        L.pt_readback.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(Stats)]
        L.pt_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.pt_device_free.argtypes = [vp, vp]
        L.pt_ipc_export.argtypes = [vp, vp, C.c_char_p]
        L.pt_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.pt_ipc_close.argtypes = [vp, vp]
        L.pt_readback_owned.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(Stats)]
        L.pt_host_register.argtypes = [vp, vp, C.c_size_t]
        L.pt_host_unregister.argtypes = [vp, vp]
        L.pt_readback_view.argtypes = [vp, C.POINTER(Stats)]
        L.pt_readback_view.restype = C.POINTER(C.c_double)
        L.pt_accum_upload.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.pt_accum_download.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.pt_accum_device_ptr.argtypes = [vp]
        L.pt_accum_device_ptr.restype = vp
        L.pt_debug_intersect.argtypes = [vp, C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.pt_debug_erand48.argtypes = [vp, C.POINTER(C.c_uint16), C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.pt_debug_philox.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_uint32)]
        L.pt_debug_ffma_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.pt_set_specialisation.argtypes = [vp, C.c_int]
        L.pt_set_acceleration.argtypes = [vp, C.c_int]
        L.pt_debug_stats.argtypes = [vp, C.POINTER(Stats)]
        L.pt_debug_specialise.argtypes = [C.POINTER(SceneDesc), C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
        L.pt_destroy.argtypes = [vp]
        L.pt_destroy.restype = None
        L.pt_last_error.argtypes = [vp]
        L.pt_last_error.restype = C.c_char_p
        L.pt_version.restype = C.c_char_p
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def render_multi(contexts, p):
    """pt_render_multi: contexts of the same scene on different devices render one image; read it back from contexts[0]."""
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    rc = lib().pt_render_multi(arr, len(contexts), C.byref(p))
    for c in contexts:
        c.last = p
    contexts[0]._check(rc, "pt_render_multi")


class PlanInfo(C.Structure):
    _fields_ = [("owned_rows", C.c_uint64), ("owned_pixels", C.c_uint64), ("row_blocks", C.c_uint64), ("block_rows", C.c_uint64),
                ("path_slots", C.c_uint64), ("run_length", C.c_uint64), ("path_indices", C.c_uint64),
                ("layout_flags", C.c_uint32), ("splits_refr_paths", C.c_uint32)]


def plan(scene, p, sm_count=148):
    """Host-only: the FP32 engine's layout of a render (pt_debug_plan) - no GPU needed."""
    d = scene.desc()
    out = PlanInfo()
    L = lib()
    L.pt_debug_plan.argtypes = [C.POINTER(SceneDesc), C.POINTER(RenderParams), C.c_int, C.POINTER(PlanInfo)]
    rc = L.pt_debug_plan(C.byref(d), C.byref(p), sm_count, C.byref(out))
    if rc:
        msg = L.pt_last_error(None)
        raise PtError(f"pt_debug_plan failed ({rc}): {msg.decode() if msg else ''}")
    return out


def specialise(scene, mode=PT_MODE_NEE_REF_RECT, layout_flags=None):
    """Host-only: (specialisation header text, cubin bytes, NVRTC seconds) for `scene` — no GPU needed.
    layout_flags: PlanInfo.layout_flags of the render the module is for (None: a module for any layout)."""
    d = scene.desc()
    buf = C.create_string_buffer(1 << 16)
    nbytes, secs = C.c_size_t(0), C.c_double(0)
    if layout_flags is not None:
        mode = mode | ((layout_flags + 1) << 8)
    rc = lib().pt_debug_specialise(C.byref(d), mode, buf, len(buf), C.byref(nbytes), C.byref(secs))
    if rc:
        msg = lib().pt_last_error(None)
        raise PtError(f"pt_debug_specialise failed ({rc}): {msg.decode() if msg else ''}")
    return buf.value.decode(), nbytes.value, secs.value


class Context:
    """pt_ctx wrapper: upload once, render many."""

    def __init__(self, scene, device=-1):
        self.scene = scene
        self._h = C.c_void_p()
        d = scene.desc()
        rc = lib().pt_scene_upload(C.byref(self._h), C.byref(d), device)
        if rc:
            msg = lib().pt_last_error(None)
            raise PtError(f"pt_scene_upload failed ({rc}): {msg.decode() if msg else ''}")
        self.last = None

    def update_scene(self, scene):
        """pt_scene_upload on the existing context: replace the scene, keep the device buffers."""
        d = scene.desc()
        self._check(lib().pt_scene_upload(C.byref(self._h), C.byref(d), -1), "pt_scene_upload(update)")
        self.scene = scene

    def _check(self, rc, what):
        if rc:
            msg = lib().pt_last_error(self._h)
            raise PtError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def set_specialisation(self, mode):
        """0 = generic kernel, 1 = scene-specialised (NVRTC) kernel for big renders (default), 2 = always."""
        self._check(lib().pt_set_specialisation(self._h, mode), "pt_set_specialisation")

    def set_acceleration(self, mode):
        """0 = brute force only, 1 = uniform grid beyond 512 small spheres (default), 2 = always the grid.  Re-uploads the scene."""
        self._check(lib().pt_set_acceleration(self._h, mode), "pt_set_acceleration")
        self.update_scene(self.scene)

    def render(self, p):
        self.last = p
        self._check(lib().pt_render(self._h, C.byref(p)), "pt_render")

    def render_into(self, p, dev_ptr, stream=0):
        self.last = p
        self._check(lib().pt_render_into(self._h, C.byref(p), C.c_void_p(dev_ptr), C.c_void_p(stream)), "pt_render_into")

    def readback(self, want_sumsq=False, out=None):
        """Mean image (H, W, 3) float64 [+ sum of squares] + stats.  `out`: a caller-owned float64 array to fill (a
        renderer that reads back every frame re-uses one buffer instead of faulting in fresh pages each time)."""
        p = self.last
        n = p.width * p.height * 3
        if out is not None:
            mean = out.reshape(-1)
            assert mean.dtype == np.float64 and mean.size == n and mean.flags.c_contiguous
        else:
            mean = np.empty(n, dtype=np.float64)
        sq = np.empty(n, dtype=np.float64) if want_sumsq else None
        st = Stats()
        self._check(lib().pt_readback(self._h, _dp(mean), _dp(sq) if want_sumsq else None, C.byref(st)), "pt_readback")
        mean = mean.reshape(p.height, p.width, 3)
        if want_sumsq:
            return mean, sq.reshape(p.height, p.width, 3), st
        return mean, st

    def readback_view(self):
        """Mean image as a numpy VIEW of the context's pinned host buffer (no copy; valid until the next call)."""
        p = self.last
        st = Stats()
        ptr = lib().pt_readback_view(self._h, C.byref(st))
        if not ptr:
            self._check(-3, "pt_readback_view")
        return np.ctypeslib.as_array(ptr, shape=(p.height, p.width, 3)), st

    def readback_owned(self, host_image):
        """This rank's rows (means) into a full-size (H, W, 3) float64 host image shared by the ranks (dist.HostImage)."""
        st = Stats()
        assert host_image.dtype == np.float64 and host_image.flags.c_contiguous
        self._check(lib().pt_readback_owned(self._h, _dp(host_image), C.byref(st)), "pt_readback_owned")
        return st

    def host_register(self, arr):
        self._check(lib().pt_host_register(self._h, C.c_void_p(arr.ctypes.data), arr.nbytes), "pt_host_register")

    def host_unregister(self, arr):
        self._check(lib().pt_host_unregister(self._h, C.c_void_p(arr.ctypes.data)), "pt_host_unregister")

    # ---- peer-memory plumbing (fused resolve + gather, dist.SharedImage)
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(lib().pt_device_alloc(self._h, nbytes, C.byref(p)), "pt_device_alloc")
        return p.value

    def device_free(self, ptr):
        self._check(lib().pt_device_free(self._h, C.c_void_p(ptr)), "pt_device_free")

    def ipc_export(self, ptr):
        buf = C.create_string_buffer(64)
        self._check(lib().pt_ipc_export(self._h, C.c_void_p(ptr), buf), "pt_ipc_export")
        return buf.raw

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._check(lib().pt_ipc_open(self._h, C.c_char_p(handle), C.byref(p)), "pt_ipc_open")
        return p.value

    def ipc_close(self, ptr):
        self._check(lib().pt_ipc_close(self._h, C.c_void_p(ptr)), "pt_ipc_close")

    def stats_raw(self):
        """pt_stats without requiring a finished render (debug entries update `specialised`)."""
        st = Stats()
        lib().pt_debug_stats(self._h, C.byref(st))
        return st

    def accum_download(self, want_sumsq=False):
        """Checkpoint: (per-pixel sums (H, W, 3), sums of squares | None, samples per pixel in them)."""
        p = self.last
        n = p.width * p.height * 3
        s = np.empty(n, dtype=np.float64)
        sq = np.empty(n, dtype=np.float64) if want_sumsq else None
        done = C.c_int(0)
        self._check(lib().pt_accum_download(self._h, _dp(s), _dp(sq) if want_sumsq else None, C.byref(done)), "pt_accum_download")
        shape = (p.height, p.width, 3)
        return s.reshape(shape), (sq.reshape(shape) if want_sumsq else None), done.value

    def accum_upload(self, sums, spp_done, sumsq=None):
        """Resume: load a checkpoint; follow with render(params(..., sample_offset=spp_done, accumulate=1))."""
        a = np.ascontiguousarray(sums, dtype=np.float64)
        h, w = a.shape[0], a.shape[1]
        q = np.ascontiguousarray(sumsq, dtype=np.float64) if sumsq is not None else None
        self._check(lib().pt_accum_upload(self._h, w, h, _dp(a), _dp(q) if q is not None else None, spp_done), "pt_accum_upload")
        self.last = params(w, h, spp_done)

    def stats(self):
        st = Stats()
        self._check(lib().pt_readback(self._h, None, None, C.byref(st)), "pt_readback")
        return st

    def accum_ptr(self):
        return lib().pt_accum_device_ptr(self._h)

    def intersect(self, rays_od, precision=64):
        r = np.ascontiguousarray(rays_od, dtype=np.float64).reshape(-1, 6)
        n = r.shape[0]
        t = np.empty(n, dtype=np.float64)
        ids = np.empty(n, dtype=np.int32)
        self._check(lib().pt_debug_intersect(self._h, _dp(r), n, precision, _dp(t), ids.ctypes.data_as(C.POINTER(C.c_int))),
                    "pt_debug_intersect")
        return t, ids

    def erand48(self, seeds, draws):
        s = np.ascontiguousarray(seeds, dtype=np.uint16).reshape(-1, 3)
        out = np.empty((s.shape[0], draws), dtype=np.float64)
        self._check(lib().pt_debug_erand48(self._h, s.ctypes.data_as(C.POINTER(C.c_uint16)), s.shape[0], draws, _dp(out)),
                    "pt_debug_erand48")
        return out

    def philox(self, ctr, key):
        c = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
        k = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
        out = np.empty_like(c)
        u32p = C.POINTER(C.c_uint32)
        self._check(lib().pt_debug_philox(self._h, c.ctypes.data_as(u32p), k.ctypes.data_as(u32p), c.shape[0],
                                          out.ctypes.data_as(u32p)), "pt_debug_philox")
        return out

    def ffma_peak(self):
        tf, mhz = C.c_double(), C.c_double()
        self._check(lib().pt_debug_ffma_peak(self._h, C.byref(tf), C.byref(mhz)), "pt_debug_ffma_peak")
        return tf.value, mhz.value

    def close(self):
        if self._h:
            lib().pt_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
