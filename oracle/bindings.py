"""ctypes bindings of the CPU checkers -- TEST INFRASTRUCTURE ONLY.

Two shared libraries expose the same C surface under different prefixes:

* ``oracle/liboracle.so``  (prefix ``orc_``): this repo's own restatement of the reference algorithm (oracle.cpp);
* ``oracle/_ref/libref.so`` / ``libref_strict.so`` (prefix ``ref_``): the unmodified reference compiled from
  /root/reference/tp2 (ref_harness.cpp).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs import this
module.  The product package (raytracercpp_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


class RtMaterial(C.Structure):
    _fields_ = [
        ("ambient_coeff", C.c_float * 3),
        ("diffuse", C.c_float * 3),
        ("specular", C.c_float * 3),
        ("emission", C.c_float * 3),
        ("reflection", C.c_float),
        ("roughness", C.c_float),
        ("ns", C.c_float),
        ("specular_threshold", C.c_float),
    ]


SETTINGS_FIELDS = [
    "image_width", "image_height", "enable_ssaa", "ssaa_factor", "hybrid_rasterization_tracing",
    "shading_method", "compute_shadows", "max_recursion_depth", "enable_bvh", "bvh_max_depth",
    "bvh_leaf_object_count", "enable_ssao", "enable_ambient", "enable_diffuse", "enable_specular",
    "enable_emissive", "rough_reflections_sample_count", "enable_ao_mapping", "enable_diffuse_mapping",
    "enable_normal_mapping", "enable_displacement_mapping", "enable_roughness_mapping", "enable_skysphere",
    "enable_skybox",
]


class RtSettings(C.Structure):
    _fields_ = [(n, C.c_int32) for n in SETTINGS_FIELDS] + [("rng_seed", C.c_uint32), ("displacement_mapping_strength", C.c_float),
                                                              ("parallax_mapping_steps", C.c_int32),
                                                              ("ssao_sample_count", C.c_int32), ("ssao_radius", C.c_float), ("ssao_amount", C.c_float),
                                                             ("enable_clipping", C.c_int32)]


def default_settings(**kw) -> RtSettings:
    """RenderSettings defaults, rendererSettings.h:29-102."""
    s = RtSettings()
    s.image_width = 1024
    s.image_height = 1024
    s.enable_ssaa = 0
    s.ssaa_factor = 2
    s.shading_method = 0
    s.compute_shadows = 0
    s.max_recursion_depth = 5
    s.enable_bvh = 1
    s.bvh_max_depth = 12
    s.bvh_leaf_object_count = 40
    s.enable_ambient = s.enable_diffuse = s.enable_specular = s.enable_emissive = 1
    s.rough_reflections_sample_count = 3
    s.rng_seed = 0
    s.displacement_mapping_strength = 0.02
    s.parallax_mapping_steps = 32
    s.ssao_sample_count, s.ssao_radius, s.ssao_amount = 64, 0.5, 1.0          # rendererSettings.h:69-73
    s.enable_clipping = 1                                                     # rendererSettings.h:40
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, float(v) if k in ("displacement_mapping_strength", "ssao_radius", "ssao_amount") else int(v))
    return s


def materials_array(mats) -> C.Array:
    arr = (RtMaterial * len(mats))()
    for i, m in enumerate(mats):
        for key in ("ambient_coeff", "diffuse", "specular", "emission"):
            v = m.get(key, (1.0, 1.0, 1.0) if key == "ambient_coeff" else (0.0, 0.0, 0.0))
            for j in range(3):
                getattr(arr[i], key)[j] = float(v[j])
        arr[i].reflection = float(m.get("reflection", 0.0))
        arr[i].roughness = float(m.get("roughness", 0.0))
        arr[i].ns = float(m.get("ns", 0.0))
        arr[i].specular_threshold = float(m.get("specular_threshold", 0.0))
    return arr


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int32)
UP = C.POINTER(C.c_uint32)
BP = C.POINTER(C.c_uint8)


def lib_path(kind: str) -> Path:
    return {"oracle": HERE / "liboracle.so", "ref": HERE / "_ref" / "libref.so",
            "ref_strict": HERE / "_ref" / "libref_strict.so"}[kind]


def available(kind: str) -> bool:
    return lib_path(kind).exists()


class CpuTracer:
    """One of the CPU checkers.  kind in {"oracle", "ref", "ref_strict"}."""

    def __init__(self, kind: str = "oracle"):
        self.kind = kind
        self.prefix = "orc_" if kind == "oracle" else "ref_"
        path = lib_path(kind)
        if not path.exists():
            raise FileNotFoundError(f"{path} not built (make -C oracle)")
        self.lib = C.CDLL(str(path), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        L, p = self.lib, self.prefix
        self._fn = lambda name: getattr(L, p + name)
        f = self._fn
        f("bvh_create").restype = C.c_void_p
        f("bvh_create").argtypes = [FP, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]
        f("bvh_destroy").argtypes = [C.c_void_p]
        f("bvh_stats").argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        f("bvh_intersect").restype = C.c_double
        f("bvh_intersect").argtypes = [C.c_void_p, FP, FP, C.c_size_t, IP, FP, FP, FP, C.c_int]
        f("triangle_intersect").restype = C.c_int
        f("triangle_intersect").argtypes = [FP, FP, FP, FP, FP, FP]
        f("renderer_create").restype = C.c_void_p
        f("renderer_destroy").argtypes = [C.c_void_p]
        f("renderer_configure").argtypes = [C.c_void_p, C.POINTER(RtSettings), C.c_float]
        f("renderer_set_triangles").argtypes = [C.c_void_p, FP, FP, IP, C.c_size_t]
        f("renderer_set_materials").argtypes = [C.c_void_p, C.POINTER(RtMaterial), C.c_size_t]
        f("renderer_set_texture_f32").argtypes = [C.c_void_p, C.c_int, FP, C.c_int, C.c_int]
        f("renderer_set_texture_u8").argtypes = [C.c_void_p, C.c_int, BP, C.c_int, C.c_int]
        f("renderer_set_camera_transform").argtypes = [C.c_void_p, FP]
        f("renderer_set_light").argtypes = [C.c_void_p, FP]
        f("renderer_add_sphere").argtypes = [C.c_void_p, FP, C.c_float, C.c_int]
        f("renderer_add_plane").argtypes = [C.c_void_p, FP, FP, C.c_int]
        f("camera_matrices").argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, FP, FP]
        f("transform_inverse").argtypes = [FP, FP]
        f("renderer_render").restype = C.c_double
        f("renderer_render").argtypes = [C.c_void_p, UP, C.c_int]
        f("renderer_trace_rows").restype = C.c_double
        f("renderer_trace_rows").argtypes = [C.c_void_p, FP, UP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        f("downscale").argtypes = [UP, C.c_int, C.c_int, C.c_int, UP]
        f("omp_max_threads").restype = C.c_int
        if kind != "oracle":
            f("last_hit_count").restype = C.c_longlong
            f("renderer_render_ssao").restype = C.c_double
            f("renderer_render_ssao").argtypes = [C.c_void_p, UP, C.c_uint, UP, C.c_int]
            f("renderer_raster").restype = C.c_double
            f("renderer_raster").argtypes = [C.c_void_p, UP, C.c_uint, UP, C.c_int]
            f("load_obj").restype = C.c_int
            f("load_obj").argtypes = [C.c_char_p, FP, FP, FP, IP, C.POINTER(RtMaterial), C.c_int, IP]
        else:
            f("renderer_render_ssao").argtypes = [C.c_void_p, UP, C.c_int, C.c_int, UP]
            f("renderer_raster").argtypes = [C.c_void_p, UP, C.c_int, UP, C.POINTER(C.c_uint64)]
            f("bvh_count").restype = C.c_double
            f("bvh_count").argtypes = [C.c_void_p, FP, FP, C.c_size_t, C.POINTER(C.c_uint64), C.c_int]
            f("renderer_count_rows").argtypes = [C.c_void_p, FP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_int]

    # ---- misc ------------------------------------------------------------------------------------------
    def max_threads(self) -> int:
        return int(self._fn("omp_max_threads")())

    def load_obj(self, path: str, transform16=None):
        """Reference OBJ loader (ref libs only): returns xyz9, uv6, mat, [material dicts]."""
        tr = _f32(transform16) if transform16 is not None else None
        n_mats = C.c_int(0)
        n = self._fn("load_obj")(path.encode(), _ptr(tr, C.c_float), None, None, None, None, 0, C.byref(n_mats))
        if n < 0:
            raise IOError(path)
        xyz9 = np.zeros((n, 9), np.float32)
        uv6 = np.zeros((n, 6), np.float32)
        mat = np.zeros(n, np.int32)
        mats = (RtMaterial * n_mats.value)()
        self._fn("load_obj")(path.encode(), _ptr(tr, C.c_float), _ptr(xyz9, C.c_float), _ptr(uv6, C.c_float),
                             _ptr(mat, C.c_int32), mats, n_mats.value, C.byref(n_mats))
        out = []
        for m in mats:
            out.append(dict(ambient_coeff=tuple(m.ambient_coeff), diffuse=tuple(m.diffuse), specular=tuple(m.specular),
                            emission=tuple(m.emission), reflection=m.reflection, roughness=m.roughness, ns=m.ns))
        return xyz9, uv6, mat, out

    def triangle_intersect(self, xyz9, o, d):
        t, u, v = C.c_float(), C.c_float(), C.c_float()
        r = self._fn("triangle_intersect")(_ptr(_f32(xyz9), C.c_float), _ptr(_f32(o), C.c_float), _ptr(_f32(d), C.c_float),
                                           C.byref(t), C.byref(u), C.byref(v))
        return bool(r), t.value, u.value, v.value

    def camera_matrices(self, fov, aspect, znear=0.1, zfar=1000.0):
        p = np.zeros(16, np.float32)
        pi = np.zeros(16, np.float32)
        self._fn("camera_matrices")(fov, aspect, znear, zfar, _ptr(p, C.c_float), _ptr(pi, C.c_float))
        return p.reshape(4, 4), pi.reshape(4, 4)

    def transform_inverse(self, m):
        m = _f32(m).reshape(16)
        out = np.zeros(16, np.float32)
        self._fn("transform_inverse")(_ptr(m, C.c_float), _ptr(out, C.c_float))
        return out.reshape(4, 4)

    def downscale(self, argb, factor):
        argb = np.ascontiguousarray(argb, dtype=np.uint32)
        h, w = argb.shape
        out = np.zeros((h // factor, w // factor), np.uint32)
        self._fn("downscale")(_ptr(argb, C.c_uint32), w, h, factor, _ptr(out, C.c_uint32))
        return out

    # ---- BVH ------------------------------------------------------------------------------------------
    def bvh(self, xyz9, max_depth=10, leaf_max=8):
        return CpuBvh(self, xyz9, max_depth, leaf_max)

    def renderer(self):
        return CpuRenderer(self)


class CpuBvh:
    def __init__(self, tracer: CpuTracer, xyz9, max_depth, leaf_max):
        self.tr = tracer
        self.xyz9 = _f32(xyz9).reshape(-1, 9)
        ms = C.c_double(0)
        self.h = tracer._fn("bvh_create")(_ptr(self.xyz9, C.c_float), len(self.xyz9), max_depth, leaf_max, C.byref(ms))
        self.build_ms = ms.value

    def close(self):
        if self.h:
            self.tr._fn("bvh_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        out = (C.c_uint64 * 6)()
        self.tr._fn("bvh_stats")(self.h, out)
        keys = ("nodes", "leaves", "empty_leaves", "interior", "max_depth_reached", "max_leaf_size")
        return dict(zip(keys, [int(x) for x in out]))

    def intersect(self, o3, d3, threads=0):
        o3 = _f32(o3).reshape(-1, 3)
        d3 = _f32(d3).reshape(-1, 3)
        n = len(o3)
        tri = np.empty(n, np.int32)
        t = np.empty(n, np.float32)
        u = np.empty(n, np.float32)
        v = np.empty(n, np.float32)
        ms = self.tr._fn("bvh_intersect")(self.h, _ptr(o3, C.c_float), _ptr(d3, C.c_float), n, _ptr(tri, C.c_int32),
                                          _ptr(t, C.c_float), _ptr(u, C.c_float), _ptr(v, C.c_float), threads)
        self.last_ms = ms
        return tri, t, u, v

    def count(self, o3, d3, threads=0):
        """Oracle only: total (volume tests, triangle tests) of the reference traversal for these rays."""
        o3 = _f32(o3).reshape(-1, 3)
        d3 = _f32(d3).reshape(-1, 3)
        out = (C.c_uint64 * 2)()
        self.tr._fn("bvh_count")(self.h, _ptr(o3, C.c_float), _ptr(d3, C.c_float), len(o3), out, threads)
        return int(out[0]), int(out[1])


class CpuRenderer:
    def __init__(self, tracer: CpuTracer):
        self.tr = tracer
        self.h = tracer._fn("renderer_create")()
        self.settings = None
        self.cam_to_world = np.eye(4, dtype=np.float32)

    def close(self):
        if self.h:
            self.tr._fn("renderer_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, settings: RtSettings, fov: float):
        self.settings = settings
        self.tr._fn("renderer_configure")(self.h, C.byref(settings), fov)

    def set_triangles(self, xyz9, uv6=None, mat=None):
        xyz9 = _f32(xyz9).reshape(-1, 9)
        uv6 = _f32(uv6).reshape(-1, 6) if uv6 is not None else None
        mat = np.ascontiguousarray(mat, dtype=np.int32) if mat is not None else None
        self.tr._fn("renderer_set_triangles")(self.h, _ptr(xyz9, C.c_float), _ptr(uv6, C.c_float), _ptr(mat, C.c_int32), len(xyz9))

    def set_materials(self, mats):
        arr = materials_array(mats)
        self.tr._fn("renderer_set_materials")(self.h, arr, len(mats))

    def set_texture(self, slot, rgba):
        rgba = np.ascontiguousarray(rgba)
        h, w = rgba.shape[:2]
        if rgba.dtype == np.uint8:
            self.tr._fn("renderer_set_texture_u8")(self.h, slot, _ptr(rgba, C.c_uint8), w, h)
        else:
            rgba = _f32(rgba)
            self.tr._fn("renderer_set_texture_f32")(self.h, slot, _ptr(rgba, C.c_float), w, h)

    def set_camera_transform(self, m):
        self.cam_to_world = _f32(m).reshape(4, 4).copy()
        self.tr._fn("renderer_set_camera_transform")(self.h, _ptr(self.cam_to_world, C.c_float))

    def add_sphere(self, center, radius, mat_index):
        c = _f32(center)
        self.tr._fn("renderer_add_sphere")(self.h, _ptr(c, C.c_float), float(radius), int(mat_index))

    def add_plane(self, point, normal, mat_index):
        p, n = _f32(point), _f32(normal)
        self.tr._fn("renderer_add_plane")(self.h, _ptr(p, C.c_float), _ptr(n, C.c_float), int(mat_index))

    def set_light(self, p):
        p = _f32(p)
        self.tr._fn("renderer_set_light")(self.h, _ptr(p, C.c_float))

    def _super_dims(self):
        s = self.settings
        f = s.ssaa_factor if s.enable_ssaa else 1
        return s.image_width * f, s.image_height * f

    def render(self, threads=0):
        """ray_trace() + post_process(); returns (argb[H,W] bottom-up, ms)."""
        s = self.settings
        out = np.zeros((s.image_height, s.image_width), np.uint32)
        ms = self.tr._fn("renderer_render")(self.h, _ptr(out, C.c_uint32), threads)
        return out, ms

    def render_ssao(self, srand_seed=None, ref_seeds9=None, n_rand=64, threads=0):
        """ray_trace() with the G-buffers + post_process() with enable_ssao (renderer.cpp:1229-1434).
        Compiled reference: the SSAO pass runs on one thread after srand(srand_seed); returns (argb, rand_values) where
        rand_values are the first n_rand values std::rand() gives after that srand.
        Oracle: ref_seeds9 = None -> the per-pixel stream (settings.rng_seed); else the 9 generator seeds of a
        single-threaded reference run (8 lanes + scalar), reference order; returns (argb, None)."""
        s = self.settings
        out = np.zeros((s.image_height, s.image_width), np.uint32)
        if self.tr.kind != "oracle":
            rv = np.zeros(n_rand, np.uint32)
            self.tr._fn("renderer_render_ssao")(self.h, _ptr(out, C.c_uint32), int(srand_seed), _ptr(rv, C.c_uint32), n_rand)
            return out, rv
        seeds = None if ref_seeds9 is None else np.ascontiguousarray(ref_seeds9, dtype=np.uint32)
        self.tr._fn("renderer_render_ssao")(self.h, _ptr(out, C.c_uint32), threads, 0 if seeds is None else 1, _ptr(seeds, C.c_uint32))
        return out, None

    def raster(self, srand_seed=1, ref_seeds9=None, n_rand=64):
        """raster_trace() + post_process() with hybrid_rasterization_tracing (renderer.cpp:869-1006) in sequential triangle
        order (the compiled reference runs it on one thread).  Returns (argb, extra): extra = rand_values for the compiled
        reference (see render_ssao), the dict of ray counters for the oracle."""
        s = self.settings
        out = np.zeros((s.image_height, s.image_width), np.uint32)
        if self.tr.kind != "oracle":
            rv = np.zeros(n_rand, np.uint32)
            self.tr._fn("renderer_raster")(self.h, _ptr(out, C.c_uint32), int(srand_seed), _ptr(rv, C.c_uint32), n_rand)
            return out, rv
        seeds = None if ref_seeds9 is None else np.ascontiguousarray(ref_seeds9, dtype=np.uint32)
        cnt = (C.c_uint64 * 5)()
        self.tr._fn("renderer_raster")(self.h, _ptr(out, C.c_uint32), 0 if seeds is None else 1, _ptr(seeds, C.c_uint32), cnt)
        return out, dict(zip(("fragments", "fragment_hits", "shadow_rays", "reflection_rays", "reflection_shadow_rays"), [int(x) for x in cnt]))

    def trace_rows(self, row_begin=0, row_end=None, row_step=1, reseed=True, threads=0, want_image=True):
        """Seeded pixel loop on the supersampled frame; returns (argb_super[H',W'], ms)."""
        rw, rh = self._super_dims()
        if row_end is None:
            row_end = rh
        out = np.zeros((rh, rw), np.uint32) if want_image else None
        ms = self.tr._fn("renderer_trace_rows")(self.h, _ptr(self.cam_to_world, C.c_float), _ptr(out, C.c_uint32),
                                                row_begin, row_end, row_step, int(reseed), threads)
        return out, ms

    def last_hit_count(self) -> int:
        """Compiled reference only: primary hits of the last trace_rows() call."""
        return int(self.tr._fn("last_hit_count")())

    def count_rows(self, row_begin=0, row_end=None, row_step=1, threads=0):
        """Oracle only: ray / test counters for the same pixel loop (see oracle.h OrcCounters)."""
        rw, rh = self._super_dims()
        if row_end is None:
            row_end = rh
        out = (C.c_uint64 * 16)()
        self.tr._fn("renderer_count_rows")(self.h, _ptr(self.cam_to_world, C.c_float), row_begin, row_end, row_step, out, threads)
        keys = ("primary_rays", "shadow_rays", "reflection_rays", "reflection_shadow_rays", "primary_hits",
                "primary_volume_tests", "primary_triangle_tests", "shadow_volume_tests", "shadow_triangle_tests",
                "reflection_volume_tests", "reflection_triangle_tests")
        return dict(zip(keys, [int(x) for x in out]))
