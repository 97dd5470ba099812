"""go-raytracing_b200 — thin ctypes glue over the product's two native libraries.

  csrc/librtx_b200.so   hand-written sm_100a CUDA wavefront path tracer behind the C-ABI of include/rtx_b200.h
  csrc/librt_host.so    C++ host mirror of the reference's Go `rt` package (scene types, camera builder, the five
                        configured scenes, flattener, BucketRenderer) — see host/rt.hpp

Python is only plumbing here (tests, bench, torch.distributed for the multi-GPU reduce). All arithmetic of the
hot path runs in the CUDA library; there is no CPU fallback, and a missing extension is a hard error.
The package directory name contains a hyphen (it mirrors the reference's name); import it with
`importlib.import_module("go-raytracing_b200")`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("RTX_B200_LIB") or os.path.join(_HERE, "csrc", "librtx_b200.so")  # env override: A/B builds of the same library
HOST_LIB_PATH = os.path.join(_HERE, "csrc", "librt_host.so")

RTX_ABI_VERSION = 2
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT, MAT_ISOTROPIC = range(5)
TEX_SOLID, TEX_CHECKER, TEX_NOISE, TEX_IMAGE = 0, 1, 2, 3
GEOM_SPHERE, GEOM_QUAD, GEOM_TRIANGLE, GEOM_PLANE, GEOM_LIST, GEOM_MESH, GEOM_CIRCLE = range(7)
XF_TRANSLATE, XF_ROTATE_Y, XF_SCALE = 0, 1, 2

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class SceneDesc(C.Structure):
    """rtx_scene_desc (include/rtx_b200.h)."""
    _fields_ = [
        ("abi_version", C.c_uint32), ("world_is_bvh", C.c_int32),
        ("n_textures", C.c_int32), ("tex_type", _pi), ("tex_color", _pd), ("tex_inv_scale", _pd), ("tex_even", _pi), ("tex_odd", _pi),
        ("n_materials", C.c_int32), ("mat_type", _pi), ("mat_tex", _pi), ("mat_albedo", _pd), ("mat_fuzz", _pd), ("mat_ior", _pd),
        ("n_spheres", C.c_int32), ("sph_center", _pd), ("sph_velocity", _pd), ("sph_radius", _pd), ("sph_mat", _pi),
        ("n_quads", C.c_int32), ("quad_q", _pd), ("quad_u", _pd), ("quad_v", _pd), ("quad_mat", _pi),
        ("n_tris", C.c_int32), ("tri_v0", _pd), ("tri_v1", _pd), ("tri_v2", _pd), ("tri_mat", _pi), ("tri_rank", _pi),
        ("n_planes", C.c_int32), ("plane_point", _pd), ("plane_normal", _pd), ("plane_mat", _pi),
        ("n_groups", C.c_int32), ("group_kind", _pi), ("group_begin", _pi), ("group_count", _pi),
        ("n_list_items", C.c_int32), ("list_item_kind", _pi), ("list_item_index", _pi),
        ("n_xforms", C.c_int32), ("xf_type", _pi), ("xf_a", _pd), ("xf_b", _pd),
        ("n_volumes", C.c_int32), ("vol_neg_inv_density", _pd), ("vol_mat", _pi),
        ("n_entries", C.c_int32), ("entry_geom_kind", _pi), ("entry_geom_index", _pi), ("entry_xf_begin", _pi), ("entry_xf_count", _pi),
        ("entry_volume", _pi), ("entry_rank", _pi),
        ("n_lights", C.c_int32), ("light_quad", _pi),
        ("env_width", C.c_int32), ("env_height", C.c_int32), ("env_rgb", _pd), ("env_rotation", C.c_double),
        ("env_importance_sampling", C.c_int32),
        ("n_circles", C.c_int32), ("circle_center", _pd), ("circle_normal", _pd), ("circle_radius", _pd), ("circle_mat", _pi),
        ("n_perlin", C.c_int32), ("perlin_vec", _pd), ("perlin_perm", _pi),
        ("n_images", C.c_int32), ("image_width", _pi), ("image_height", _pi), ("image_offset", C.POINTER(C.c_int64)), ("image_rgb", _pd),
    ]


class CameraDesc(C.Structure):
    """rtx_camera_desc (include/rtx_b200.h)."""
    _fields_ = [
        ("aspect_ratio", C.c_double), ("image_width", C.c_int32), ("samples_per_pixel", C.c_int32), ("max_depth", C.c_int32),
        ("vfov", C.c_double), ("look_from", C.c_double * 3), ("look_at", C.c_double * 3), ("vup", C.c_double * 3),
        ("defocus_angle", C.c_double), ("focus_dist", C.c_double), ("look_from2", C.c_double * 3), ("look_at2", C.c_double * 3),
        ("camera_motion", C.c_int32), ("free_camera", C.c_int32), ("forward", C.c_double * 3), ("background", C.c_double * 3),
        ("use_sky_gradient", C.c_int32), ("phantom_hdri", C.c_int32),
        ("has_derived", C.c_int32), ("image_height", C.c_int32),
        ("center", C.c_double * 3), ("pixel00_loc", C.c_double * 3), ("pixel_delta_u", C.c_double * 3), ("pixel_delta_v", C.c_double * 3),
        ("u", C.c_double * 3), ("v", C.c_double * 3), ("w", C.c_double * 3),
        ("defocus_radius", C.c_double), ("viewport_width", C.c_double), ("viewport_height", C.c_double),
    ]


class Stats(C.Structure):
    """rtx_stats (include/rtx_b200.h)."""
    _fields_ = [
        ("paths", C.c_uint64), ("extension_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("nodes_visited", C.c_uint64),
        ("tri_tests", C.c_uint64), ("sphere_tests", C.c_uint64), ("quad_tests", C.c_uint64), ("plane_tests", C.c_uint64),
        ("wavefront_iterations", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("ms_generate", C.c_double), ("ms_extend", C.c_double), ("ms_shade", C.c_double), ("ms_connect", C.c_double), ("ms_total", C.c_double),
        ("tlas_nodes", C.c_uint32), ("blas_nodes", C.c_uint32), ("n_entries", C.c_uint32), ("n_tris", C.c_uint32),
        ("blas_depth", C.c_uint32), ("bvh_on_device", C.c_uint32), ("ms_bvh_build", C.c_double), ("ms_scene_upload", C.c_double),
        ("ms_tail", C.c_double), ("ms_reduce", C.c_double), ("ms_resolve", C.c_double), ("tail_iterations", C.c_uint64),
        ("n_devices", C.c_uint32), ("checked_build", C.c_uint32), ("checked_violations", C.c_uint64), ("checked_by_kind", C.c_uint64 * 8),
    ]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "checked_by_kind" else getattr(self, k)) for k, _ in self._fields_}


# every symbol include/rtx_b200.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "rtx_create", "rtx_destroy", "rtx_last_error", "rtx_abi_version", "rtx_scene_upload", "rtx_camera_set", "rtx_image_size",
    "rtx_render_pass", "rtx_accum_clear", "rtx_accum_enable_moments", "rtx_accum_device_ptr", "rtx_resolve_rgba8", "rtx_resolve_accum",
    "rtx_trace_closest", "rtx_camera_rays", "rtx_hdri_sample", "rtx_hdri_pdf", "rtx_hdri_lookup", "rtx_hdri_total_power",
    "rtx_get_stats", "rtx_set_option", "rtx_set_stream", "rtx_create_multi", "rtx_device_count", "rtx_mesh_test_order",
]

_lib_cache = None
_host_cache = None


class ExtensionMissing(RuntimeError):
    pass


def lib() -> C.CDLL:
    """The CUDA library. Raises loudly when it has not been built (there is no fallback path)."""
    global _lib_cache
    if _lib_cache is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(make -C go-raytracing_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.rtx_last_error.restype = C.c_char_p
        L.rtx_last_error.argtypes = [C.c_void_p]
        L.rtx_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
        L.rtx_create_multi.argtypes = [_pi, C.c_int32, C.POINTER(C.c_void_p)]
        L.rtx_destroy.argtypes = [C.c_void_p]
        L.rtx_scene_upload.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        L.rtx_camera_set.argtypes = [C.c_void_p, C.POINTER(CameraDesc)]
        L.rtx_image_size.argtypes = [C.c_void_p, _pi, _pi]
        L.rtx_render_pass.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32]
        L.rtx_accum_clear.argtypes = [C.c_void_p]
        L.rtx_accum_enable_moments.argtypes = [C.c_void_p, C.c_int32]
        L.rtx_accum_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.rtx_resolve_rgba8.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
        L.rtx_resolve_accum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rtx_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double] + [C.c_void_p] * 7
        L.rtx_camera_rays.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
        L.rtx_hdri_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rtx_hdri_pdf.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.rtx_hdri_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.rtx_hdri_total_power.argtypes = [C.c_void_p, _pd]
        L.rtx_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.rtx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.rtx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        _lib_cache = L
    return _lib_cache


def host() -> C.CDLL:
    """The C++ host mirror of the Go rt package."""
    global _host_cache
    if _host_cache is None:
        lib()
        if not os.path.exists(HOST_LIB_PATH):
            raise ExtensionMissing(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        H = C.CDLL(HOST_LIB_PATH)
        H.rth_last_error.restype = C.c_char_p
        H.rth_scene_named.restype = C.c_void_p
        H.rth_scene_named.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32]
        H.rth_scene_free.argtypes = [C.c_void_p]
        H.rth_scene_desc.restype = C.POINTER(SceneDesc)
        H.rth_scene_desc.argtypes = [C.c_void_p]
        H.rth_camera_desc.restype = C.POINTER(CameraDesc)
        H.rth_camera_desc.argtypes = [C.c_void_p]
        H.rth_image_height.argtypes = [C.c_void_p]
        H.rth_image_height_for.argtypes = [C.c_int32, C.c_double]
        H.rth_bucket_render.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64, _pd, C.c_char_p]
        H.rth_bucket_render_progressive.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        H.rth_load_hdr.argtypes = [C.c_char_p, _pi, _pi, C.c_void_p, C.c_int64]
        H.rth_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32]
        H.rth_stats_bar.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_char_p]
        H.rth_parse_obj.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _pd]
        _host_cache = H
    return _host_cache


# ------------------------------------------------------------------------------------------------------------
# scene descriptions
# ------------------------------------------------------------------------------------------------------------
def _d(a, cols=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def _i(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))


class SceneBuilder:
    """Builds an rtx_scene_desc from plain Python/numpy data (tests; arbitrary small scenes).

    Mirrors what the Go flattener emits: primitives are added to typed arrays, `entry()` appends a world entry in
    insertion order. Materials: ('lambertian', rgb | tex_id) / ('metal', rgb, fuzz) / ('dielectric', ior) /
    ('light', rgb) / ('isotropic', rgb)."""

    def __init__(self, world_is_bvh: bool = True):
        self.world_is_bvh = world_is_bvh
        self.tex = []      # (type, color, inv_scale, even, odd)
        self.mat = []      # (type, tex, albedo, fuzz, ior)
        self.sph, self.quad, self.tri, self.plane, self.circ = [], [], [], [], []
        self.perlin = []   # (vec[256][3], perm[3][256])
        self.images = []   # float64 [H][W][3], ImageLoader.data (after the reference's sqrt on load)
        self.groups, self.items = [], []
        self.xf = []
        self.vol = []
        self.entries = []
        self.lights = []
        self.env = None

    def solid(self, rgb):
        self.tex.append((TEX_SOLID, tuple(rgb), 0.0, -1, -1))
        return len(self.tex) - 1

    def checker(self, scale, c1, c2):
        e, o = self.solid(c1), self.solid(c2)
        self.tex.append((TEX_CHECKER, (0, 0, 0), 1.0 / scale, e, o))
        return len(self.tex) - 1

    def noise(self, scale, seed=1):
        """NoiseTexture (rt/texture.go:24-29, :81-85) with a Perlin table drawn here from a seeded generator in the reference's
        draw order (rt/noise.go:15-28: 256 unit vectors from [-1,1)^3, then Fisher-Yates permutations X, Y, Z)."""
        rng = np.random.default_rng(seed)
        vec = rng.random((256, 3)) * 2.0 - 1.0
        vec /= np.sqrt((vec * vec).sum(axis=1, keepdims=True))
        perm = np.zeros((3, 256), dtype=np.int32)
        for a in range(3):
            p = np.arange(256, dtype=np.int32)
            for i in range(255, 0, -1):
                t = int(rng.integers(0, i + 1))
                p[i], p[t] = p[t], p[i]
            perm[a] = p
        self.perlin.append((vec, perm))
        self.tex.append((TEX_NOISE, (0, 0, 0), float(scale), len(self.perlin) - 1, -1))
        return len(self.tex) - 1

    def image(self, rgb8):
        """ImageTexture over an 8-bit RGB image [H][W][3]: ImageLoader.Load converts to sqrt(v / 255) (rt/image_loader.go:62-70)."""
        a = np.asarray(rgb8)
        assert a.ndim == 3 and a.shape[2] == 3
        self.images.append(np.ascontiguousarray(np.sqrt(a.astype(np.float64) / 255.0)))
        self.tex.append((TEX_IMAGE, (0, 0, 0), 0.0, len(self.images) - 1, -1))
        return len(self.tex) - 1

    def material(self, kind, *args):
        if kind == "lambertian":
            tex = args[0] if isinstance(args[0], int) else self.solid(args[0])
            self.mat.append((MAT_LAMBERTIAN, tex, (0, 0, 0), 0.0, 0.0))
        elif kind == "metal":
            self.mat.append((MAT_METAL, -1, tuple(args[0]), min(float(args[1]), 1.0), 0.0))
        elif kind == "dielectric":
            self.mat.append((MAT_DIELECTRIC, -1, (0, 0, 0), 0.0, float(args[0])))
        elif kind == "light":
            self.mat.append((MAT_DIFFUSE_LIGHT, args[0] if isinstance(args[0], int) else self.solid(args[0]), (0, 0, 0), 0.0, 0.0))
        elif kind == "isotropic":
            self.mat.append((MAT_ISOTROPIC, self.solid(args[0]), (0, 0, 0), 0.0, 0.0))
        else:
            raise ValueError(kind)
        return len(self.mat) - 1

    def sphere(self, center, radius, mat, center2=None):
        vel = (0, 0, 0) if center2 is None else tuple(np.subtract(center2, center))
        self.sph.append((tuple(center), vel, float(radius), mat))
        return len(self.sph) - 1

    def quadp(self, Q, u, v, mat):
        self.quad.append((tuple(Q), tuple(u), tuple(v), mat))
        return len(self.quad) - 1

    def triangle(self, v0, v1, v2, mat):
        self.tri.append((tuple(v0), tuple(v1), tuple(v2), mat))
        return len(self.tri) - 1

    def planep(self, point, normal, mat):
        n = np.asarray(normal, dtype=np.float64)
        n = n / np.sqrt((n * n).sum())
        self.plane.append((tuple(point), tuple(n), mat))
        return len(self.plane) - 1

    def circle(self, center, normal, radius, mat):
        n = np.asarray(normal, dtype=np.float64)
        n = n / np.sqrt((n * n).sum())
        self.circ.append((tuple(center), tuple(n), float(radius), mat))
        return len(self.circ) - 1

    def pyramid_group(self, base_center, base_size, height, mat):
        """rt.Pyramid (rt/primitives.go:39-71): a base quad and four triangles in one HittableList."""
        cx, cy, cz = (float(v) for v in base_center)
        h = base_size / 2
        items = [(GEOM_QUAD, self.quadp((cx - base_size / 2, cy, cz - base_size / 2), (base_size, 0, 0), (0, 0, base_size), mat))]
        apex = (cx, cy + height, cz)
        corners = [(cx + h, cy, cz - h), (cx + h, cy, cz + h), (cx - h, cy, cz + h), (cx - h, cy, cz - h)]
        for i in range(4):
            items.append((GEOM_TRIANGLE, self.triangle(corners[i], corners[(i + 1) % 4], apex, mat)))
        return self.list_group(items)

    def list_group(self, items: Sequence[tuple]):
        begin = len(self.items)
        self.items.extend(items)
        self.groups.append((GEOM_LIST, begin, len(items)))
        return len(self.groups) - 1

    def box_group(self, a, b, mat):
        """rt.Box (rt/primitives.go:5-37): six quads front,right,back,left,top,bottom."""
        mn, mx = np.minimum(a, b).astype(float), np.maximum(a, b).astype(float)
        dx, dy, dz = np.array([mx[0] - mn[0], 0, 0]), np.array([0, mx[1] - mn[1], 0]), np.array([0, 0, mx[2] - mn[2]])
        qs = [((mn[0], mn[1], mx[2]), dx, dy), ((mx[0], mn[1], mx[2]), -dz, dy), ((mx[0], mn[1], mn[2]), -dx, dy),
              ((mn[0], mn[1], mn[2]), dz, dy), ((mn[0], mx[1], mx[2]), dx, -dz), ((mn[0], mn[1], mn[2]), dx, dz)]
        return self.list_group([(GEOM_QUAD, self.quadp(Q, u, v, mat)) for Q, u, v in qs])

    def mesh_group(self, v0, v1, v2, mat):
        v0, v1, v2 = _d(v0, 3), _d(v1, 3), _d(v2, 3)
        begin = len(self.tri)
        for a, b, c in zip(v0, v1, v2):
            self.tri.append((tuple(a), tuple(b), tuple(c), mat))
        self.groups.append((GEOM_MESH, begin, len(v0)))
        return len(self.groups) - 1

    def entry(self, kind, index, xforms=(), volume=None):
        """xforms: outermost first: ('translate', off) / ('rotate_y', degrees) / ('scale', factor3)."""
        xb = len(self.xf)
        for x in xforms:
            if x[0] == "translate":
                self.xf.append((XF_TRANSLATE, tuple(x[1]), (0, 0, 0)))
            elif x[0] == "rotate_y":
                rad = x[1] * 3.1415926535897932385 / 180.0
                self.xf.append((XF_ROTATE_Y, (float(np.sin(rad)), float(np.cos(rad)), 0.0), (0, 0, 0)))
            elif x[0] == "scale":
                f = np.asarray(x[1], dtype=np.float64) * np.ones(3)
                self.xf.append((XF_SCALE, tuple(f), tuple(1.0 / f)))
            else:
                raise ValueError(x)
        vol = -1
        if volume is not None:
            density, mat = volume
            self.vol.append((-1.0 / density, mat))
            vol = len(self.vol) - 1
        self.entries.append((kind, index, xb, len(xforms), vol))
        return len(self.entries) - 1

    def light(self, quad_index):
        self.lights.append(quad_index)

    def environment(self, rgb, rotation_rad=0.0, importance_sampling=True):
        rgb = np.ascontiguousarray(rgb, dtype=np.float64)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        self.env = (rgb, float(rotation_rad), bool(importance_sampling))

    def build(self) -> "BuiltScene":
        return BuiltScene(self)


class BuiltScene:
    """Owns the numpy arrays an rtx_scene_desc points to."""

    def __init__(self, b: SceneBuilder):
        k = self._keep = {}

        def P(name, arr, dtype):
            arr = np.ascontiguousarray(np.asarray(arr, dtype=dtype).reshape(-1))
            if arr.size == 0:
                arr = np.zeros(1, dtype=dtype)
            k[name] = arr
            return arr.ctypes.data_as(_pd if dtype == np.float64 else _pi)

        d = SceneDesc()
        d.abi_version = RTX_ABI_VERSION
        d.world_is_bvh = int(b.world_is_bvh)
        d.n_textures = len(b.tex)
        d.tex_type = P("tex_type", [t[0] for t in b.tex], np.int32)
        d.tex_color = P("tex_color", [t[1] for t in b.tex], np.float64)
        d.tex_inv_scale = P("tex_inv_scale", [t[2] for t in b.tex], np.float64)
        d.tex_even = P("tex_even", [t[3] for t in b.tex], np.int32)
        d.tex_odd = P("tex_odd", [t[4] for t in b.tex], np.int32)
        d.n_materials = len(b.mat)
        d.mat_type = P("mat_type", [m[0] for m in b.mat], np.int32)
        d.mat_tex = P("mat_tex", [m[1] for m in b.mat], np.int32)
        d.mat_albedo = P("mat_albedo", [m[2] for m in b.mat], np.float64)
        d.mat_fuzz = P("mat_fuzz", [m[3] for m in b.mat], np.float64)
        d.mat_ior = P("mat_ior", [m[4] for m in b.mat], np.float64)
        d.n_spheres = len(b.sph)
        d.sph_center = P("sph_center", [s[0] for s in b.sph], np.float64)
        d.sph_velocity = P("sph_velocity", [s[1] for s in b.sph], np.float64)
        d.sph_radius = P("sph_radius", [s[2] for s in b.sph], np.float64)
        d.sph_mat = P("sph_mat", [s[3] for s in b.sph], np.int32)
        d.n_quads = len(b.quad)
        d.quad_q = P("quad_q", [q[0] for q in b.quad], np.float64)
        d.quad_u = P("quad_u", [q[1] for q in b.quad], np.float64)
        d.quad_v = P("quad_v", [q[2] for q in b.quad], np.float64)
        d.quad_mat = P("quad_mat", [q[3] for q in b.quad], np.int32)
        d.n_tris = len(b.tri)
        d.tri_v0 = P("tri_v0", [t[0] for t in b.tri], np.float64)
        d.tri_v1 = P("tri_v1", [t[1] for t in b.tri], np.float64)
        d.tri_v2 = P("tri_v2", [t[2] for t in b.tri], np.float64)
        d.tri_mat = P("tri_mat", [t[3] for t in b.tri], np.int32)
        d.tri_rank = None
        d.n_planes = len(b.plane)
        d.plane_point = P("plane_point", [p[0] for p in b.plane], np.float64)
        d.plane_normal = P("plane_normal", [p[1] for p in b.plane], np.float64)
        d.plane_mat = P("plane_mat", [p[2] for p in b.plane], np.int32)
        d.n_groups = len(b.groups)
        d.group_kind = P("group_kind", [g[0] for g in b.groups], np.int32)
        d.group_begin = P("group_begin", [g[1] for g in b.groups], np.int32)
        d.group_count = P("group_count", [g[2] for g in b.groups], np.int32)
        d.n_list_items = len(b.items)
        d.list_item_kind = P("list_item_kind", [i[0] for i in b.items], np.int32)
        d.list_item_index = P("list_item_index", [i[1] for i in b.items], np.int32)
        d.n_xforms = len(b.xf)
        d.xf_type = P("xf_type", [x[0] for x in b.xf], np.int32)
        d.xf_a = P("xf_a", [x[1] for x in b.xf], np.float64)
        d.xf_b = P("xf_b", [x[2] for x in b.xf], np.float64)
        d.n_volumes = len(b.vol)
        d.vol_neg_inv_density = P("vol_nid", [v[0] for v in b.vol], np.float64)
        d.vol_mat = P("vol_mat", [v[1] for v in b.vol], np.int32)
        d.n_entries = len(b.entries)
        d.entry_geom_kind = P("e_kind", [e[0] for e in b.entries], np.int32)
        d.entry_geom_index = P("e_index", [e[1] for e in b.entries], np.int32)
        d.entry_xf_begin = P("e_xb", [e[2] for e in b.entries], np.int32)
        d.entry_xf_count = P("e_xc", [e[3] for e in b.entries], np.int32)
        d.entry_volume = P("e_vol", [e[4] for e in b.entries], np.int32)
        d.entry_rank = None
        d.n_lights = len(b.lights)
        d.light_quad = P("lights", b.lights, np.int32)
        d.n_circles = len(b.circ)
        d.circle_center = P("circ_c", [c[0] for c in b.circ], np.float64)
        d.circle_normal = P("circ_n", [c[1] for c in b.circ], np.float64)
        d.circle_radius = P("circ_r", [c[2] for c in b.circ], np.float64)
        d.circle_mat = P("circ_m", [c[3] for c in b.circ], np.int32)
        d.n_perlin = len(b.perlin)
        d.perlin_vec = P("perlin_vec", [pv[0] for pv in b.perlin], np.float64)
        d.perlin_perm = P("perlin_perm", [pv[1] for pv in b.perlin], np.int32)
        d.n_images = len(b.images)
        d.image_width = P("img_w", [im.shape[1] for im in b.images], np.int32)
        d.image_height = P("img_h", [im.shape[0] for im in b.images], np.int32)
        offs = np.cumsum([0] + [im.shape[0] * im.shape[1] for im in b.images])[:-1] if b.images else np.zeros(0)
        k["img_off"] = np.ascontiguousarray(np.asarray(offs if len(offs) else [0], dtype=np.int64))
        d.image_offset = k["img_off"].ctypes.data_as(C.POINTER(C.c_int64))
        d.image_rgb = P("img_rgb", np.concatenate([im.reshape(-1) for im in b.images]) if b.images else [], np.float64)
        if b.env is not None:
            rgb, rot, is_ = b.env
            k["env"] = rgb
            d.env_height, d.env_width = rgb.shape[0], rgb.shape[1]
            d.env_rgb = rgb.ctypes.data_as(_pd)
            d.env_rotation = rot
            d.env_importance_sampling = int(is_)
        self.desc = d

    @property
    def desc_ptr(self):
        return C.pointer(self.desc)


def make_camera(width, aspect, spp, depth, vfov, look_from, look_at, vup=(0, 1, 0), defocus_angle=0.0, focus_dist=10.0,
                background=(0, 0, 0), sky=False, phantom=False, motion=None, free_forward=None) -> CameraDesc:
    """CameraBuilder equivalent (rt/camera.go:175-280) producing an rtx_camera_desc."""
    c = CameraDesc()
    c.aspect_ratio, c.image_width, c.samples_per_pixel, c.max_depth, c.vfov = aspect, width, spp, depth, vfov
    c.look_from[:] = look_from
    c.look_at[:] = look_at
    c.vup[:] = vup
    c.defocus_angle, c.focus_dist = defocus_angle, focus_dist
    c.background[:] = background
    c.use_sky_gradient, c.phantom_hdri = int(sky), int(phantom)
    c.forward[:] = (0, 0, -1)
    if motion is not None:
        c.look_from2[:], c.look_at2[:] = motion
        c.camera_motion = 1
    if free_forward is not None:
        f = np.asarray(free_forward, dtype=np.float64)
        c.forward[:] = f / np.sqrt((f * f).sum())
        c.free_camera = 1
    return c


class NamedScene:
    """A scene function of rt/scenes.go built by the C++ host mirror (rt_scenes.cpp) and flattened."""

    def __init__(self, name: str, width: int = 0, aspect: float = 16.0 / 9.0, spp: int = 0, depth: int = 0, seed: int = 0x5EED,
                 use_bvh: bool = True, asset_root: Optional[str] = None):
        H = host()
        root = (asset_root or REPO_ROOT).encode()
        self._h = H.rth_scene_named(name.encode(), root, seed, int(use_bvh), width, aspect, spp, depth)
        if not self._h:
            raise RuntimeError(f"scene '{name}': {H.rth_last_error().decode()}")
        self.name = name
        self.desc_ptr = H.rth_scene_desc(self._h)
        self.cam_ptr = H.rth_camera_desc(self._h)
        self.desc = self.desc_ptr.contents
        self.cam = self.cam_ptr.contents
        self.width = self.cam.image_width
        self.height = H.rth_image_height(self._h)

    def bucket_render(self, seed=1, save_path: str = ""):
        """Runs the BucketRenderer mirror (3 passes) and returns (RGBA8 image, seconds)."""
        pix = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        sec = C.c_double(0)
        rc = host().rth_bucket_render(self._h, seed, pix.ctypes.data, pix.nbytes, C.byref(sec), save_path.encode())
        if rc != 0:
            raise RuntimeError(host().rth_last_error().decode())
        return pix, sec.value

    def bucket_render_progressive(self, seed=1):
        """The display-loop use: Update() is ticked without ever blocking on a pass and the framebuffer is copied every tick.
        Returns (final RGBA8 image, ticks that returned while a pass was running, distinct finished passes seen on the way)."""
        pix = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        ticks, frames = C.c_int32(0), C.c_int32(0)
        rc = host().rth_bucket_render_progressive(self._h, seed, pix.ctypes.data, pix.nbytes, C.byref(ticks), C.byref(frames))
        if rc != 0:
            raise RuntimeError(host().rth_last_error().decode())
        return pix, ticks.value, frames.value

    def close(self):
        if self._h:
            host().rth_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------------------
# device context
# ------------------------------------------------------------------------------------------------------------
class RtxError(RuntimeError):
    pass


def device_count() -> int:
    return int(lib().rtx_device_count())


class Context:
    """One context of the CUDA library (rtx_ctx): one GPU, or — `devices=[...]` — several GPUs of the box behind one context
    (rtx_create_multi: the library slices the samples of a pass over the devices and sums the buffers with one ncclReduce)."""

    def __init__(self, device: int = 0, devices: Optional[Sequence[int]] = None):
        self._L = lib()
        h = C.c_void_p()
        if devices is not None:
            ids = (C.c_int32 * len(devices))(*devices)
            rc = self._L.rtx_create_multi(ids, len(devices), C.byref(h))
            what, device = f"rtx_create_multi({list(devices)})", devices[0]
        else:
            rc = self._L.rtx_create(device, C.byref(h))
            what = f"rtx_create({device})"
        if rc != 0:
            raise RtxError(f"{what} = {rc}: {self._L.rtx_last_error(None).decode()}")
        self._h = h
        self.device = device
        self.devices = list(devices) if devices is not None else [device]
        self.width = self.height = 0

    def _check(self, rc, what):
        if rc != 0:
            raise RtxError(f"{what} = {rc}: {self._L.rtx_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.rtx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        self._check(self._L.rtx_set_option(self._h, key.encode(), int(value)), f"rtx_set_option({key})")

    def set_stream(self, cuda_stream: int):
        """Run on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream); 0/None = own stream."""
        self._check(self._L.rtx_set_stream(self._h, C.c_void_p(cuda_stream or None)), "rtx_set_stream")

    def upload(self, desc_ptr):
        self._check(self._L.rtx_scene_upload(self._h, desc_ptr), "rtx_scene_upload")

    def set_camera(self, cam):
        ptr = cam if not isinstance(cam, CameraDesc) else C.pointer(cam)
        self._check(self._L.rtx_camera_set(self._h, ptr), "rtx_camera_set")
        w, h = C.c_int32(), C.c_int32()
        self._check(self._L.rtx_image_size(self._h, C.byref(w), C.byref(h)), "rtx_image_size")
        self.width, self.height = w.value, h.value

    def load(self, scene):
        """scene: NamedScene or (BuiltScene, CameraDesc)."""
        if isinstance(scene, NamedScene):
            self.upload(scene.desc_ptr)
            self.set_camera(scene.cam_ptr)
        else:
            built, cam = scene
            self.upload(built.desc_ptr)
            self.set_camera(cam)

    def clear(self):
        self._check(self._L.rtx_accum_clear(self._h), "rtx_accum_clear")

    def enable_moments(self, on=True):
        self._check(self._L.rtx_accum_enable_moments(self._h, int(on)), "rtx_accum_enable_moments")

    def render_pass(self, spp, max_depth, camera_max_depth=None, seed=1, sample_base=0):
        cmd = max_depth if camera_max_depth is None else camera_max_depth
        self._check(self._L.rtx_render_pass(self._h, spp, max_depth, cmd, seed, sample_base), "rtx_render_pass")

    def accum_device_ptr(self):
        s, q, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._check(self._L.rtx_accum_device_ptr(self._h, C.byref(s), C.byref(q), C.byref(n)), "rtx_accum_device_ptr")
        return s.value, q.value, n.value

    def resolve_rgba8(self, total_spp, out: Optional[np.ndarray] = None):
        if out is None:
            out = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        self._check(self._L.rtx_resolve_rgba8(self._h, total_spp, out.ctypes.data, out.nbytes), "rtx_resolve_rgba8")
        return out

    def resolve_accum(self, moments=False):
        n = self.width * self.height
        s = np.zeros((self.height, self.width, 3), dtype=np.float32)
        q = np.zeros((self.height, self.width, 3), dtype=np.float32) if moments else None
        cnt = np.zeros((self.height, self.width), dtype=np.uint32)
        self._check(self._L.rtx_resolve_accum(self._h, s.ctypes.data, q.ctypes.data if moments else None, cnt.ctypes.data), "rtx_resolve_accum")
        return s, q, cnt

    def trace_closest(self, rays, tmin=0.001, tmax=float("inf")):
        rays = _d(rays, 7)
        n = len(rays)
        out = dict(entry=np.full(n, -2, np.int32), prim=np.full(n, -2, np.int32), t=np.zeros(n), normal=np.zeros((n, 3)),
                   front=np.zeros(n, np.uint8), uv=np.zeros((n, 2)), p=np.zeros((n, 3)))
        self._check(self._L.rtx_trace_closest(self._h, rays.ctypes.data, n, tmin, tmax, out["entry"].ctypes.data, out["prim"].ctypes.data,
                                              out["t"].ctypes.data, out["normal"].ctypes.data, out["front"].ctypes.data,
                                              out["uv"].ctypes.data, out["p"].ctypes.data), "rtx_trace_closest")
        return out

    def camera_rays(self, ij, sq, disk, tm):
        ij, sq, disk, tm = _i(ij), _d(sq, 2), _d(disk, 2), _d(tm)
        n = len(tm)
        out = np.zeros((n, 7))
        self._check(self._L.rtx_camera_rays(self._h, ij.ctypes.data, sq.ctypes.data, disk.ctypes.data, tm.ctypes.data, n, out.ctypes.data),
                    "rtx_camera_rays")
        return out

    def hdri_sample(self, xi):
        xi = _d(xi, 2)
        n = len(xi)
        d, e, p = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
        self._check(self._L.rtx_hdri_sample(self._h, xi.ctypes.data, n, d.ctypes.data, e.ctypes.data, p.ctypes.data), "rtx_hdri_sample")
        return d, e, p

    def hdri_pdf(self, dirs):
        dirs = _d(dirs, 3)
        p = np.zeros(len(dirs))
        self._check(self._L.rtx_hdri_pdf(self._h, dirs.ctypes.data, len(dirs), p.ctypes.data), "rtx_hdri_pdf")
        return p

    def hdri_lookup(self, dirs):
        dirs = _d(dirs, 3)
        rgb = np.zeros((len(dirs), 3))
        self._check(self._L.rtx_hdri_lookup(self._h, dirs.ctypes.data, len(dirs), rgb.ctypes.data), "rtx_hdri_lookup")
        return rgb

    def hdri_total_power(self):
        v = C.c_double()
        self._check(self._L.rtx_hdri_total_power(self._h, C.byref(v)), "rtx_hdri_total_power")
        return v.value

    def mesh_test_order(self, n_tris):
        """rtx_mesh_test_order: the ranks the device derived for the mesh triangles of the scene just loaded."""
        out = np.empty(int(n_tris), dtype=np.int32)
        self._check(self._L.rtx_mesh_test_order(self._h, out.ctypes.data, int(n_tris)), "rtx_mesh_test_order")
        return out

    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.rtx_get_stats(self._h, C.byref(s)), "rtx_get_stats")
        return s.as_dict()


# BASELINE.json configs: scene pose/lens/background from the scene function; width, aspect, spp, depth from the config
# (SURVEY.md §8d). 600x338 in BASELINE/README is 600x337 by the reference's own rule (rt/camera.go:299).
CONFIGS = {
    "cornell": dict(scene="cornell", width=400, aspect=16.0 / 9.0, spp=10, depth=10),
    "random": dict(scene="random", width=600, aspect=16.0 / 9.0, spp=100, depth=50),
    "cornell-glossy": dict(scene="cornell-glossy", width=600, aspect=16.0 / 9.0, spp=256, depth=5),
    "cornell-lucy": dict(scene="cornell-lucy", width=1200, aspect=16.0 / 9.0, spp=500, depth=50),
    "hdri-test": dict(scene="hdri-test", width=3840, aspect=16.0 / 9.0, spp=1024, depth=20),
}
# the other scene functions of rt/scenes.go the device vocabulary covers, at the scenes' own resolution / quality
CONFIGS.update({
    "checkered": dict(scene="checkered", width=600, aspect=16.0 / 9.0, spp=100, depth=50),
    "simple": dict(scene="simple", width=400, aspect=16.0 / 9.0, spp=100, depth=50),
    "quads": dict(scene="quads", width=400, aspect=1.0, spp=100, depth=50),
    "glossy-metal": dict(scene="glossy-metal", width=640, aspect=16.0 / 9.0, spp=100, depth=10),
    "cornell-smoke": dict(scene="cornell-smoke", width=600, aspect=1.0, spp=150, depth=5),
    "perlin": dict(scene="perlin", width=600, aspect=16.0 / 9.0, spp=100, depth=50),
    "earth": dict(scene="earth", width=800, aspect=16.0 / 9.0, spp=100, depth=50),
    "primitives": dict(scene="primitives", width=800, aspect=16.0 / 9.0, spp=300, depth=25),
})


def stats_bar(pix: np.ndarray, spp: int, depth: int, seconds: float, workers: int) -> str:
    """drawStatsToFramebuffer (rt/bucket_renderer.go:375-407) on an [H, W, 4] uint8 image, in place; returns the stats line."""
    assert pix.dtype == np.uint8 and pix.ndim == 3 and pix.shape[2] == 4 and pix.flags.c_contiguous
    buf = C.create_string_buffer(256)
    host().rth_stats_bar(pix.ctypes.data, pix.shape[1], pix.shape[0], spp, depth, seconds, workers, buf)
    return buf.value.decode()


def parse_obj(path: str, threads: int = 0):
    """The text parse of the host mirror's LoadOBJ (rt/obj_loader.go:15-102) on `threads` threads (0 = all, 1 = the reference's
    sequential scan): (vertices [n, 3] float64, triangle vertex indices [m, 3] uint32, seconds). Raises RuntimeError with the
    reference's message on malformed input."""
    H = host()
    nv, nt, sec = C.c_int64(0), C.c_int64(0), C.c_double(0)
    if H.rth_parse_obj(path.encode(), threads, C.byref(nv), C.byref(nt), None, 0, None, 0, C.byref(sec)) != 0:
        raise RuntimeError(H.rth_last_error().decode())
    v, t = np.zeros((nv.value, 3), np.float64), np.zeros((nt.value, 3), np.uint32)
    if H.rth_parse_obj(path.encode(), threads, C.byref(nv), C.byref(nt), v.ctypes.data, nv.value, t.ctypes.data, nt.value, C.byref(sec)) != 0:
        raise RuntimeError(H.rth_last_error().decode())
    return v, t, sec.value


def config_scene(name: str, width: Optional[int] = None, spp: Optional[int] = None, depth: Optional[int] = None, seed: int = 0x5EED) -> NamedScene:
    cfg = CONFIGS[name]
    return NamedScene(cfg["scene"], width or cfg["width"], cfg["aspect"], spp or cfg["spp"], depth or cfg["depth"], seed=seed)
