"""ctypes wrapper of oracle/liboracle.so — the CPU float64 restatement of the reference (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        L = C.CDLL(ORACLE_LIB)
        L.orc_scene_from_desc.restype = C.c_void_p
        L.orc_scene_from_desc.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_image_size.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_bvh_nodes.restype = C.c_int64
        L.orc_bvh_nodes.argtypes = [C.c_void_p]
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double] + [C.c_void_p] * 7
        L.orc_camera_rays.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
        L.orc_render.restype = C.c_double
        L.orc_render.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_resolve_rgba8.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_hdri_total_power.restype = C.c_double
        L.orc_hdri_total_power.argtypes = [C.c_void_p]
        L.orc_hdri_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_hdri_pdf.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_hdri_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_hdri_search_cdf.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        L.orc_load_hdr.restype = C.c_double
        L.orc_load_hdr.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]
        L.orc_aabb_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.orc_gamma_byte.argtypes = [C.c_double]
        L.orc_image_height.argtypes = [C.c_int32, C.c_double]
        L.orc_reflectance.restype = C.c_double
        L.orc_reflectance.argtypes = [C.c_double, C.c_double]
        L.orc_refract.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_texture_value.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_checker.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_prim_hit.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        _lib = L
    return _lib


def _d(a, cols=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a.reshape(-1, cols) if cols else a


class OracleScene:
    """The reference-style pointer graph rebuilt from an rtx_scene_desc / rtx_camera_desc."""

    def __init__(self, desc_ptr, cam_ptr):
        self._L = lib()
        self._keep = (desc_ptr, cam_ptr)
        self._h = self._L.orc_scene_from_desc(C.cast(desc_ptr, C.c_void_p), C.cast(cam_ptr, C.c_void_p))
        w, h = C.c_int32(), C.c_int32()
        self._L.orc_image_size(self._h, C.byref(w), C.byref(h))
        self.width, self.height = w.value, h.value

    def close(self):
        if self._h:
            self._L.orc_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trace_closest(self, rays, tmin=0.001, tmax=float("inf")):
        rays = _d(rays, 7)
        n = len(rays)
        out = dict(entry=np.full(n, -2, np.int32), prim=np.full(n, -2, np.int32), t=np.zeros(n), normal=np.zeros((n, 3)),
                   front=np.zeros(n, np.uint8), uv=np.zeros((n, 2)), p=np.zeros((n, 3)))
        self._L.orc_trace_closest(self._h, rays.ctypes.data, n, tmin, tmax, out["entry"].ctypes.data, out["prim"].ctypes.data,
                                  out["t"].ctypes.data, out["normal"].ctypes.data, out["front"].ctypes.data, out["uv"].ctypes.data,
                                  out["p"].ctypes.data)
        return out

    def camera_rays(self, ij, sq, disk, tm):
        ij = np.ascontiguousarray(np.asarray(ij, dtype=np.int32).reshape(-1))
        sq, disk, tm = _d(sq, 2), _d(disk, 2), _d(tm)
        out = np.zeros((len(tm), 7))
        self._L.orc_camera_rays(self._h, ij.ctypes.data, sq.ctypes.data, disk.ctypes.data, tm.ctypes.data, len(tm), out.ctypes.data)
        return out

    def render(self, spp, depth, seed=1, threads=0, use_atomics=True, moments=True):
        """One BucketRenderer pass. Returns dict(sum, sumsq, seconds, counters)."""
        s = np.zeros((self.height, self.width, 3))
        q = np.zeros((self.height, self.width, 3)) if moments else None
        cnt = np.zeros(5, dtype=np.int64)
        sec = self._L.orc_render(self._h, spp, depth, seed, threads, int(use_atomics), s.ctypes.data, q.ctypes.data if moments else None,
                                 cnt.ctypes.data)
        return dict(sum=s, sumsq=q, seconds=sec,
                    counters=dict(RayCount=int(cnt[0]), BVHIntersections=int(cnt[1]), SamplesComputed=int(cnt[2]),
                                  PixelsRendered=int(cnt[3]), ShadowQueries=int(cnt[4])))

    def texture_value(self, tex, u, v, p):
        pp, out = _d(p), np.zeros(3)
        lib().orc_texture_value(self._h, int(tex), float(u), float(v), pp.ctypes.data, out.ctypes.data)
        return out

    def hdri_total_power(self):
        return self._L.orc_hdri_total_power(self._h)

    def hdri_sample(self, xi):
        xi = _d(xi, 2)
        n = len(xi)
        d, e, p = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
        self._L.orc_hdri_sample(self._h, xi.ctypes.data, n, d.ctypes.data, e.ctypes.data, p.ctypes.data)
        return d, e, p

    def hdri_pdf(self, dirs):
        dirs = _d(dirs, 3)
        p = np.zeros(len(dirs))
        self._L.orc_hdri_pdf(self._h, dirs.ctypes.data, len(dirs), p.ctypes.data)
        return p

    def hdri_lookup(self, dirs):
        dirs = _d(dirs, 3)
        rgb = np.zeros((len(dirs), 3))
        self._L.orc_hdri_lookup(self._h, dirs.ctypes.data, len(dirs), rgb.ctypes.data)
        return rgb


def resolve_rgba8(sum_rgb, spp):
    s = _d(sum_rgb)
    H, W = s.shape[0], s.shape[1]
    pix = np.zeros((H, W, 4), dtype=np.uint8)
    lib().orc_resolve_rgba8(s.ctypes.data, spp, W, H, pix.ctypes.data)
    return pix


def load_hdr(path, want_pixels=False):
    w, h = C.c_int32(), C.c_int32()
    power = lib().orc_load_hdr(path.encode(), C.byref(w), C.byref(h), None)
    if power < 0:
        raise RuntimeError("orc_load_hdr failed for " + path)
    rgb = None
    if want_pixels:
        rgb = np.zeros((h.value, w.value, 3))
        lib().orc_load_hdr(path.encode(), C.byref(w), C.byref(h), rgb.ctypes.data)
    return w.value, h.value, power, rgb


def hardware_threads():
    return lib().orc_hardware_threads()


def aabb_hit(box6, ray6, tmin, tmax):
    b, r = _d(box6), _d(ray6)
    return bool(lib().orc_aabb_hit(b.ctypes.data, r.ctypes.data, tmin, tmax))


def prim_hit(kind, params, ray7, tmin, tmax):
    p, r = _d(params), _d(ray7)
    out = np.zeros(7)
    hit = lib().orc_prim_hit(kind, p.ctypes.data, r.ctypes.data, tmin, tmax, out.ctypes.data)
    return bool(hit), out


def gamma_byte(x):
    return lib().orc_gamma_byte(float(x))


def image_height(w, aspect):
    return lib().orc_image_height(int(w), float(aspect))


def reflectance(c, ri):
    return lib().orc_reflectance(float(c), float(ri))


def refract(uv, n, eta):
    a, b, o = _d(uv), _d(n), np.zeros(3)
    lib().orc_refract(a.ctypes.data, b.ctypes.data, eta, o.ctypes.data)
    return o


def checker(scale, even, odd, p):
    e, o, pp, out = _d(even), _d(odd), _d(p), np.zeros(3)
    lib().orc_checker(scale, e.ctypes.data, o.ctypes.data, pp.ctypes.data, out.ctypes.data)
    return out


def search_cdf(cdf, xi):
    c = _d(cdf)
    return lib().orc_hdri_search_cdf(c.ctypes.data, len(c), float(xi))
