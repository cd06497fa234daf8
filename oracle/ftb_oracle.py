"""ctypes face of oracle/libftb_oracle.so — TEST INFRASTRUCTURE (see ftb_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package functracer_b200 never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from functracer_b200 import abi  # noqa: E402  (struct layouts of the shared C ABI header)

_LIB = None


class OracleHit(C.Structure):
    _fields_ = [("t", C.c_double), ("p", C.c_double * 3), ("n", C.c_double * 3), ("uv", C.c_double * 2),
                ("colour", C.c_double * 3), ("roughness", C.c_double), ("reflectance", C.c_double),
                ("shineyness", C.c_double), ("apply_lighting", C.c_int32), ("prim", C.c_int32),
                ("sub", C.c_int32), ("reserved", C.c_int32)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libftb_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.ftbo_last_error.restype = C.c_char_p
        L.ftbo_render.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), dp,
                                  C.POINTER(abi.DebugOut), C.POINTER(abi.Stats), C.c_int]
        L.ftbo_render_window.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(abi.Camera), C.POINTER(abi.RenderParams),
                                         C.c_int, C.c_int, C.c_int, C.c_int, dp, C.POINTER(abi.DebugOut),
                                         C.POINTER(abi.Stats), C.c_int]
        L.ftbo_shade_rays.argtypes = [C.POINTER(abi.SceneDesc), dp, C.c_int64, C.POINTER(abi.RenderParams), dp,
                                      C.POINTER(abi.DebugOut), C.POINTER(abi.Stats), C.c_int]
        L.ftbo_node_hits.argtypes = [C.POINTER(abi.SceneDesc), C.c_int, dp, dp, C.POINTER(OracleHit), C.c_int]
        L.ftbo_quadratic.argtypes = [C.c_double, C.c_double, C.c_double, dp]
        L.ftbo_aabb_intersects.argtypes = [dp, dp, dp, dp]
        L.ftbo_attenuate.argtypes = [dp, C.c_double]
        L.ftbo_attenuate.restype = C.c_double
        L.ftbo_texture.argtypes = [C.POINTER(abi.SceneDesc), C.c_int, C.c_double, C.c_double, dp]
        L.ftbo_texture.restype = None
        L.ftbo_to_byte.argtypes = [C.c_double]
        L.ftbo_to_byte.restype = C.c_uint8
        L.ftbo_hue_shift.argtypes = [dp, dp]
        L.ftbo_hue_shift.restype = None
        L.ftbo_jitter_vector.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, dp, dp]
        L.ftbo_jitter_vector.restype = None
        L.ftbo_lambert.argtypes = [dp] * 5
        L.ftbo_lambert.restype = None
        L.ftbo_specular.argtypes = [dp, dp, dp, dp, C.c_double, dp]
        L.ftbo_specular.restype = None
        L.ftbo_rough_diffuse.argtypes = [dp, dp, dp, dp, C.c_double, dp]
        L.ftbo_rough_diffuse.restype = None
        L.ftbo_primary_ray.argtypes = [C.POINTER(abi.Camera), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, dp]
        L.ftbo_primary_ray.restype = None
        L.ftbo_render_windows_packed.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), C.c_int, C.POINTER(C.c_int),
                                                 dp, C.POINTER(abi.DebugOut), C.POINTER(abi.Stats), C.c_int]
        L.ftbo_render_windows_packed.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _vec(v):
    return np.ascontiguousarray(v, dtype=np.float64)


def make_params(width, height, spp, jitter_xy=None, sampling=abi.SAMPLING_JITTER, recursion_limit=8, seed=1234,
                precision=abi.PRECISION_FP32, out_format=abi.OUT_RGB_F64, shard_index=0, shard_count=1, n_gpus=0,
                collect_stats=0):
    p = abi.RenderParams()
    p.width, p.height, p.spp, p.sampling = width, height, spp, sampling
    keep = None
    if jitter_xy is not None:
        keep = np.ascontiguousarray(jitter_xy, dtype=np.float64)
        assert keep.size == 2 * spp
        p.jitter_xy = _dp(keep)
    p.recursion_limit, p.precision, p.seed, p.out_format = recursion_limit, precision, seed, out_format
    p.shard_index, p.shard_count, p.n_gpus, p.collect_stats = shard_index, shard_count, n_gpus, collect_stats
    p._keep = keep  # keep the jitter array alive
    return p


def render(scene, params, window=None, threads=0, debug=True):
    """ftbo_render on a functracer_b200.frontend.ParsedScene.  Returns dict(rgb[H,W,3], prim, sub, t, stats)."""
    L = lib()
    W, H = params.width, params.height
    rgb = np.zeros((H, W, 3), dtype=np.float64)
    n = (W + 1) * (H + 1) if params.sampling == abi.SAMPLING_CORNER else W * H * params.spp
    dbg = None
    prim = sub = t = None
    if debug:
        prim = np.full(n, -2, dtype=np.int32)
        sub = np.zeros(n, dtype=np.int32)
        t = np.zeros(n, dtype=np.float64)
        dbg = abi.DebugOut(prim.ctypes.data_as(C.POINTER(C.c_int32)), sub.ctypes.data_as(C.POINTER(C.c_int32)), _dp(t))
    st = abi.Stats()
    x0, y0, x1, y1 = window if window else (0, 0, W, H)
    rc = L.ftbo_render_window(scene.desc_ptr, scene.camera_ptr, C.byref(params), x0, y0, x1, y1, _dp(rgb),
                              C.byref(dbg) if dbg else None, C.byref(st), threads)
    if rc != 0:
        raise RuntimeError("oracle: %s" % L.ftbo_last_error().decode())
    return dict(rgb=rgb, prim=prim, sub=sub, t=t, stats=st)


def render_windows(scene, params, windows, threads=0, debug=True):
    """Several pixel windows [(x0, y0, x1, y1), ...] of one frame in one call, outputs PACKED per window.  Returns
    dict(windows=[dict(rect, rgb[h,w,3], prim[h,w,spp], sub[h,w,spp])...], stats, seconds)."""
    import time
    L = lib()
    W, H, spp = params.width, params.height, params.spp
    rects = np.ascontiguousarray([[max(0, x0), max(0, y0), min(W, x1), min(H, y1)] for x0, y0, x1, y1 in windows], dtype=np.int32).reshape(-1, 4)
    npix = int(((rects[:, 2] - rects[:, 0]).clip(0) * (rects[:, 3] - rects[:, 1]).clip(0)).sum())
    rgb = np.zeros((npix, 3), dtype=np.float64)
    prim = sub = dbg = None
    if debug:
        prim = np.full(npix * spp, -2, dtype=np.int32)
        sub = np.zeros(npix * spp, dtype=np.int32)
        dbg = abi.DebugOut(prim.ctypes.data_as(C.POINTER(C.c_int32)), sub.ctypes.data_as(C.POINTER(C.c_int32)), None)
    st = abi.Stats()
    t0 = time.perf_counter()
    rc = L.ftbo_render_windows_packed(scene.desc_ptr, scene.camera_ptr, C.byref(params), len(rects), rects.ctypes.data_as(C.POINTER(C.c_int)), _dp(rgb),
                                      C.byref(dbg) if dbg else None, C.byref(st), threads)
    secs = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError("oracle: %s" % L.ftbo_last_error().decode())
    out, at = [], 0
    for x0, y0, x1, y1 in rects.tolist():
        w, h = max(0, x1 - x0), max(0, y1 - y0)
        d = dict(rect=(x0, y0, x1, y1), rgb=rgb[at:at + w * h].reshape(h, w, 3))
        if debug:
            d["prim"] = prim[at * spp:(at + w * h) * spp].reshape(h, w, spp)
            d["sub"] = sub[at * spp:(at + w * h) * spp].reshape(h, w, spp)
        out.append(d)
        at += w * h
    return dict(windows=out, stats=st, seconds=secs)


def shade_rays(scene, rays_od, params, threads=0):
    L = lib()
    rays = np.ascontiguousarray(rays_od, dtype=np.float64).reshape(-1, 6)
    n = rays.shape[0]
    rgb = np.zeros((n, 3), dtype=np.float64)
    prim = np.full(n, -2, dtype=np.int32)
    sub = np.zeros(n, dtype=np.int32)
    t = np.zeros(n, dtype=np.float64)
    dbg = abi.DebugOut(prim.ctypes.data_as(C.POINTER(C.c_int32)), sub.ctypes.data_as(C.POINTER(C.c_int32)), _dp(t))
    st = abi.Stats()
    rc = L.ftbo_shade_rays(scene.desc_ptr, _dp(rays), n, C.byref(params), _dp(rgb), C.byref(dbg), C.byref(st), threads)
    if rc != 0:
        raise RuntimeError("oracle: %s" % L.ftbo_last_error().decode())
    return dict(rgb=rgb, prim=prim, sub=sub, t=t, stats=st)


def node_hits(scene, o, d, node=-1, max_hits=256):
    """All hits of a node (default: the scene root) in the reference's sequence order."""
    L = lib()
    buf = (OracleHit * max_hits)()
    n = L.ftbo_node_hits(scene.desc_ptr, node, _dp(_vec(o)), _dp(_vec(d)), buf, max_hits)
    if n < 0:
        raise RuntimeError("oracle: %s" % L.ftbo_last_error().decode())
    out = []
    for i in range(min(n, max_hits)):
        h = buf[i]
        out.append(dict(t=h.t, p=tuple(h.p), n=tuple(h.n), uv=tuple(h.uv), colour=tuple(h.colour), roughness=h.roughness,
                        reflectance=h.reflectance, shineyness=h.shineyness, apply_lighting=bool(h.apply_lighting),
                        prim=h.prim, sub=h.sub))
    return out


def quadratic(a, b, c):
    out = np.zeros(2)
    n = lib().ftbo_quadratic(a, b, c, _dp(out))
    return list(out[:n])


def aabb_intersects(bmin, bmax, o, d):
    return bool(lib().ftbo_aabb_intersects(_dp(_vec(bmin)), _dp(_vec(bmax)), _dp(_vec(o)), _dp(_vec(d))))


def attenuate(falloff, distance):
    return lib().ftbo_attenuate(_dp(_vec(falloff)), distance)


def texture(scene, tex, u, v):
    out = np.zeros(3)
    lib().ftbo_texture(scene.desc_ptr, tex, u, v, _dp(out))
    return tuple(out)


def to_byte(c):
    return int(lib().ftbo_to_byte(c))


def hue_shift(c):
    out = np.zeros(3)
    lib().ftbo_hue_shift(_dp(_vec(c)), _dp(out))
    return tuple(out)


def jitter_vector(seed, sample, depth, light, idx, max_angle, v):
    out = np.zeros(3)
    lib().ftbo_jitter_vector(seed, sample, depth, light, idx, max_angle, _dp(_vec(v)), _dp(out))
    return out


def lambert(n, ld, lc, colour):
    out = np.zeros(3)
    lib().ftbo_lambert(_dp(_vec(n)), _dp(_vec(ld)), _dp(_vec(lc)), _dp(_vec(colour)), _dp(out))
    return tuple(out)


def specular(n, ld, lc, view_d, shineyness):
    out = np.zeros(3)
    lib().ftbo_specular(_dp(_vec(n)), _dp(_vec(ld)), _dp(_vec(lc)), _dp(_vec(view_d)), shineyness, _dp(out))
    return tuple(out)


def rough_diffuse(n, ld, view_d, colour, roughness):
    out = np.zeros(3)
    lib().ftbo_rough_diffuse(_dp(_vec(n)), _dp(_vec(ld)), _dp(_vec(view_d)), _dp(_vec(colour)), roughness, _dp(out))
    return tuple(out)


def primary_ray(camera, width, height, px, py, jx=0.0, jy=0.0):
    out = np.zeros(6)
    lib().ftbo_primary_ray(C.byref(camera), width, height, px, py, jx, jy, _dp(out))
    return out


def quantise(rgb):
    """Image.write's toByte on a whole frame (Image.fs:36): clamp, *255, truncate."""
    c = np.where(rgb > 1.0, 1.0, np.where(rgb < 0.0, 0.0, rgb))
    return (c * 255.0).astype(np.uint8)
