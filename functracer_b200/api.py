"""ctypes face of libfunctracer_b200.so — the C ABI of include/functracer_b200.h.

This is the binding the F# shim makes with P/Invoke, restated for the Python
harness.  The names follow the reference's call site (Program.fs:54-64): `Scene.render` is
generateRays + shade + blendPixels, `Scene.shade` is Shading.shade on explicit rays.

There is no fallback: if the CUDA library is missing or no GPU is visible, every call raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class FtbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("functracer_b200 status %d: %s" % (status, message))
        self.status = status


def lib():
    """Loads libfunctracer_b200.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("FTB_LIB") or os.path.join(_HERE, "libfunctracer_b200.so")  # FTB_LIB: A/B builds of the same ABI
        if not os.path.exists(path):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % path)
        L = C.CDLL(path)
        vp = C.c_void_p
        L.ftb_abi_version.restype = C.c_int
        L.ftb_device_count.restype = C.c_int
        L.ftb_last_error.restype = C.c_char_p
        L.ftb_scene_create.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(vp)]
        L.ftb_scene_create.restype = C.c_int
        L.ftb_scene_destroy.argtypes = [vp]
        L.ftb_scene_destroy.restype = None
        L.ftb_render.argtypes = [vp, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), vp, C.POINTER(abi.DebugOut),
                                 C.POINTER(abi.Stats)]
        L.ftb_render.restype = C.c_int
        L.ftb_tile_buffer_bytes.argtypes = [C.POINTER(abi.RenderParams)]
        L.ftb_tile_buffer_bytes.restype = C.c_int64
        L.ftb_render_tiles_device.argtypes = [vp, C.POINTER(abi.Camera), C.POINTER(abi.RenderParams), vp,
                                              C.POINTER(abi.DebugOut), C.POINTER(abi.Stats), vp]
        L.ftb_render_tiles_device.restype = C.c_int
        L.ftb_assemble_device.argtypes = [C.POINTER(abi.RenderParams), C.POINTER(vp), vp, vp]
        L.ftb_assemble_device.restype = C.c_int
        L.ftb_shade_rays.argtypes = [vp, C.POINTER(C.c_double), C.c_int64, C.POINTER(abi.RenderParams),
                                     C.POINTER(C.c_double), C.POINTER(abi.DebugOut), C.POINTER(abi.Stats)]
        L.ftb_shade_rays.restype = C.c_int
        L.ftb_band_rows.argtypes = [C.POINTER(abi.RenderParams), C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ftb_band_rows.restype = C.c_int
        L.ftb_assemble_rows_device.argtypes = [C.POINTER(abi.RenderParams), C.POINTER(vp), vp, C.c_int, C.c_int, vp]
        L.ftb_assemble_rows_device.restype = C.c_int
        L.ftb_host_copy_begin.argtypes = [vp, vp, vp, C.c_int64, vp]
        L.ftb_host_copy_begin.restype = C.c_int
        L.ftb_host_copy_finish.argtypes = [vp]
        L.ftb_host_copy_finish.restype = C.c_int
        L.ftb_check_overflow.argtypes = [vp, vp]
        L.ftb_check_overflow.restype = C.c_int
        if os.environ.get("FTB_LIB") and not hasattr(L, "ftb_scene_build_info"):
            L.ftb_scene_build_info = None  # an older A/B build of the same ABI (tools/ab_build.sh REV)
        else:
            L.ftb_scene_build_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
            L.ftb_scene_build_info.restype = C.c_int
        if L.ftb_abi_version() != abi.ABI_VERSION:
            raise RuntimeError("ABI version mismatch: library %d, bindings %d" % (L.ftb_abi_version(), abi.ABI_VERSION))
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise FtbError(rc, lib().ftb_last_error().decode(errors="replace"))


def make_params(width, height, spp, jitter_xy=None, sampling=abi.SAMPLING_JITTER, recursion_limit=8, seed=1234,
                precision=abi.PRECISION_FP32, out_format=abi.OUT_RGB_F64, shard_index=0, shard_count=1, n_gpus=0,
                collect_stats=0, band_index=0, band_count=0):
    p = abi.RenderParams()
    p.width, p.height, p.spp, p.sampling = width, height, spp, sampling
    keep = None
    if jitter_xy is not None:
        keep = np.ascontiguousarray(jitter_xy, dtype=np.float64)
        assert keep.size == 2 * spp
        p.jitter_xy = keep.ctypes.data_as(C.POINTER(C.c_double))
    p.recursion_limit, p.precision, p.seed, p.out_format = recursion_limit, precision, seed, out_format
    p.shard_index, p.shard_count, p.n_gpus, p.collect_stats = shard_index, shard_count, n_gpus, collect_stats
    p.band_index, p.band_count = band_index, band_count
    p._keep = keep
    return p


_OUT_DTYPE = {abi.OUT_RGB_F64: (np.float64, 3), abi.OUT_RGB_F32: (np.float32, 3), abi.OUT_RGBA8: (np.uint8, 4)}


class Scene:
    """An ftb_scene: the flattened SceneGraph resident on the GPU(s)."""

    def __init__(self, parsed):
        """parsed: anything with .desc_ptr / .camera_ptr (functracer_b200.frontend.ParsedScene)."""
        self._h = C.c_void_p()
        self.parsed = parsed
        _check(lib().ftb_scene_create(parsed.desc_ptr, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().ftb_scene_destroy(self._h)
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

    # -- Program.fs:54-64: generateRays -> shade -> blendPixels ----------------------------------
    def render(self, width, height, spp, jitter_xy=None, debug=False, stats=False, out=None, camera=None, **kw):
        """Host-buffer render.  Returns dict(rgb[H,W,C], prim, sub, t, stats)."""
        if stats:
            kw.setdefault("collect_stats", 1)
        p = make_params(width, height, spp, jitter_xy, **kw)
        return self.render_params(p, debug=debug, stats=stats, out=out, camera=camera)

    def render_params(self, p, debug=False, stats=False, out=None, camera=None):
        dt, ch = _OUT_DTYPE[p.out_format]
        W, H = p.width, p.height
        if out is None:
            out = np.empty((H, W, ch), dtype=dt)
        assert out.dtype == dt and out.size == W * H * ch and out.flags["C_CONTIGUOUS"]
        n = (W + 1) * (H + 1) if p.sampling == abi.SAMPLING_CORNER else W * H * p.spp
        prim = sub = t = dbg = None
        if debug:
            prim = np.full(n, -2, dtype=np.int32)
            sub = np.zeros(n, dtype=np.int32)
            t = np.zeros(n, dtype=np.float64)
            dbg = abi.DebugOut(prim.ctypes.data_as(C.POINTER(C.c_int32)), sub.ctypes.data_as(C.POINTER(C.c_int32)),
                               t.ctypes.data_as(C.POINTER(C.c_double)))
        st = abi.Stats() if stats else None
        cam = camera if camera is not None else self.parsed.camera_ptr
        _check(lib().ftb_render(self._h, cam, C.byref(p), out.ctypes.data_as(C.c_void_p), C.byref(dbg) if dbg else None,
                                C.byref(st) if st is not None else None))
        return dict(rgb=out, prim=prim, sub=sub, t=t, stats=st)

    # -- Shading.shade (Shading.fs:141-147) on explicit rays -----------------------------------------
    def shade(self, rays_od, debug=True, stats=False, **kw):
        rays = np.ascontiguousarray(rays_od, dtype=np.float64).reshape(-1, 6)
        n = rays.shape[0]
        if stats:
            kw.setdefault("collect_stats", 1)
        p = make_params(1, 1, 1, None, **kw)
        rgb = np.zeros((n, 3), dtype=np.float64)
        prim = sub = t = dbg = None
        if debug:
            prim = np.full(n, -2, dtype=np.int32)
            sub = np.zeros(n, dtype=np.int32)
            t = np.zeros(n, dtype=np.float64)
            dbg = abi.DebugOut(prim.ctypes.data_as(C.POINTER(C.c_int32)), sub.ctypes.data_as(C.POINTER(C.c_int32)),
                               t.ctypes.data_as(C.POINTER(C.c_double)))
        st = abi.Stats() if stats else None
        dp = C.POINTER(C.c_double)
        _check(lib().ftb_shade_rays(self._h, rays.ctypes.data_as(dp), n, C.byref(p), rgb.ctypes.data_as(dp),
                                    C.byref(dbg) if dbg else None, C.byref(st) if st is not None else None))
        return dict(rgb=rgb, prim=prim, sub=sub, t=t, stats=st)

    # -- device-buffer entry points (one process per GPU; pointers are raw device addresses) -----
    def render_tiles_device(self, p, d_tiles_ptr, stream=0, stats=False, camera=None, d_dbg=None):
        """d_dbg: optional (prim_ptr, sub_ptr, t_ptr) DEVICE addresses of per-sample debug planes (0 / None = not wanted)."""
        st = abi.Stats() if stats else None
        cam = camera if camera is not None else self.parsed.camera_ptr
        dbg = None
        if d_dbg is not None:
            dbg = abi.DebugOut(C.cast(C.c_void_p(d_dbg[0] or None), C.POINTER(C.c_int32)), C.cast(C.c_void_p(d_dbg[1] or None), C.POINTER(C.c_int32)),
                               C.cast(C.c_void_p(d_dbg[2] or None), C.POINTER(C.c_double)))
        _check(lib().ftb_render_tiles_device(self._h, cam, C.byref(p), C.c_void_p(d_tiles_ptr), C.byref(dbg) if dbg is not None else None,
                                             C.byref(st) if st is not None else None, C.c_void_p(stream)))
        return st

    def build_info(self):
        """dict(bvh_on_device, bvh_build_ms, bvh_total_ms): how the scene's mesh index was built at create time."""
        on, b, t = C.c_int32(), C.c_double(), C.c_double()
        if lib().ftb_scene_build_info is None:
            return dict(bvh_on_device=False, bvh_build_ms=0.0, bvh_total_ms=0.0)
        _check(lib().ftb_scene_build_info(self._h, C.byref(on), C.byref(b), C.byref(t)))
        return dict(bvh_on_device=bool(on.value), bvh_build_ms=b.value, bvh_total_ms=t.value)

    def check_overflow(self, stream=0):
        """Waits for `stream`; raises FtbError(ERR_HIT_OVERFLOW) if a frame rendered through the device entry points on the
        current device overflowed a per-ray stack since the last check."""
        _check(lib().ftb_check_overflow(self._h, C.c_void_p(stream)))

    def host_copy_begin(self, d_src_ptr, host_array, stream=0, offset=0, nbytes=None):
        """Queues device -> host_array[offset : offset + nbytes] (bytes) behind `stream`; pageable destinations are staged
        through the scene's page-locked ring.  Complete with host_copy_finish()."""
        n = host_array.nbytes - offset if nbytes is None else nbytes
        _check(lib().ftb_host_copy_begin(self._h, C.c_void_p(d_src_ptr), C.c_void_p(host_array.ctypes.data + offset), n, C.c_void_p(stream)))

    def host_copy_finish(self):
        _check(lib().ftb_host_copy_finish(self._h))


def band_rows(p, band_index, band_count):
    y0, y1 = C.c_int(), C.c_int()
    _check(lib().ftb_band_rows(C.byref(p), band_index, band_count, C.byref(y0), C.byref(y1)))
    return y0.value, y1.value


def assemble_rows_device(p, d_tile_ptrs, d_out_ptr, y0, y1, stream=0):
    arr = (C.c_void_p * len(d_tile_ptrs))(*[C.c_void_p(x) for x in d_tile_ptrs])
    _check(lib().ftb_assemble_rows_device(C.byref(p), arr, C.c_void_p(d_out_ptr), y0, y1, C.c_void_p(stream)))


def tile_buffer_bytes(p):
    n = lib().ftb_tile_buffer_bytes(C.byref(p))
    if n < 0:
        _check(int(n))
    return int(n)


def assemble_device(p, d_tile_ptrs, d_out_ptr, stream=0):
    arr = (C.c_void_p * len(d_tile_ptrs))(*[C.c_void_p(x) for x in d_tile_ptrs])
    _check(lib().ftb_assemble_device(C.byref(p), arr, C.c_void_p(d_out_ptr), C.c_void_p(stream)))


def device_count():
    return lib().ftb_device_count()
