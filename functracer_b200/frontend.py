"""Python face of libftb_frontend.so — the host-side stand-in for SceneParser.fs / PlyParser.fs /
BspMesh.compile (see csrc/frontend/ftb_frontend.h).  Produces the ftb_scene_desc the F# flattener
would hand to the render library."""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SceneParseError(Exception):
    """Mirrors Program.readScene's failure path (Program.fs:10-16): message + exit code 1."""


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libftb_frontend.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.ftbf_last_error.restype = C.c_char_p
        L.ftbf_parse.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.ftbf_parse.restype = C.c_int
        L.ftbf_destroy.argtypes = [C.c_void_p]
        L.ftbf_destroy.restype = None
        L.ftbf_desc.argtypes = [C.c_void_p]
        L.ftbf_desc.restype = C.POINTER(abi.SceneDesc)
        L.ftbf_camera.argtypes = [C.c_void_p]
        L.ftbf_camera.restype = C.POINTER(abi.Camera)
        L.ftbf_options.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.ftbf_options.restype = None
        L.ftbf_jitter_pattern.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_double)]
        L.ftbf_jitter_pattern.restype = None
        L.ftbf_slice_triangle.argtypes = [C.POINTER(C.c_double)] * 4 + [C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ftbf_slice_triangle.restype = None
        L.ftbf_parse_colour.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
        L.ftbf_parse_colour.restype = C.c_int
        L.ftbf_write_png.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
        L.ftbf_write_png.restype = C.c_int
        _LIB = L
    return _LIB


class ParsedScene:
    """(SceneOptions, Scene) as SceneParser.parse returns them (SceneParser.fs:360-366), already
    flattened: .desc is an ftb_scene_desc whose arrays live as long as this object."""

    def __init__(self, text, asset_dir=None):
        L = lib()
        h = C.c_void_p()
        rc = L.ftbf_parse(text.encode("utf-8"), asset_dir.encode() if asset_dir else None, C.byref(h))
        if rc != 0:
            raise SceneParseError(L.ftbf_last_error().decode())
        self._h = h
        self.desc_ptr = L.ftbf_desc(h)
        self.desc = self.desc_ptr.contents
        self.camera_ptr = L.ftbf_camera(h)
        self.camera = self.camera_ptr.contents
        w, hh, spp, smp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.ftbf_options(h, C.byref(w), C.byref(hh), C.byref(spp), C.byref(smp))
        self.width, self.height, self.spp, self.sampling = w.value, hh.value, spp.value, smp.value

    def close(self):
        if getattr(self, "_h", None):
            lib().ftbf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # convenience views for tests
    def nodes(self):
        d = self.desc
        return [(d.nodes[i].kind, d.nodes[i].a, d.nodes[i].b) for i in range(d.n_nodes)]

    def group_children(self, node):
        d = self.desc
        n = d.nodes[node]
        assert n.kind == abi.NODE_GROUP
        return [d.children[n.a + i] for i in range(n.b)]


def jitter_pattern(seed, spp):
    """Jitter.pattern random Jitter.circle spp (Image.fs:101-105) from a seeded generator."""
    xy = np.zeros(2 * spp, dtype=np.float64)
    lib().ftbf_jitter_pattern(seed, spp, xy.ctypes.data_as(C.POINTER(C.c_double)))
    return xy


def slice_triangle(p0, n, tri):
    """Triangle.slice (Triangle.fs:24-41) -> (above, below) lists of 3x3 arrays."""
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    n = np.ascontiguousarray(n, dtype=np.float64)
    tri = np.ascontiguousarray(tri, dtype=np.float64).reshape(9)
    above = np.zeros(18)
    below = np.zeros(18)
    na, nb = C.c_int(), C.c_int()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lib().ftbf_slice_triangle(dp(p0), dp(n), dp(tri), dp(above), C.byref(na), dp(below), C.byref(nb))
    return ([above[9 * i:9 * i + 9].reshape(3, 3).copy() for i in range(na.value)],
            [below[9 * i:9 * i + 9].reshape(3, 3).copy() for i in range(nb.value)])


def parse_colour(text):
    rgb = (C.c_double * 3)()
    rc = lib().ftbf_parse_colour(text.encode(), rgb)
    if rc != 0:
        raise SceneParseError(lib().ftbf_last_error().decode())
    return tuple(rgb)


def write_png(path, rgba):
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w, c = rgba.shape
    assert c == 4
    rc = lib().ftbf_write_png(path.encode(), w, h, rgba.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise RuntimeError(lib().ftbf_last_error().decode())
