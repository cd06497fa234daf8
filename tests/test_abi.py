"""CPU-side checks of the drop-in boundary: the ctypes mirror matches include/functracer_b200.h
field for field, the shared library loads and exports every declared symbol, and the product
path fails loudly (no CPU fallback) when no GPU is visible.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from functracer_b200 import abi, api, frontend, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "functracer_b200.h")

STRUCTS = {
    "ftb_node": abi.Node, "ftb_transform": abi.Transform, "ftb_material": abi.Material, "ftb_texture": abi.Texture,
    "ftb_image": abi.Image, "ftb_bsp_node": abi.BspNode, "ftb_bsp_leaf": abi.BspLeaf, "ftb_mesh": abi.Mesh,
    "ftb_light": abi.Light, "ftb_scene_desc": abi.SceneDesc, "ftb_camera": abi.Camera,
    "ftb_render_params": abi.RenderParams, "ftb_debug_out": abi.DebugOut, "ftb_stats": abi.Stats,
}


def test_struct_layouts_match_header():
    """sizeof + offsetof of every field, as the C compiler sees the header."""
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, "int main(void){"]
    for cname, cls in STRUCTS.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "l.c"), os.path.join(d, "l")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-o", exe, src])
        out = subprocess.check_output([exe], text=True)
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, cls in STRUCTS.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_header_declares_exactly_the_exports():
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(ftb_[a-z_0-9]+)\s*\(", text))
    assert declared == set(abi.EXPORTS)
    m = re.search(r"#define FTB_ABI_VERSION (\d+)", text)
    assert int(m.group(1)) == abi.ABI_VERSION


def test_library_loads_and_exports_every_symbol():
    L = api.lib()
    for name in abi.EXPORTS:
        assert hasattr(L, name), name
    assert L.ftb_abi_version() == abi.ABI_VERSION


def test_enums_match_header():
    text = open(HEADER).read()
    for name, val in [("FTB_NODE_EXCLUDE", abi.NODE_EXCLUDE), ("FTB_PRIM_TRIANGLE", abi.PRIM_TRIANGLE),
                      ("FTB_TEX_ROTATE", abi.TEX_ROTATE), ("FTB_LIGHT_POINT", abi.LIGHT_POINT),
                      ("FTB_OUT_RGBA8", abi.OUT_RGBA8)]:
        m = re.search(r"%s\s*=\s*(\d+)" % name, text)
        assert m and int(m.group(1)) == val, name
    assert abi.TILE_W == 16 and "#define FTB_TILE_W 16" in text


def test_no_cpu_fallback_without_a_gpu():
    """On a box without a GPU the product path must refuse, not fall back."""
    if api.device_count() > 0:
        pytest.skip("GPU present")
    parsed = frontend.ParsedScene(scenes.hollow_sphere(res=(8, 8), spp=1), scenes.asset_dir())
    with pytest.raises(api.FtbError) as e:
        api.Scene(parsed)
    assert e.value.status == abi.ERR_NO_DEVICE


def test_argument_validation_needs_no_device():
    p = api.make_params(0, 10, 1, [0.0, 0.0])
    with pytest.raises(api.FtbError) as e:
        api.tile_buffer_bytes(p)
    assert e.value.status == abi.ERR_BAD_ARG
    p = api.make_params(33, 17, 2, [0.0] * 4, shard_index=1, shard_count=2)
    # 3 x 2 tiles, shard 1 of 2 owns tiles 1, 3, 5
    assert api.tile_buffer_bytes(p) == 3 * 256 * 3 * 4
    p = api.make_params(33, 17, 2, [0.0] * 4, precision=abi.PRECISION_FP64_VERIFY)
    assert api.tile_buffer_bytes(p) == 6 * 256 * 3 * 8


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under functracer_b200/ may reference it."""
    pkg = os.path.join(ROOT, "functracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "ftb_oracle" not in text and "oracle/" not in text and "ftbo_" not in text, os.path.join(dirpath, f)
