"""Which kernel a scene runs on is decided on the host: lower.cpp derives the scene's feature mask, api.cu launches the smallest
compiled variant that covers it (csrc/Makefile F32_FEATS) and falls back to the generic 0xfff kernel - correct, but 1.3-2x
slower - when no specialised variant does.  A drifting feature bit would therefore cost the headline silently.  This test
runs the real lowering (lower.cpp compiled with g++ into a throw-away probe, no device involved) on every BASELINE.json config
and on the bundled scenes' small test sizes, replays api.cu's choice, and pins the variant each config is documented to use
(BENCH.md).  CPU only."""
import ctypes as C
import os
import re
import subprocess

import pytest

from functracer_b200 import abi, frontend, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "functracer_b200", "csrc")

PROBE = r"""
#include "%(csrc)s/cuda/lower.h"
#include <cstring>
extern "C" int ftb_probe_lower(const ftb_scene_desc* d, unsigned* features, int* counts, char* err, int errlen)
{
    ftb::Lowered L;
    std::string e;
    const int rc = ftb::lower_scene(*d, L, e, true);  // with the host-built mesh index: its size decides the large-mesh walk
    std::strncpy(err, e.c_str(), (size_t)errlen - 1);
    err[errlen - 1] = 0;
    if (rc != 0) return rc;
    *features = L.features;
    counts[0] = (int)L.items.size();
    counts[1] = (int)L.leaves.size();
    counts[2] = (int)L.bvh_tri.size();
    counts[3] = (int)L.ops.size();
    return 0;
}
"""

FT_TABLE, FT_RNG, FT_MESHPK, FT_ALL = 0x200, 0x40, 0x800, 0xfff
LARGE_MESH = 32768  # api.cu kLargeMesh


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    d = tmp_path_factory.mktemp("lower_probe")
    src, so = os.path.join(str(d), "probe.cpp"), os.path.join(str(d), "libprobe.so")
    open(src, "w").write(PROBE % dict(csrc=CSRC))
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src, os.path.join(CSRC, "cuda", "lower.cpp")])
    lib = C.CDLL(so)
    lib.ftb_probe_lower.argtypes = [C.POINTER(abi.SceneDesc), C.POINTER(C.c_uint), C.POINTER(C.c_int), C.c_char_p, C.c_int]
    lib.ftb_probe_lower.restype = C.c_int
    return lib


def compiled_variants():
    mk = open(os.path.join(CSRC, "Makefile")).read()
    return [int(x, 16) for x in re.search(r"^F32_FEATS := (.*)$", mk, flags=re.M).group(1).split()]


def constants_agree_with_the_sources():
    api_cu = open(os.path.join(CSRC, "cuda", "api.cu")).read()
    dev = open(os.path.join(CSRC, "cuda", "device_scene.h")).read()
    assert int(re.search(r"kLargeMesh = (\d+)", api_cu).group(1)) == LARGE_MESH
    for name, val in (("FT_TABLE", FT_TABLE), ("FT_RNG", FT_RNG), ("FT_MESHPK", FT_MESHPK), ("FT_ALL", FT_ALL)):
        assert int(re.search(r"\b%s = (0x[0-9a-f]+)" % name, dev).group(1), 16) == val, name


def pick(need, variants):
    """api.cu pickVariant: the cover with the fewest feature bits; the table bit is dropped rather than going generic."""
    def cover(n):
        best = None
        for v in variants:
            if v & n == n and (best is None or bin(v).count("1") < bin(best).count("1")):
                best = v
        return best
    v = cover(need)
    if need & FT_TABLE and (v is None or v == FT_ALL):
        w = cover(need & ~FT_TABLE)
        if w is not None and w != FT_ALL:
            return w
    return v


def lowered(probe, text):
    sc = frontend.ParsedScene(text, scenes.asset_dir())
    feats, counts, err = C.c_uint(0), (C.c_int * 4)(), C.create_string_buffer(256)
    rc = probe.ftb_probe_lower(sc.desc_ptr, C.byref(feats), counts, err, 256)
    assert rc == 0, err.value
    need = feats.value
    if counts[2] >= LARGE_MESH:
        need |= FT_MESHPK  # api.cu: a large mesh is walked by the whole warp
    if sc.camera.has_focus:
        need |= FT_RNG      # api.cu: depth of field draws random numbers
    return need, list(counts)


# the variant every BASELINE.json config is measured on (BENCH.md section 2)
EXPECTED = {
    "cfg1-sample": 0x050, "cfg2-hollow-sphere": 0x209, "cfg3-house": 0x74b, "cfg3-night-house": 0x74b, "cfg4-bunny": 0x004,
    "cfg4-bunny-d12": 0x004, "cfg4-bunny-full-d14": 0x804, "cfg5-repeat": 0x74b, "cfg5-moon": 0x030,
}


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_every_baseline_config_runs_on_its_specialised_kernel(probe, name):
    constants_agree_with_the_sources()
    variants = compiled_variants()
    assert FT_ALL in variants  # the generic kernel (also the counting kernel) must exist
    need, counts = lowered(probe, scenes.config_text(name, res=(64, 48), spp=1))  # the feature mask does not depend on the frame size
    v = pick(need, variants)
    print("%s: %d items, %d leaves, %d mesh slots, needs 0x%03x -> variant 0x%03x" % (name, counts[0], counts[1], counts[2], need, v))
    assert v == EXPECTED[name], "%s needs 0x%03x and would run on 0x%03x" % (name, need, v)
    assert v != FT_ALL


def test_the_table_is_only_asked_for_when_it_fits():
    """lower.h wantsOriginTable: at least 8 items, and (1 + lights) rows of the padded item count within 256 slots."""
    hdr = open(os.path.join(CSRC, "cuda", "lower.h")).read()
    assert re.search(r"kOriginCap = 256\b", hdr) and re.search(r"kOriginMinItems = 8\b", hdr)
    assert "((n_items + 1) & ~1)" in hdr  # the rows of an origin are padded to an even count (render.cuh tabStride)
