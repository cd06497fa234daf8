"""Randomised scenes (seeded, built from the reference grammar: nested transforms, materials, textures, hue shifts,
ignoreLight, groups, every CSG operator, every primitive, all three light kinds) rendered by the CUDA path and by
the CPU oracle.  The bundled scenes exercise a handful of graph shapes; this walks the space between them."""
import numpy as np
import pytest

from functracer_b200 import abi, api, frontend
from oracle import ftb_oracle as orc
from util import parse

pytestmark = pytest.mark.gpu

PRIMS = ["sphere", "cube", "cylinder", "solidCylinder", "cone", "circle", "square"]


def _num(x):
    s = "%.4f" % x
    return s if not s.startswith("-0.0000") else "0.0000"


def _t(rng, lo, hi):
    return "(" + ",".join(_num(rng.uniform(lo, hi)) for _ in range(3)) + ")"


def _colour(rng):
    return "(" + ",".join(_num(rng.uniform(0.05, 1.0)) for _ in range(3)) + ")"


def _material(rng):
    s = "material diffuse %s " % _colour(rng)
    if rng.random() < 0.3:
        s += "roughness %s " % _num(rng.uniform(0.1, 0.8))
    s += "reflectance %s shineyness %s" % (_num(rng.choice([0.0, 0.0, 0.3, 0.6])), _num(float(rng.choice([0, 0, 5, 20]))))
    return s


def _texture(rng):
    t = "grid %s %s" % (_colour(rng), _colour(rng))
    if rng.random() < 0.5:
        t = "(scale (%s, %s) %s)" % (_num(rng.uniform(0.1, 0.9)), _num(rng.uniform(0.1, 0.9)), t)
    if rng.random() < 0.3:
        t = "(rotate %s %s)" % (_num(rng.uniform(5, 85)), t)
    return t


def _geometry(rng, depth, in_csg=False):
    r = rng.random()
    if depth <= 0 or r < 0.22:
        return rng.choice(PRIMS)
    if r < 0.50:  # transform
        k = rng.random()
        if k < 0.4:
            f = "translate %s" % _t(rng, -1.5, 1.5)
        elif k < 0.7:
            f = "scale %s " % (_num(rng.uniform(0.4, 1.6)) if rng.random() < 0.5 else "(%s,%s,%s)" % tuple(_num(rng.uniform(0.4, 1.6)) for _ in range(3)))
        else:
            f = "rotate %s %s" % (_t(rng, -1, 1).replace("0.0000", "0.3000"), _num(rng.uniform(-170, 170)))
        return "(%s %s)" % (f, _geometry(rng, depth - 1, in_csg))
    if r < 0.62:
        return "(%s %s)" % (_material(rng), _geometry(rng, depth - 1, in_csg))
    if r < 0.68:
        return "(texture %s %s)" % (_texture(rng), _geometry(rng, depth - 1, in_csg))
    if r < 0.72:
        return "(hueShift 1 %s)" % _geometry(rng, depth - 1, in_csg)
    if r < 0.75:
        return "(ignoreLight %s)" % _geometry(rng, depth - 1, in_csg)
    if r < 0.90:
        op = rng.choice(["union", "subtract", "intersect", "exclude"])
        # operands get their own small random offsets: canonical primitives share planes (y = 0 caps, cube faces), and
        # exactly coincident surfaces are tie cases whose order flips with the last ulp of t (DESIGN.md "grazing")
        a = "(translate %s %s)" % (_t(rng, -0.35, 0.35), _geometry(rng, depth - 1, True))
        b = "(translate %s %s)" % (_t(rng, -0.35, 0.35), _geometry(rng, depth - 1, True))
        return "(%s %s %s)" % (op, a, b)
    n = int(rng.integers(0, 4))
    return "(group %s)" % " ".join("(translate %s %s)" % (_t(rng, -0.35, 0.35), _geometry(rng, depth - 1, in_csg)) for _ in range(n))


def _scene(seed):
    rng = np.random.default_rng(seed)
    cam = "camera pos (%s,%s,%s) lookat (0,0,0) up (0,1,0) fov 55 ratio 1" % (_num(rng.uniform(-1, 1)), _num(rng.uniform(0.5, 2.5)), _num(rng.uniform(-6.5, -5)))
    if rng.random() < 0.2:  # depth of field (Image.fs:91-94)
        cam += " focus (%s,%s)" % (_num(rng.uniform(4, 8)), _num(rng.uniform(0.5, 3)))
    samples = "samples corner" if rng.random() < 0.1 else "samples %d" % int(rng.integers(1, 4))
    objs = []
    for _ in range(int(rng.integers(1, 5))):
        objs.append("(translate %s %s)" % (_t(rng, -1.8, 1.8), _geometry(rng, 4)))
    if rng.random() < 0.2:  # a mesh (never a CSG operand: Triangle.fs:62 emits no t < 0 crossings)
        objs.append('(%s (translate %s (scale %s (translate (0.017,-0.11,0) %s "bunny_tiny.ply"))))' % (
            _material(rng), _t(rng, -1.5, 1.5), _num(rng.uniform(6, 12)), rng.choice(["mesh", "bspMesh 0", "bspMesh 3"])))
    if rng.random() < 0.15:  # image texture on a sphere
        objs.append('(texture image "moon.ppm" (translate %s sphere))' % _t(rng, -1.5, 1.5))
    if rng.random() < 0.7:
        objs.append("(%s (translate (0,-2.2,0) plane))" % _material(rng))
    lights = []
    for _ in range(int(rng.integers(1, 4))):
        k = rng.random()
        if k < 0.4:
            lights.append("positional pos %s falloff (1,0.02,0.01) colour %s" % (_t(rng, -5, 5).replace("(", "(", 1), _colour(rng)))
        elif k < 0.7:
            lights.append("directional dir (%s,%s,%s) colour %s" % (_num(rng.uniform(-1, 1)), _num(rng.uniform(-1.5, -0.3)), _num(rng.uniform(-1, 1)), _colour(rng)))
        else:
            lights.append("softdirectional dir (%s,%s,%s) samples %d scatter %s colour %s" % (_num(rng.uniform(-1, 1)), _num(rng.uniform(-1.5, -0.3)), _num(rng.uniform(-1, 1)),
                                                                                           int(rng.integers(1, 4)), _num(rng.uniform(2, 30)), _colour(rng)))
    return cam + "\n" + samples + "\nres 128 96\n\n" + "\n\n".join(objs) + "\n\n" + "\n".join(lights) + "\n"


# Seeds whose frames miss the default bars (FP32: colour within 1/255 on >= 99.9 % of the 128x96 pixels and <= 0.5 % primary
# id mismatches; FP64-verify: colour within 1e-6 on >= 99.9 %, <= 0.1 % id mismatches), each with the reason found by looking
# at the frame (tools/fuzz_debug.py SEED) and the bar it does meet: seed -> (ok64, mism64, ok32, mism32, why).
KNOWN = {
}


@pytest.mark.parametrize("seed", range(160))
def test_random_scene(seed):
    text = _scene(1000 + seed)
    sc = parse(text)
    jit = frontend.jitter_pattern(seed + 1, sc.spp)
    ref = orc.render(sc, orc.make_params(sc.width, sc.height, sc.spp, jit, seed=77, sampling=sc.sampling))
    try:
        scene = api.Scene(sc)
    except api.FtbError as e:
        assert e.status == abi.ERR_UNSUPPORTED, text  # e.g. CSG nesting beyond the documented limits
        pytest.skip("unsupported by the device path: %s" % e)
    with scene:
        try:
            g64 = scene.render(sc.width, sc.height, sc.spp, jit, seed=77, precision=abi.PRECISION_FP64_VERIFY, debug=True, sampling=sc.sampling)
            g32 = scene.render(sc.width, sc.height, sc.spp, jit, seed=77, precision=abi.PRECISION_FP32, debug=True, sampling=sc.sampling)
        except api.FtbError as e:
            assert e.status == abi.ERR_HIT_OVERFLOW, text
            pytest.skip("hit-stack overflow reported: %s" % e)
    finite = np.isfinite(ref["rgb"]).all(axis=-1)  # NaN pixels (pow of a negative base, acos of 1 + ulp) are grazing cases
    mism64 = float((g64["prim"] != ref["prim"]).mean())
    d64 = np.abs(g64["rgb"] - ref["rgb"]).max(axis=-1)
    d32 = np.abs(g32["rgb"] - ref["rgb"]).max(axis=-1)
    ok64 = float(((d64 <= 1e-6) | ~finite).mean())
    ok32 = float(((d32 <= 1.0 / 255.0) | ~finite).mean())
    mism32 = float((g32["prim"] != ref["prim"]).mean())
    print("seed %d: fp64 id-mismatch %.2e colour-ok %.4f | fp32 id-mismatch %.2e colour-ok %.4f | leaves %d" % (seed, mism64, ok64, mism32, ok32, sc.desc.n_nodes))
    b_ok64, b_m64, b_ok32, b_m32 = KNOWN.get(seed, (0.999, 1e-3, 0.999, 5e-3, ""))[:4]
    assert mism64 <= b_m64 and ok64 >= b_ok64, text
    assert mism32 <= b_m32 and ok32 >= b_ok32, text
