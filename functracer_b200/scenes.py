"""The benchmark / parity workloads: `.scene` texts equivalent to the reference's seven bundled
scenes (/root/reference/Scenes/*.scene, described in SURVEY.md Appendix C), emitted by small
builders with the option lines the BASELINE.json configs ask for (`res W H`, `samples N`), plus
deterministic synthetic stand-ins for the three assets the reference does not ship
(bun_zipper_res4.ply, env4.jpg, the moon JPEG).

The texts use only the reference grammar (SceneParser.fs), so the F# parser accepts them too.
"""
import math
import os
import tempfile

import numpy as np

_ASSET_DIR = None


# ---------------------------------------------------------------------------------- assets
def _value_noise(h, w, rng, octaves=5):
    out = np.zeros((h, w))
    amp, total = 1.0, 0.0
    for o in range(octaves):
        gh, gw = 4 * 2 ** o + 1, 8 * 2 ** o + 1
        g = rng.random((gh, gw))
        g[:, -1] = g[:, 0]  # wrap in u so the sphere seam is continuous
        ys = np.linspace(0, gh - 1, h, endpoint=False)
        xs = np.linspace(0, gw - 1, w, endpoint=False)
        y0 = np.floor(ys).astype(int)
        x0 = np.floor(xs).astype(int)
        fy = (ys - y0)[:, None]
        fx = (xs - x0)[None, :]
        y1 = np.minimum(y0 + 1, gh - 1)
        x1 = np.minimum(x0 + 1, gw - 1)
        fy = fy * fy * (3 - 2 * fy)
        fx = fx * fx * (3 - 2 * fx)
        v = (g[y0][:, x0] * (1 - fy) * (1 - fx) + g[y0][:, x1] * (1 - fy) * fx +
             g[y1][:, x0] * fy * (1 - fx) + g[y1][:, x1] * fy * fx)
        out += amp * v
        total += amp
        amp *= 0.5
    return out / total


def _write_ppm(path, rgb):
    h, w, _ = rgb.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgb, dtype=np.uint8).tobytes())


def _env_map(path, w=2048, h=1024, seed=11):
    rng = np.random.default_rng(seed)
    n = _value_noise(h, w, rng)
    v = np.linspace(0, 1, h)[:, None]
    sky = np.stack([0.35 + 0.3 * v + 0.3 * n, 0.5 + 0.2 * v + 0.25 * n, 0.9 - 0.3 * v + 0.1 * n], axis=-1)
    ground = np.stack([0.3 + 0.4 * n, 0.25 + 0.35 * n, 0.15 + 0.2 * n], axis=-1)
    img = np.where(v[..., None] < 0.5, sky, ground)
    _write_ppm(path, np.clip(img * 255.0, 0, 255).astype(np.uint8))


def _moon_map(path, w=1024, h=512, seed=12):
    rng = np.random.default_rng(seed)
    n = _value_noise(h, w, rng, octaves=6)
    img = np.clip(0.25 + 0.7 * n, 0, 1)
    # a few craters
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(40):
        cx, cy, r = rng.integers(0, w), rng.integers(0, h), rng.integers(6, 40)
        d = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2) / r
        img = np.where(d < 1, img * (0.6 + 0.4 * d), img)
    g = (img * 255.0).astype(np.uint8)
    _write_ppm(path, np.stack([g, g, (g * 0.95).astype(np.uint8)], axis=-1))


def _blob_ply(path, n_lon, n_lat, seed=13):
    """A closed, bunny-sized blob in the exact format PlyParser.fs reads: 5 floats per vertex
    (x y z confidence intensity), faces as `3 a b c`."""
    rng = np.random.default_rng(seed)
    coef = rng.normal(size=(6, 4)) * 0.12
    centre = np.array([-0.017, 0.11, 0.0])

    def radius(theta, phi):
        r = 1.0
        for k in range(6):
            a, b, c, d = coef[k]
            r += a * np.sin((k + 1) * theta + b * 5) * np.sin(phi) * np.cos((k % 3 + 1) * phi + c * 3 + d)
        # two "ears"
        r += 0.55 * np.exp(-((theta - 1.2) ** 2 + (phi - 0.55) ** 2) / 0.03)
        r += 0.55 * np.exp(-((theta - 1.9) ** 2 + (phi - 0.55) ** 2) / 0.03)
        return 0.065 * r

    verts = [centre + np.array([0.0, radius(0.0, 0.0), 0.0])]
    for i in range(1, n_lat):
        phi = math.pi * i / n_lat
        for j in range(n_lon):
            theta = 2 * math.pi * j / n_lon
            r = radius(theta, phi)
            verts.append(centre + r * np.array([math.sin(phi) * math.cos(theta), math.cos(phi), math.sin(phi) * math.sin(theta)]))
    verts.append(centre - np.array([0.0, radius(0.0, math.pi), 0.0]))
    faces = []
    ring = lambda i, j: 1 + (i - 1) * n_lon + (j % n_lon)
    for j in range(n_lon):
        faces.append((0, ring(1, j + 1), ring(1, j)))
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            faces.append((a, b, d))
            faces.append((a, d, c))
    last = len(verts) - 1
    for j in range(n_lon):
        faces.append((last, ring(n_lat - 1, j), ring(n_lat - 1, j + 1)))
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment synthetic stand-in for bun_zipper (functracer_b200)\n")
        f.write("element vertex %d\nproperty float x\nproperty float y\nproperty float z\n" % len(verts))
        f.write("property float confidence\nproperty float intensity\n")
        f.write("element face %d\nproperty list uchar int vertex_indices\nend_header\n" % len(faces))
        for v in verts:
            f.write("%.9g %.9g %.9g 1 0.5\n" % (v[0], v[1], v[2]))
        for a, b, c in faces:
            f.write("3 %d %d %d \n" % (a, b, c))
    return len(verts), len(faces)


def asset_dir():
    """Directory with the generated stand-in assets (created on first use, deterministic)."""
    global _ASSET_DIR
    if _ASSET_DIR is None:
        d = os.environ.get("FTB_ASSET_DIR") or os.path.join(tempfile.gettempdir(), "functracer_b200_assets_v1")
        os.makedirs(d, exist_ok=True)
        jobs = {
            "env4.ppm": lambda p: _env_map(p),
            "moon.ppm": lambda p: _moon_map(p),
            "bunny_res4.ply": lambda p: _blob_ply(p, 24, 21),     # 482 verts / 960 faces  (res4-like)
            "bunny_full.ply": lambda p: _blob_ply(p, 264, 133),   # 34 850 verts / 69 696 faces (full-like)
            "bunny_tiny.ply": lambda p: _blob_ply(p, 8, 6),       # 42 verts / 80 faces (unit tests)
        }
        for name, fn in jobs.items():
            p = os.path.join(d, name)
            if not os.path.exists(p):
                tmp = p + ".tmp%d" % os.getpid()
                fn(tmp)
                os.replace(tmp, p)
        _ASSET_DIR = d
    return _ASSET_DIR


# ---------------------------------------------------------------------------------- scene texts
def _t(v):
    return "(" + ",".join(("%g" % x) for x in v) + ")"


def _options(camera, res, spp):
    lines = [camera]
    if spp is not None:
        lines.append("samples %s" % spp)
    if res is not None:
        lines.append("res %d %d" % res)
    return "\n".join(lines) + "\n\n"


def _mat(colour, refl=0, shin=0, rough=None):
    s = "material diffuse %s " % (_t(colour) if not isinstance(colour, (int, float)) else "%g" % colour)
    if rough is not None:
        s += "roughness %g " % rough
    return s + "reflectance %g shineyness %g" % (refl, shin)


def sample(res=None, spp=1, focus=True):
    cam = "camera pos (0,3,-5) lookat (0,0,10) up (0,1,0) fov 60 ratio 1" + (" focus (12,2)" if focus else "")
    objs = [
        ";skybox",
        '(ignoreLight\n    (texture image "env4.ppm"\n        (translate (0,3,-5) (scale 200 sphere))\n    )\n)',
        "(%s \n    (translate (0,1,10) (scale (3,3,3) sphere )) \n)" % _mat((0.1, 0.1, 1.2), 0.2, 30),
        "(%s \n    ((translate (-15,3,40)) . (scale (3,3,3) ) sphere ) \n)" % _mat((0.2, 0.8, 0), 0, 30),
        "(%s \n    (translate (1,3,6)  sphere ) \n)" % _mat((0.8, 0, 0), 0.2, 30),
        "(texture (scale (0.2, 0.2) grid #8cff69 #c882ff)\n    (%s \n        (translate (5,1,7)  sphere ) \n    )\n)" % _mat((0, 0, 0), 0.1, 30),
        "(%s \n    (translate (-3,1,0)  sphere ) \n)" % _mat((0, 0, 0.5), 0, 0),
    ]
    lights = ["softDirectional dir (1,-3,-3) samples 1 scatter 36 colour (0.5,0.5,0.5)",
              "softDirectional dir (-3,-2,3) samples 1 scatter 36 colour (1,1,1)"]
    return _options(cam, res, spp) + "\n\n".join(objs) + "\n\n" + "\n".join(lights) + "\n"


def hollow_sphere(res=None, spp=1):
    cam = "camera pos (5.7,5.7,-5.7) lookat (0,0,4) up (0,1,0) fov 60 ratio 1"
    objs = ["(%s (subtract (scale 11 sphere) (scale 10 sphere )))" % _mat((0.4, 0.4, 0.4), 0, 0)]
    colours = [(1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 1, 1), (0, 0, 1)]
    k = 0
    for row, y in enumerate([2.6, 1.3, 0, -1.3, -2.6]):
        for col, x in enumerate([-2.6, -1.3, 0, 1.3, 2.6]):
            op = "subtract " if k % 2 == 0 else "intersect"
            objs.append("(%s  (translate %s (%s cube (scale 0.65 sphere))))" % (_mat(colours[col], 0.4, 10), _t((x, y, 0)), op))
            k += 1
        objs.append("; row %d" % row)
    lights = ["positional pos (0,0,-8) falloff (1,0.01,0.02) colour (1,1,1)"]
    return _options(cam, res, spp) + "\n".join(objs) + "\n\n" + "\n".join(lights) + "\n"


_HOUSE = """(group
    (%s
        (subtract (scale 4 solidCylinder) (scale (1.5,3,50) sphere))
    )
    (%s
        (translate (0,4,0) (scale (5,3,5) cone))
    )
)""" % (_mat((0.39, 0.58, 0.92)), _mat((1, 0, 0)))

_TREE = """(translate (-8,0,-5)
    (group
        (%s
            (scale (0.5,4,0.5) cylinder)
        )
        (%s
            (group
                (translate (0,7,0) (scale 3 sphere))
                (translate (2,7,-2) (scale 0.5 sphere))
                (translate (2,9,0) (scale 0.5 sphere))
            )
        )
    )
)""" % (_mat((0.65, 0.17, 0.17)), _mat((0, 1, 0)))

_CRATES = """(%s
(translate (9,0,0)
(group
    (translate (-5.5, 0.5, -4.0) cube)
    (translate (-4.8, 1.5, -4.0) (rotate (0,1,0) 10 cube))
    (translate (-4.1, 0.5, -4.3) (rotate (0,1,0) 0 cube))
)))""" % _mat((0.65, 0.17, 0.17))

_GROUND = "(%s\n    plane \n)" % _mat((1, 1, 1))


def house(res=None, spp=1):
    """house.scene with the comment inside the group (house.scene:17, a parse error at HEAD) left out."""
    cam = "camera pos (0,5,-20) lookat (-2,0,0) up (0,1,0) fov 60 ratio 1"
    lights = ["softdirectional dir (2,-1,1) samples 1 scatter 5 colour (0.8,0.8,0.8)",
              "positional pos (0.5,2,3) falloff (1,1,1) colour (1,1,1)"]
    return _options(cam, res, spp) + "; House\n" + _HOUSE + "\n\n; Tree 1\n" + _TREE + "\n\n; Crates\n" + _CRATES + "\n\n" + _GROUND + "\n\n" + "\n".join(lights) + "\n"


def night_house(res=None, spp=1):
    cam = "camera pos (15,11,-20) lookat (-2,0,0) up (0,1,0) fov 60 ratio 1"
    fence = """(%s
    (repeat 8 translate (-0.4,0,-1)
        (translate (-2,0,-5) (scale (0.1,1.5,0.1) solidCylinder))
    )
)""" % _mat((0, 1, 1), 0.5, 0)
    lights = ["softdirectional dir (2,-1,1) samples 1 scatter 5 colour (0.1,0.1,0.1)",
              "positional pos (0.5,2,2) falloff (1,0.01,0.02) colour (0.7,0.7,0.2)",
              "positional pos (-6,2.9,-6) falloff (1,0.01,0.02) colour (0.7,0.7,0.7)"]
    return (_options(cam, res, spp) + "; House\n" + _HOUSE + "\n\n; Fence\n" + fence + "\n\n; Tree 1\n" + _TREE +
            "\n\n; Crates\n" + _CRATES + "\n\n" + _GROUND.replace("\n    plane \n", " plane ") + "\n\n" + "\n".join(lights) + "\n")


def repeat(res=None, spp=1):
    cam = "camera pos (-4,3,-10) lookat (0,0,200) up (0,1,0) fov 60 ratio 1"
    body = "(repeat 5 (translate (-5,0,10)) . (hueshift 1) \n    " + _HOUSE.replace("\n", "\n    ") + ") "
    lights = ["softdirectional dir (2,-1,1) samples 1 scatter 5 colour (0.8,0.8,0.8)",
              "positional pos (0.5,2,3) falloff (1,1,1) colour (1,1,1)"]
    return _options(cam, res, spp) + body + "\n\n" + _GROUND + "\n\n" + "\n".join(lights) + "\n"


def moon(res=(400, 400), spp=1):
    cam = "camera pos (0,0,-100) lookat (0,0,10) up (0,1,0) fov 6 ratio 1 "
    balls = []
    for (x, y), rough in zip([(-3, 3), (3, 3), (-3, -3), (3, -3)], [0, 0.2, 0.4, 0.6]):
        balls.append("    (%s  \n        (translate %s (scale 2 sphere ))  \n    ) " % (_mat(1, 0, 0, rough=rough), _t((x, y, 0))))
    body = '(texture image "moon.ppm" \n(group \n' + "\n \n".join(balls) + "\n)\n)"
    return _options(cam, res, spp) + body + "\ndirectional dir (0.5,0,1) colour 1 \n"


def bunny(res=None, spp=1, depth=0, mesh="bunny_res4.ply"):
    cam = "camera pos (0,2,-2) lookat (0,0,3) up (0,1,0) fov 60 ratio 1 "
    body = '(%s \n    (scale 8 (rotate (0,1,0) 180\n        bspMesh %d "%s"\n    ))\n)' % (_mat(1, 0, 0), depth, mesh)
    return _options(cam, res, spp) + body + "\n\ndirectional dir (-3,-2,3) colour (1,1,1)\n"


# BASELINE.json configs -> (builder, kwargs); jitter seed = 1 + cfg index
CONFIGS = {
    "cfg1-sample": dict(build=sample, res=(640, 480), spp=1, seed=1),
    "cfg2-hollow-sphere": dict(build=hollow_sphere, res=(1920, 1080), spp=4, seed=2),
    "cfg3-house": dict(build=house, res=(1920, 1080), spp=16, seed=3),
    "cfg3-night-house": dict(build=night_house, res=(1920, 1080), spp=16, seed=3),
    "cfg4-bunny": dict(build=bunny, res=(3840, 2160), spp=16, seed=4),
    # the same config with a real BSP (depth is a scene-file argument, bunny.scene:7) and with the full-size mesh
    "cfg4-bunny-d12": dict(build=bunny, res=(3840, 2160), spp=16, seed=4, kw=dict(depth=12)),
    "cfg4-bunny-full-d14": dict(build=bunny, res=(3840, 2160), spp=16, seed=4, kw=dict(depth=14, mesh="bunny_full.ply")),
    "cfg5-repeat": dict(build=repeat, res=(7680, 4320), spp=64, seed=5),
    "cfg5-moon": dict(build=moon, res=(7680, 4320), spp=64, seed=5),
}


def config_text(name, res=None, spp=None, **kw):
    c = CONFIGS[name]
    args = dict(c.get("kw", {}))
    args.update(kw)
    return c["build"](res=res or c["res"], spp=spp or c["spp"], **args)
