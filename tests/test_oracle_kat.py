"""Pins the CPU oracle: every known-answer fact the reference's own tests hold for the hot path
(FuncTracer.Tests/Geometry/BoundingBox.fs:11-27, Sphere.fs:18-21) plus the hand-derived vectors of
SURVEY.md Appendix D (derived from the reference source, cited per test).  CPU only."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from functracer_b200 import abi
from oracle import ftb_oracle as orc
from util import one_object_scene, parse

finite = st.floats(min_value=-1e3, max_value=1e3, allow_nan=False, allow_infinity=False)


# ---- reference tests ---------------------------------------------------------------------------
def test_ref_aabb_hit():  # BoundingBox.fs (tests):11-18
    assert orc.aabb_intersects([-.5, -.5, -.5], [.5, .5, .5], [-10, -10, -10], [1, 1, 1]) is True


def test_ref_aabb_miss():  # BoundingBox.fs (tests):20-27
    assert orc.aabb_intersects([-.5, -.5, -.5], [.5, .5, .5], [0, 0, -10], [0.1, 10, 0.1]) is False


@settings(max_examples=300, deadline=None)
@given(finite, finite, finite, finite, finite, finite)
def test_ref_sphere_zero_or_two_hits(ox, oy, oz, dx, dy, dz):  # Sphere.fs (tests):18-21
    sc = parse(one_object_scene("sphere"))
    hits = orc.node_hits(sc, [ox, oy, oz], [dx, dy, dz])
    assert len(hits) in (0, 2)


@settings(max_examples=200, deadline=None)
@given(finite, finite, finite, finite, finite, finite)
def test_ref_sphere_hits_on_surface(ox, oy, oz, dx, dy, dz):  # Sphere.fs (tests):23-30 (skipped upstream for d = 0)
    if abs(dx) + abs(dy) + abs(dz) < 1e-3:
        return
    sc = parse(one_object_scene("sphere"))
    for h in orc.node_hits(sc, [ox, oy, oz], [dx, dy, dz]):
        assert abs(np.linalg.norm(h["p"]) - 1.0) < 1e-5 * max(1.0, abs(h["t"]) * np.linalg.norm([dx, dy, dz]) + np.linalg.norm([ox, oy, oz])) + 1e-5


# ---- Appendix D ----------------------------------------------------------------------------------
def test_d1_quadratic_far_root_first():  # Math.fs:4-10
    assert orc.quadratic(1, 0, -1) == [1.0, -1.0]
    assert orc.quadratic(1, 0, 1) == []


def test_d2_unit_sphere():  # Sphere.fs:6-21
    sc = parse(one_object_scene("sphere"))
    h = orc.node_hits(sc, [0, 0, -5], [0, 0, 1])
    assert [x["t"] for x in h] == [6.0, 4.0]
    assert h[0]["p"] == (0, 0, 1) and h[0]["n"] == (0, 0, 1)
    assert h[1]["p"] == (0, 0, -1) and h[1]["n"] == (0, 0, -1)
    assert h[0]["uv"] == pytest.approx((0.75, 0.5))
    assert h[1]["uv"] == pytest.approx((0.25, 0.5))


def test_d3_scaled_sphere_t_invariant():  # Transform.fs:85-86
    sc = parse(one_object_scene("(scale 2 sphere)"))
    h = orc.node_hits(sc, [0, 0, -5], [0, 0, 1])
    assert [x["t"] for x in h] == [7.0, 3.0]
    assert h[1]["p"] == (0, 0, -2) and h[1]["n"] == (0, 0, -1)


def test_d4_plane_and_parallel_quirk():  # Plane.fs:9-20
    sc = parse(one_object_scene("plane"))
    h = orc.node_hits(sc, [0, 1, 0], [0, -1, 0])
    assert len(h) == 1 and h[0]["t"] == 1.0 and h[0]["p"] == (0, 0, 0) and h[0]["n"] == (0, 1, 0) and h[0]["uv"] == (0, 0)
    h = orc.node_hits(sc, [0, 1, 0], [1, 0, 0])  # parallel, above: one hit at t = 0, p = o
    assert len(h) == 1 and h[0]["t"] == 0.0 and h[0]["p"] == (0, 1, 0)
    assert orc.node_hits(sc, [0, -1, 0], [1, 0, 0]) == []  # parallel, below: none


def test_d5_cube_face_order_and_noise():  # Cube.fs:9-25
    sc = parse(one_object_scene("cube"))
    h = orc.node_hits(sc, [0, 0, -5], [0, 0, 1])
    assert [x["sub"] for x in h] == [4, 5]  # front then back
    assert [x["t"] for x in h] == pytest.approx([4.5, 5.5], abs=1e-12)
    assert h[0]["n"] == pytest.approx((0, 0, -1), abs=1e-15)
    assert h[1]["n"] == pytest.approx((0, 0, 1), abs=1e-15)
    assert h[0]["n"][1] != 0.0  # cos(90 deg) = 6.1e-17 noise is real in double


def test_d6_csg_rule_tables():  # Csg.fs:19-94
    exp = {
        "subtract": [(3, -1), (4, 1), (6, -1), (7, 1)],
        "intersect": [(4, -1), (6, 1)],
        "union": [(3, -1), (7, 1)],
        "exclude": [(3, -1), (4, 1), (6, -1), (7, 1)],
    }
    for op, want in exp.items():
        sc = parse(one_object_scene("(%s (scale 2 sphere) sphere)" % op))
        h = orc.node_hits(sc, [0, 0, -5], [0, 0, 1])
        assert [(x["t"], x["n"][2]) for x in h] == [(float(t), float(nz)) for t, nz in want], op


def test_csg_keeps_negative_t_and_nests():  # Csg.fs:76-80
    sc = parse(one_object_scene("(subtract (union (translate (0,0,-1) sphere) (translate (0,0,1) sphere)) (scale 0.5 sphere))"))
    h = orc.node_hits(sc, [0, 0, 0], [0, 0, 1])
    assert [x["t"] for x in h] == pytest.approx([-2, -0.5, 0.5, 2])


def test_d9_attenuate():  # Light.fs:16-17
    assert orc.attenuate([1, 0.01, 0.02], 8.0) == 1.0 / (1.0 + 8.0 * (0.01 + 8.0 * 0.02)) == 0.42372881355932196


def test_d10_lambert_unclamped():  # Shading.fs:65-70
    assert orc.lambert([0, 1, 0], [0, -1, 0], [1, 1, 1], [.5, .5, .5]) == (.5, .5, .5)
    assert orc.lambert([0, 1, 0], [0, 1, 0], [1, 1, 1], [.5, .5, .5]) == (-.5, -.5, -.5)


def test_d11_specular_negative_base_integral_exponent():  # Shading.fs:78-87
    # n = +y, light direction chosen so that v.(-r) = -0.5 exactly: view along -r rotated 120 deg
    n = [0, 1, 0]
    ld = [0, -1, 0]  # reflect n ld = ld - 2(ld.n)n = (0, 1, 0); -r = (0,-1,0)
    view = [math.sqrt(3) / 2, 0.5, 0]  # v.(-r) = -0.5
    got = orc.specular(n, ld, [1, 1, 1], view, 10.0)
    assert got == pytest.approx((2.0 ** -10,) * 3, rel=1e-12)
    assert orc.specular(n, ld, [1, 1, 1], view, 0.0) == (0, 0, 0)  # shineyness <= 0
    assert math.isnan(orc.specular(n, ld, [1, 1, 1], view, 2.5)[0])  # NaN passes the <= test and poisons the pixel


def test_d12_grid_texture():  # Texture.fs:8-29
    sc = parse(one_object_scene("(texture grid (1,0,0) (0,0,1) sphere)"))
    c1, c2 = (1, 0, 0), (0, 0, 1)
    assert orc.texture(sc, 0, .25, .25) == c1
    assert orc.texture(sc, 0, .25, .75) == c2
    assert orc.texture(sc, 0, .75, .75) == c1
    assert orc.texture(sc, 0, .75, .25) == c2
    assert orc.texture(sc, 0, .5, .25) == c2
    assert orc.texture(sc, 0, -0.25, 1.25) == c2


def test_texture_functions_outermost_acts_first():  # Scene.fs:68-74, Texture.fs:14-22
    sc = parse(one_object_scene("(texture (scale (0.5, 0.5) grid (1,0,0) (0,0,1)) sphere)"))
    # uv (.15,.15) -> (.3,.3) -> c1 ; uv (.3,.15) -> (.6,.3) -> c2
    top = sc.desc.n_textures - 1
    assert orc.texture(sc, top, .15, .15) == (1, 0, 0)
    assert orc.texture(sc, top, .3, .15) == (0, 0, 1)
    sc = parse(one_object_scene("(texture (rotate 90 grid (1,0,0) (0,0,1)) sphere)"))
    top = sc.desc.n_textures - 1
    # rotate: (u,v) -> (c u + s v, -s u + c v) = (v, -u): (.25,.25) -> (.25,-.25 ~ .75) -> c2
    assert orc.texture(sc, top, .25, .25) == (0, 0, 1)


def test_d13_to_byte():  # Image.fs:36, Math.fs:12-16
    assert [orc.to_byte(x) for x in (0.5, 0.999, 1.0, 1.7, -0.2)] == [127, 254, 255, 255, 0]


def test_d14_hue_shift():  # CommonTypes.fs:90
    assert orc.hue_shift([1, 2, 3]) == (3, 1, 2)


def test_d8_image_plane():  # Image.fs:67-89
    cam = abi.Camera()
    cam.o[:] = [0, 0, 0]
    cam.look_at[:] = [0, 0, 1]
    cam.up[:] = [0, 1, 0]
    cam.fov_y_rad = 60.0 * 1.0 * (math.pi / 180.0)
    cam.aspect_ratio = 1.0
    h = 2 * math.tan(cam.fov_y_rad / 2)
    ph, pw = h / 639, h / 479  # the swapped axes of Image.fs:71-72
    assert ph == pytest.approx(0.0018070430960551666, rel=1e-12)
    assert pw == pytest.approx(0.002410648305593427, rel=1e-12)
    r = orc.primary_ray(cam, 640, 480, 0, 0)
    assert r[3:] == pytest.approx([-h / 2 + pw / 2, h / 2 - ph / 2, 1.0], rel=1e-12)
    r = orc.primary_ray(cam, 640, 480, 639, 479)
    assert r[3] == pytest.approx(0.964259, abs=1e-5) and r[4] == pytest.approx(-0.289127, abs=1e-5)
    r = orc.primary_ray(cam, 640, 480, 10, 20, 0.5, -0.25)
    assert r[3] == pytest.approx(-h / 2 + pw / 2 + 10 * pw + 0.5 * pw, rel=1e-12)
    assert r[4] == pytest.approx(h / 2 - ph / 2 - 20 * ph - 0.25 * ph, rel=1e-12)


# ---- primitives beyond Appendix D (derived from the cited source) -----------------------------
def test_cylinder_and_cone_face_the_ray():  # Cylinder.fs:8-20, Cone.fs:7-28
    sc = parse(one_object_scene("cylinder"))
    h = orc.node_hits(sc, [0, .5, -5], [0, 0, 1])
    assert [x["t"] for x in h] == [6.0, 4.0]
    assert h[0]["n"] == (0, 0, -1) and h[1]["n"] == (0, 0, -1)  # both flipped to face the ray
    assert orc.node_hits(sc, [0, 1.5, -5], [0, 0, 1]) == []  # py filter
    sc = parse(one_object_scene("cone"))
    h = orc.node_hits(sc, [0, .5, -5], [0, 0, 1])  # radius at y=.5 is .5
    assert [x["t"] for x in h] == pytest.approx([5.5, 4.5])
    assert all(x["n"][2] < 0 for x in h)
    assert orc.node_hits(sc, [0, 1.5, -5], [0, 0, 1]) == []  # the mirrored nappe is filtered out


def test_solid_cylinder_parts_in_order():  # Cylinder.fs:25-29
    sc = parse(one_object_scene("solidCylinder"))
    h = orc.node_hits(sc, [0.2, 5, 0.1], [0, -1, 0])
    assert [(x["sub"], x["t"]) for x in h] == [(0, 4.0), (1, pytest.approx(5.0))]
    assert h[0]["n"] == (0, 1, 0)
    assert h[1]["n"] == pytest.approx((0, -1, 0), abs=1e-15)
    h = orc.node_hits(sc, [-5, .5, 0], [1, 0, 0])
    assert [x["sub"] for x in h] == [2, 2]


def test_surface_ops_outermost_material_wins():  # Ray.fs:47-59, Scene.fs:75-80, SceneParser.fs:99-105
    sc = parse(one_object_scene("(material diffuse (1,0,0) reflectance 0.5 shineyness 3 (ignoreLight (material diffuse (0,1,0) reflectance 0 shineyness 0 sphere)))"))
    h = orc.node_hits(sc, [0, 0, -5], [0, 0, 1])[0]
    assert h["colour"] == (1, 0, 0) and h["reflectance"] == 0.5 and h["apply_lighting"] is True
    sc = parse(one_object_scene("(ignoreLight (hueShift 1 (material diffuse (1,2,3) reflectance 0 shineyness 0 sphere)))"))
    h = orc.node_hits(sc, [0, 0, -5], [0, 0, 1])[0]
    assert h["colour"] == (3, 1, 2) and h["apply_lighting"] is False


def test_triangle_moller_trumbore(assets):  # Triangle.fs:43-66
    sc = parse(one_object_scene('mesh "bunny_tiny.ply"'), assets)
    assert sc.desc.n_triangles == 80
    tri = np.array([sc.desc.triangles[i] for i in range(9)]).reshape(3, 3)
    c = tri.mean(axis=0)
    n = np.cross(tri[1] - tri[0], tri[2] - tri[0])
    n /= np.linalg.norm(n)
    o = c + 0.3 * n
    hits = [x for x in orc.node_hits(sc, o, -2 * n) if x["prim"] == 0]
    assert len(hits) == 1
    assert hits[0]["t"] == pytest.approx(0.15)  # t is in units of |d| = 2
    assert hits[0]["p"] == pytest.approx(tuple(c), abs=1e-12)
    assert hits[0]["n"] == pytest.approx(tuple(n), abs=1e-12)  # not flipped toward the ray
    # behind the origin: rejected by t > 1e-7 (so meshes emit no negative-t hits)
    assert [x for x in orc.node_hits(sc, o, 2 * n) if x["prim"] == 0] == []


def test_bsp_traversal_visits_right_then_left(assets):  # BspMesh.fs:67-76
    flat = parse(one_object_scene('bspMesh 0 "bunny_tiny.ply"'), assets)
    deep = parse(one_object_scene('bspMesh 3 "bunny_tiny.ply"'), assets)
    assert flat.desc.n_bsp_nodes == 0 and flat.desc.n_bsp_leaves == 1
    assert deep.desc.n_bsp_nodes >= 3
    o, d = [-0.017, 0.11, -1.0], [0.001, 0.002, 1.0]
    a = sorted(x["t"] for x in orc.node_hits(flat, o, d))
    b = sorted(x["t"] for x in orc.node_hits(deep, o, d))
    assert len(a) == 2 and a == pytest.approx(b, rel=1e-9)


def test_jitter_vector_contract():  # Jitter.fs:26-39 on the ftb_rng contract
    v = orc.jitter_vector(7, 123, 0, 1, 0, math.radians(36), [0, 0, 2])
    assert np.linalg.norm(v) == pytest.approx(1.0, abs=1e-15)
    assert math.acos(v[2]) <= math.radians(18) + 1e-12
    assert (v == orc.jitter_vector(7, 123, 0, 1, 0, math.radians(36), [0, 0, 2])).all()
    assert (v != orc.jitter_vector(7, 124, 0, 1, 0, math.radians(36), [0, 0, 2])).any()


def test_shading_end_to_end_point_light():
    """One sphere, one point light straight behind the camera: colour at the centre pixel from first
    principles (Shading.fs:109-139, Light.fs:16-17)."""
    text = one_object_scene("(material diffuse (0.5,0.25,1) reflectance 0 shineyness 0 sphere)",
                            "positional pos (0,0,-5) falloff (1,0.5,0.25) colour (1,1,1)")
    sc = parse(text)
    p = orc.make_params(1, 1, 1, [0.0, 0.0])
    r = orc.shade_rays(sc, [[0, 0, -5, 0, 0, 1]], p)
    dist = 4.0 - 1e-4  # shadow origin p + 1e-4 n, n = (0,0,-1)
    att = 1.0 / (1 + dist * (0.5 + dist * 0.25))
    assert r["rgb"][0] == pytest.approx([0.5 * att, 0.25 * att, 1.0 * att], rel=1e-9)
    assert r["prim"][0] == 0 and r["t"][0] == pytest.approx(4.0 - 1e-4)


def test_reflection_added_once_per_light_and_depth_limit():
    """Two facing mirrors (planes) with two lights: the reflection term is added once per light
    (Shading.fs:89-98, 131-139) and recursion stops after 8 reflection generations (:133)."""
    obj = ("(material diffuse (0.2,0.2,0.2) reflectance 0.5 shineyness 0 (translate (0,-1,0) plane))\n"
           "(material diffuse (0.3,0.3,0.3) reflectance 0.5 shineyness 0 (translate (0,1,0) (rotate (1,0,0) 180 plane)))")
    lights = "directional dir (0,-1,0) colour (1,1,1)\ndirectional dir (0,1,0) colour (0.5,0.5,0.5)"
    sc = parse(one_object_scene(obj, lights))
    p = orc.make_params(1, 1, 1, [0.0, 0.0])
    r = orc.shade_rays(sc, [[0, 0, 0, 0, -1, 0]], p)
    assert r["stats"].reflection_rays == 8
    assert r["stats"].primary_rays == 1
    p0 = orc.make_params(1, 1, 1, [0.0, 0.0], recursion_limit=0)
    r0 = orc.shade_rays(sc, [[0, 0, 0, 0, -1, 0]], p0)
    assert r0["stats"].reflection_rays == 0
    # level colours: floor lit by light 0 only (light 1 comes from below: negative Lambert, unclamped)
    p1 = orc.make_params(1, 1, 1, [0.0, 0.0], recursion_limit=1)
    r1 = orc.shade_rays(sc, [[0, 0, 0, 0, -1, 0]], p1)
    L = 2
    # local(level) = sum over lights of Lambert; weight per bounce = L * reflectance = 1.0
    assert r1["rgb"][0][0] == pytest.approx(r0["rgb"][0][0] + L * 0.5 * orc.shade_rays(sc, [[0, -1, 0, 0, 1, 0]], p0)["rgb"][0][0], rel=1e-12)
