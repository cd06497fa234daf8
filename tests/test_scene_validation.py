"""ftb_scene_create validates the flattened SceneGraph before it touches a device: a malformed description is refused with
FTB_ERR_BAD_SCENE (or FTB_ERR_UNSUPPORTED) and a message, never dereferenced.  On a box without a GPU a well-formed
scene gets as far as FTB_ERR_NO_DEVICE, which is how these tests tell "accepted" from "refused" without any compute.
The reference has no such boundary (its SceneGraph is a typed DU, SceneParser.fs builds it); the C ABI does."""
import contextlib

import pytest

from functracer_b200 import abi, api, frontend, scenes


def _parsed(text=None):
    return frontend.ParsedScene(text or scenes.house(res=(16, 16), spp=1), scenes.asset_dir())


def _status(parsed):
    try:
        api.Scene(parsed).close()
        return 0, ""
    except api.FtbError as e:
        return e.status, str(e)


ACCEPTED = (0, abi.ERR_NO_DEVICE)  # with / without a GPU


@contextlib.contextmanager
def _patched(obj, field, value):
    old = getattr(obj, field)
    setattr(obj, field, value)
    try:
        yield
    finally:
        setattr(obj, field, old)


def _first(d, kind):
    for i in range(d.n_nodes):
        if d.nodes[i].kind == kind:
            return d.nodes[i]
    raise AssertionError("no node of kind %d" % kind)


def test_well_formed_scenes_are_accepted():
    for text in (scenes.house(res=(16, 16), spp=1), scenes.hollow_sphere(res=(16, 16), spp=1), scenes.sample(res=(16, 16), spp=1)):
        assert _status(_parsed(text))[0] in ACCEPTED


@pytest.mark.parametrize("field,value,needle", [
    ("root", -1, "node"), ("root", 1 << 20, "node"), ("n_nodes", 0, "no nodes"),
])
def test_bad_roots(field, value, needle):
    p = _parsed()
    with _patched(p.desc, field, value):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and needle in msg, (st, msg)
    assert _status(p)[0] in ACCEPTED  # restored: the description itself was not damaged


def test_bad_indices_inside_nodes():
    p = _parsed()
    d = p.desc
    cases = [
        (_first(d, abi.NODE_TRANSFORM), "a", d.n_transforms, "transform"),
        (_first(d, abi.NODE_TRANSFORM), "a", -1, "transform"),
        (_first(d, abi.NODE_MATERIAL), "a", d.n_materials + 7, "material"),
        (_first(d, abi.NODE_GROUP), "b", d.n_children + 1, "group"),
        (_first(d, abi.NODE_GROUP), "a", -3, "group"),
        (_first(d, abi.NODE_PRIMITIVE), "a", 99, "primitive"),
    ]
    for node, field, value, needle in cases:
        with _patched(node, field, value):
            st, msg = _status(p)
        assert st == abi.ERR_BAD_SCENE and needle in msg, (field, value, st, msg)
    assert _status(p)[0] in ACCEPTED


def test_bad_child_link_and_cycles():
    p = _parsed()
    d = p.desc
    old = d.children[0]
    d.children[0] = d.n_nodes  # out of range
    try:
        st, msg = _status(p)
    finally:
        d.children[0] = old
    assert st == abi.ERR_BAD_SCENE and "child" in msg
    # a transform node that is its own child: the walk must stop, not recurse for ever
    t = _first(d, abi.NODE_TRANSFORM)
    idx = [i for i in range(d.n_nodes) if d.nodes[i].kind == abi.NODE_TRANSFORM][0]
    with _patched(t, "b", idx):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and ("deep" in msg or "cyclic" in msg), (st, msg)


def test_bad_lights_and_textures():
    p = _parsed(scenes.sample(res=(16, 16), spp=1))
    d = p.desc
    with _patched(d.lights[0], "kind", 17):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and "light" in msg
    with _patched(d.lights[0], "samples", -2):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and "sample" in msg
    tex = _first(d, abi.NODE_TEXTURE)
    with _patched(tex, "a", d.n_textures):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and "texture" in msg
    assert _status(p)[0] in ACCEPTED


def test_mesh_as_csg_operand_is_refused_as_unsupported():
    text = scenes.bunny(res=(16, 16), spp=1, depth=2, mesh="bunny_tiny.ply")
    p = _parsed(text)
    assert _status(p)[0] in ACCEPTED
    d = p.desc
    # wrap: turn the node above the mesh primitive into a CSG union of the mesh with itself
    prim = [i for i in range(d.n_nodes) if d.nodes[i].kind == abi.NODE_PRIMITIVE and d.nodes[i].a == abi.PRIM_BSPMESH]
    assert prim
    parents = [i for i in range(d.n_nodes) if d.nodes[i].kind in (abi.NODE_TRANSFORM, abi.NODE_MATERIAL) and d.nodes[i].b == prim[0]]
    assert parents
    n = d.nodes[parents[0]]
    old = (n.kind, n.a, n.b)
    n.kind, n.a, n.b = abi.NODE_UNION, prim[0], prim[0]
    try:
        st, msg = _status(p)
    finally:
        n.kind, n.a, n.b = old
    assert st == abi.ERR_UNSUPPORTED and "CSG" in msg, (st, msg)


def test_bad_mesh_index_structures():
    p = _parsed(scenes.bunny(res=(16, 16), spp=1, depth=3, mesh="bunny_tiny.ply"))
    d = p.desc
    assert d.n_bsp_leaves > 0 and d.n_triangles > 0
    with _patched(d.bsp_leaves[0], "tri_count", d.n_triangles + 1):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and "triangle range" in msg, (st, msg)
    with _patched(d.bsp_leaves[0], "tri_first", -1):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE, (st, msg)
    if d.n_bsp_nodes > 0:
        with _patched(d.bsp_nodes[0], "left", d.n_bsp_nodes + 5):  # links >= 0 are nodes, < 0 leaves (functracer_b200.h)
            st, msg = _status(p)
        assert st == abi.ERR_BAD_SCENE and "BSP" in msg, (st, msg)
        with _patched(d.bsp_nodes[0], "left", 0):  # a node that is its own child
            st, msg = _status(p)
        assert st == abi.ERR_BAD_SCENE and ("deep" in msg or "cyclic" in msg), (st, msg)
    mesh_prim = [d.nodes[i] for i in range(d.n_nodes) if d.nodes[i].kind == abi.NODE_PRIMITIVE and d.nodes[i].a == abi.PRIM_BSPMESH][0]
    with _patched(mesh_prim, "b", d.n_meshes):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and "mesh" in msg, (st, msg)
    assert _status(p)[0] in ACCEPTED


def test_bad_image_texture():
    p = _parsed(scenes.sample(res=(16, 16), spp=1))
    d = p.desc
    img_tex = [d.textures[i] for i in range(d.n_textures) if d.textures[i].kind == abi.TEX_IMAGE]
    assert img_tex
    with _patched(img_tex[0], "image", d.n_images):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and "image" in msg, (st, msg)
    with _patched(d.images[img_tex[0].image], "width", 0):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and "image" in msg, (st, msg)
    with _patched(img_tex[0], "kind", 9):
        st, msg = _status(p)
    assert st == abi.ERR_BAD_SCENE and "texture kind" in msg, (st, msg)
    assert _status(p)[0] in ACCEPTED
