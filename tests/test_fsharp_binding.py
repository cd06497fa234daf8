"""The F# side of the boundary (integration/fsharp/*.fs) cannot be compiled in this image (no .NET), so it is checked
mechanically against the header instead: every [<Struct; StructLayout(LayoutKind.Sequential)>] type of Native.fs is laid out
with the C rules the CLR applies to blittable sequential structs (int = 4, float = double = 8, nativeint / int64 / uint64 = 8,
natural alignment) and compared field by field - offset and size - with the ctypes mirror that tests/test_abi.py pins to
include/functracer_b200.h through gcc; every DllImport must name an export of the header with the same number of arguments;
the record-style constructions in RunTracerNative.fs / SceneFlatten.fs may only name fields the structs have.  CPU only."""
import ctypes as C
import os
import re

from functracer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FS = os.path.join(ROOT, "integration", "fsharp")
HEADER = os.path.join(ROOT, "include", "functracer_b200.h")

FS_TYPES = {"int": 4, "int32": 4, "uint32": 4, "float": 8, "double": 8, "nativeint": 8, "int64": 8, "uint64": 8}
MIRROR = {
    "FtbNode": abi.Node, "FtbMaterial": abi.Material, "FtbTexture": abi.Texture, "FtbImage": abi.Image,
    "FtbBspNode": abi.BspNode, "FtbBspLeaf": abi.BspLeaf, "FtbMesh": abi.Mesh, "FtbLight": abi.Light,
    "FtbSceneDesc": abi.SceneDesc, "FtbCamera": abi.Camera, "FtbRenderParams": abi.RenderParams,
}


def fs_structs():
    """{type name: [(field, F# type)]} of Native.fs, in declaration order."""
    text = open(os.path.join(FS, "Native.fs")).read()
    out = {}
    for m in re.finditer(r"\[<Struct; StructLayout\(LayoutKind\.Sequential\)>\]\s*type (\w+) =\n((?:[ \t]+.*\n)+)", text):
        fields = re.findall(r"^\s+val mutable (\w+)\s*:\s*(\w+)", m.group(2), flags=re.M)
        out[m.group(1)] = fields
    return out


def sequential_layout(fields):
    """C layout of scalar fields: [(offset, size)], sizeof."""
    off, align, out = 0, 1, []
    for _, t in fields:
        size = FS_TYPES[t]
        off = (off + size - 1) // size * size
        out.append((off, size))
        off += size
        align = max(align, size)
    return out, (off + align - 1) // align * align


def ctypes_scalars(cls):
    """The scalar slots of a ctypes struct in memory order: arrays expanded, pointers as 8-byte slots."""
    out = []
    for name, typ in cls._fields_:
        base = getattr(cls, name).offset
        if issubclass(typ, C.Array):
            n, elem = typ._length_, typ._type_
            for k in range(n):
                out.append((base + k * C.sizeof(elem), C.sizeof(elem)))
        else:
            out.append((base, C.sizeof(typ)))
    return out


def test_every_fsharp_struct_has_the_headers_layout():
    structs = fs_structs()
    assert set(structs) == set(MIRROR), "Native.fs declares %s" % sorted(structs)
    for name, cls in MIRROR.items():
        layout, size = sequential_layout(structs[name])
        assert size == C.sizeof(cls), (name, size, C.sizeof(cls))
        assert layout == ctypes_scalars(cls), name


def header_prototypes():
    """{export: number of parameters} from the header."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    text = re.sub(r"//.*", "", text)
    protos = {}
    for m in re.finditer(r"\b(ftb_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", text):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_every_dllimport_is_an_export_with_the_same_arity():
    text = open(os.path.join(FS, "Native.fs")).read()
    protos = header_prototypes()
    externs = re.findall(r"\[<DllImport\(Lib, CallingConvention = CallingConvention\.Cdecl\)>\]\s*extern \w+ (ftb_\w+)\(([^)]*)\)", text)
    assert len(externs) >= 7
    for name, args in externs:
        assert name in protos and name in abi.EXPORTS, name
        n = 0 if not args.strip() else args.count(",") + 1
        assert n == protos[name], (name, n, protos[name])
    # the calls the shim cannot do without
    assert {"ftb_scene_create", "ftb_scene_destroy", "ftb_render", "ftb_last_error", "ftb_abi_version"} <= {n for n, _ in externs}


def test_named_field_constructions_only_use_declared_fields():
    structs = {k: {f for f, _ in v} for k, v in fs_structs().items()}
    for fn in ("RunTracerNative.fs", "SceneFlatten.fs"):
        text = open(os.path.join(FS, fn)).read()
        for m in re.finditer(r"\b(Ftb\w+)\s*\(([^()]*(?:\([^()]*\)[^()]*)*)\)", text):
            if m.group(1) not in structs:
                continue
            named = re.findall(r"(?:^|,)\s*(\w+)\s*=", m.group(2))
            for f in named:
                assert f in structs[m.group(1)], (fn, m.group(1), f)


def test_shim_targets_this_abi_and_the_rgba8_path():
    native = open(os.path.join(FS, "Native.fs")).read()
    m = re.search(r"let AbiVersion = (\d+)", native)
    assert m and int(m.group(1)) == abi.ABI_VERSION
    run = open(os.path.join(FS, "RunTracerNative.fs")).read()
    assert re.search(r"outFormat = %d\b" % abi.OUT_RGBA8, run)       # Image.write's quantisation on the device (what bench.py's e2e measures)
    assert "checkAbi ()" in run                                       # refuses a library of another ABI before passing structs
    assert "recursionLimit = 8" in run                                # Shading.fs:142
    assert "GCHandleType.Pinned" in run                               # the caller's buffer is ordinary pageable memory


def test_flattener_emits_the_headers_enum_values():
    """SceneFlatten.fs writes node / primitive / texture / light kinds as literals; they must be the header's enum values
    (mirrored in abi.py, which tests/test_abi.py::test_enums_match_header pins to the header)."""
    text = open(os.path.join(FS, "SceneFlatten.fs")).read()
    # primKind: `| Name [_] -> k`
    body = text[text.index("let private primKind = function"):text.index("type Tables")]
    prim = {n.lower(): int(k) for n, k in re.findall(r"\|\s*(\w+)(?:\s+_)?\s*->\s*(\d+)", body)}
    assert prim == {n.lower(): i for i, n in enumerate(abi.PRIM_NAMES)}
    # node kinds: the first argument of the FtbNode constructed under each SceneGraph case
    node = text[text.index("let rec private addNode"):text.index("let private addLight")]
    want = {"Primitive": abi.NODE_PRIMITIVE, "Transform": abi.NODE_TRANSFORM, "Material": abi.NODE_MATERIAL, "Texture": abi.NODE_TEXTURE,
            "HueShift": abi.NODE_HUESHIFT, "IgnoreLight": abi.NODE_IGNORELIGHT, "Group": abi.NODE_GROUP, "Union": abi.NODE_UNION,
            "Intersect": abi.NODE_INTERSECT, "Subtract": abi.NODE_SUBTRACT, "Exclude": abi.NODE_EXCLUDE}
    for case, kind in want.items():
        at = re.search(r"\|\s*%s\b" % case, node)  # the case's arm, then the first node it emits
        m = at and re.search(r"FtbNode \((\d+),", node[at.end():])
        assert m and int(m.group(1)) == kind, (case, m and m.group(1))
    tex = text[text.index("let rec private addTexture"):text.index("let rec private addNode")]
    for case, kind in (("Image", abi.TEX_IMAGE), ("Grid", abi.TEX_GRID), ("Scale", abi.TEX_SCALE), ("Rotate", abi.TEX_ROTATE)):
        at = re.search(r"\|[^\n]*\b%s\b[^\n]*->" % case, tex)
        m = at and re.search(r"FtbTexture \(kind = (\d+)", tex[at.end():])
        assert m and int(m.group(1)) == kind, case
    light = text[text.index("let private addLight"):text.index("type Flattened")]
    for case, kind in (("Directional", abi.LIGHT_DIRECTIONAL), ("SoftDirectional", abi.LIGHT_SOFT_DIRECTIONAL), ("Point", abi.LIGHT_POINT)):
        at = re.search(r"\|\s*%s \(" % case, light)
        m = at and re.search(r"FtbLight \(kind = (\d+)", light[at.end():])
        assert m and int(m.group(1)) == kind, case
