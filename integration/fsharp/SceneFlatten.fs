/// (SceneOptions, Scene) -> ftb_scene_desc: the same structural recursion as Scene.intersect (Scene.fs:67-104),
/// emitting rows of flat tables instead of closures.  Needs the data twins described in README.md
/// (Primitive.BspMesh of Geometry * BspData, Texture.Image of Texture * ImageData, Transform.toArray).
module SceneFlatten

open System
open System.Collections.Generic
open System.Runtime.InteropServices
open Native
open Scene
open Light

// ftb_node_kind / ftb_prim_kind (Scene.fs:8-18, 33-53 in declaration order)
let private primKind = function
    | BspMesh _ -> 0 | Circle -> 1 | Square -> 2 | Cube -> 3 | Sphere -> 4
    | Plane -> 5 | Cone -> 6 | SolidCylinder -> 7 | Cylinder -> 8 | Triangle _ -> 9

type Tables () =
    member val nodes = List<FtbNode>()
    member val children = List<int>()
    member val transforms = List<float>()   // 24 doubles per ftb_transform (m2w rows, then w2m rows): a blittable float[] pins, a struct holding arrays would not
    member val materials = List<FtbMaterial>()
    member val textures = List<FtbTexture>()
    member val images = List<byte[] * int * int>()
    member val meshes = List<FtbMesh>()
    member val bspNodes = List<FtbBspNode>()
    member val bspLeaves = List<FtbBspLeaf>()
    member val triangles = List<float>()
    member val lights = List<FtbLight>()

let private rows34 (m: float[,]) = [| for j in 0..2 do for i in 0..3 -> m.[j, i] |]

let private addTriangle (t: Tables) (Triangle.Triangle (a, b, c)) =
    let idx = t.triangles.Count / 9
    for p in [a; b; c] do
        let (Point (x, y, z)) = p
        t.triangles.Add x; t.triangles.Add y; t.triangles.Add z
    idx

// BspMesh.compile's tree (BspMesh.fs:51-65) as flat nodes; link >= 0 = branch, < 0 = ~leaf
let rec private addBsp (t: Tables) = function
    | BspMesh.LeafD tris ->
        let first = t.triangles.Count / 9
        tris |> Array.iter (addTriangle t >> ignore)
        t.bspLeaves.Add (FtbBspLeaf (triFirst = first, triCount = tris.Length))
        ~~~(t.bspLeaves.Count - 1)
    | BspMesh.BranchD (aabb, left, right) ->
        let idx = t.bspNodes.Count
        t.bspNodes.Add (FtbBspNode ())
        let l = addBsp t left
        let r = addBsp t right
        let (Point (x0, y0, z0)), (Point (x1, y1, z1)) = aabb.min, aabb.max
        t.bspNodes.[idx] <- FtbBspNode (minX = x0, minY = y0, minZ = z0, maxX = x1, maxY = y1, maxZ = z1, left = l, right = r)
        idx

let rec private addTexture (t: Tables) = function
    | Image (_, data: ImageTexture.ImageData) ->
        t.images.Add (data.rgb24, data.width, data.height)
        t.textures.Add (FtbTexture (kind = 0, image = t.images.Count - 1)); t.textures.Count - 1
    | Grid (Colour (r1, g1, b1), Colour (r2, g2, b2)) ->
        t.textures.Add (FtbTexture (kind = 1, p0 = r1, p1 = g1, p2 = b1, p3 = r2, p4 = g2, p5 = b2)); t.textures.Count - 1
    | TextureFunction (inner, Scale (x, y)) ->
        let i = addTexture t inner
        t.textures.Add (FtbTexture (kind = 2, inner = i, p0 = x, p1 = y)); t.textures.Count - 1
    | TextureFunction (inner, Rotate angle) ->
        let i = addTexture t inner
        // Texture.rotate builds matrix (rotate unitY angle) per lookup (Texture.fs:18-22): hand over its cos / sin entries
        let m = Transform.matrix (Transform.rotate Vector.unitY angle) |> Transform.toArray
        t.textures.Add (FtbTexture (kind = 3, inner = i, p0 = float angle, p1 = m.[0, 0], p2 = m.[0, 2])); t.textures.Count - 1

let rec private addNode (t: Tables) (g: SceneGraph) : int =
    let emit n = t.nodes.Add n; t.nodes.Count - 1
    match g with
    | Primitive p ->
        let payload =
            match p with
            | BspMesh (_, data) -> t.meshes.Add (FtbMesh (root = addBsp t data)); t.meshes.Count - 1
            | Triangle tri -> addTriangle t tri
            | _ -> 0
        emit (FtbNode (0, primKind p, payload))
    | SceneFunction (f, child) ->
        let c = addNode t child
        match f with
        | Transform tr ->
            t.transforms.AddRange (rows34 (Transform.matrix tr |> Transform.toArray))
            t.transforms.AddRange (rows34 (Transform.matrix (Transform.inverse tr) |> Transform.toArray))
            emit (FtbNode (1, t.transforms.Count / 24 - 1, c))
        | Material m ->
            let (Colour (r, g, b)) = m.colour
            t.materials.Add (FtbMaterial (r = r, g = g, b = b, roughness = m.roughness, reflectance = m.reflectance,
                                          shineyness = m.shineyness, applyLighting = (if m.applyLighting then 1 else 0)))
            emit (FtbNode (2, t.materials.Count - 1, c))
        | Texture tex -> emit (FtbNode (3, addTexture t tex, c))
        | HueShift _ -> emit (FtbNode (4, 0, c))
        | IgnoreLight -> emit (FtbNode (5, 0, c))
    | Group nodes ->
        let kids = nodes |> List.map (addNode t)      // children first: their own groups use the table too
        let first = t.children.Count
        t.children.AddRange kids
        emit (FtbNode (6, first, kids.Length))
    | Union (a, b) -> let x, y = addNode t a, addNode t b in emit (FtbNode (7, x, y))
    | Intersect (a, b) -> let x, y = addNode t a, addNode t b in emit (FtbNode (8, x, y))
    | Subtract (a, b) -> let x, y = addNode t a, addNode t b in emit (FtbNode (9, x, y))
    | Exclude (a, b) -> let x, y = addNode t a, addNode t b in emit (FtbNode (10, x, y))

let private addLight (t: Tables) (Light (cfg, Colour (r, g, b))) =
    let l =
        match cfg with
        | Directional (Vector (x, y, z)) -> FtbLight (kind = 0, vx = x, vy = y, vz = z)
        | SoftDirectional (Vector (x, y, z), samples, scatter) -> FtbLight (kind = 1, samples = samples, vx = x, vy = y, vz = z, scatterRad = float scatter)
        | Point (Point (x, y, z), Falloff (c, lin, q)) -> FtbLight (kind = 2, vx = x, vy = y, vz = z, falloffC = c, falloffL = lin, falloffQ = q)
    let mutable l = l
    l.cr <- r; l.cg <- g; l.cb <- b
    t.lights.Add l

/// Pins every table for the duration of ftb_scene_create (which copies) and ftb_render.
type Flattened (scene: Scene, camera: Image.Camera) =
    let t = Tables ()
    let root = addNode t scene.objects
    do scene.lights |> List.iter (addLight t)
    let pins = List<GCHandle> ()
    let pin (a: Array) = (let h = GCHandle.Alloc (a, GCHandleType.Pinned) in pins.Add h; h.AddrOfPinnedObject ())
    let imageRows = t.images |> Seq.map (fun (bytes, w, h) -> FtbImage (rgb24 = pin bytes, width = w, height = h)) |> Seq.toArray
    member val desc =
        FtbSceneDesc (root = root,
                      nNodes = t.nodes.Count, nodes = pin (t.nodes.ToArray ()),
                      nChildren = t.children.Count, children = pin (t.children.ToArray ()),
                      nTransforms = t.transforms.Count / 24, transforms = pin (t.transforms.ToArray ()),
                      nMaterials = t.materials.Count, materials = pin (t.materials.ToArray ()),
                      nTextures = t.textures.Count, textures = pin (t.textures.ToArray ()),
                      nImages = imageRows.Length, images = pin imageRows,
                      nMeshes = t.meshes.Count, meshes = pin (t.meshes.ToArray ()),
                      nBspNodes = t.bspNodes.Count, bspNodes = pin (t.bspNodes.ToArray ()),
                      nBspLeaves = t.bspLeaves.Count, bspLeaves = pin (t.bspLeaves.ToArray ()),
                      nTriangles = t.triangles.Count / 9, triangles = pin (t.triangles.ToArray ()),
                      nLights = t.lights.Count, lights = pin (t.lights.ToArray ())) with get, set
    member val camera =
        let (Point (ox, oy, oz)), (Point (lx, ly, lz)), (Vector (ux, uy, uz)) = camera.o, camera.lookAt, camera.up
        let f = camera.focus
        FtbCamera (ox = ox, oy = oy, oz = oz, lx = lx, ly = ly, lz = lz, ux = ux, uy = uy, uz = uz,
                   fovYRad = float camera.fovY, aspect = camera.aspectRatio,
                   hasFocus = (if f.IsSome then 1 else 0),
                   focalLength = (match f with Some x -> x.focalLength | None -> 0.0),
                   apertureRad = (match f with Some x -> float x.apetureAngularSize | None -> 0.0)) with get, set
    interface IDisposable with
        member __.Dispose () = for h in pins do h.Free ()
