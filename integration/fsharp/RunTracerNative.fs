/// Replacement for Program.runTracer lines 54-64 (image plane, generateRays, depth-of-field jitter, shader,
/// sceneGeometry, shade, blendPixels): one call into libfunctracer_b200.  readScene and Image.write stay as they are.
module RunTracerNative

open System
open System.IO
open System.Runtime.InteropServices
open Image
open Native

/// SceneOptions.samplingStrategy is a closure; the parser must also record what it built it from
/// (SceneParser.fs:300-316: `samples N` -> JitteredSampling.strategy N, `samples corner` -> CornerSampling.strategy).
type SamplingChoice = Jittered of int | Corner

let run (timer: Diagnostics.Stopwatch) (options: Scene.SceneOptions) (sampling: SamplingChoice) (scene: Scene.Scene) (output: Stream) =
    let resH, resV = Image.resH options.resolution, Image.resV options.resolution
    // the sample pattern is drawn on the host exactly as JitteredSampling.generateRays does (Image.fs:101-105)
    let spp, mode, jitter =
        match sampling with
        | Jittered n ->
            let pattern = Jitter.pattern (Random ()) Jitter.circle n
            n, 0, pattern |> List.collect (fun (Jitter.JitterOffset (x, y)) -> [x; y]) |> List.toArray
        | Corner -> 1, 1, [| 0.0; 0.0 |]
    eprintfn "Generated rays: %ims" timer.ElapsedMilliseconds
    checkAbi ()
    use flat = new SceneFlatten.Flattened (scene, options.camera)
    let mutable desc = flat.desc
    let mutable camera = flat.camera
    let mutable handle = 0n
    check (ftb_scene_create (&desc, &handle))
    eprintfn "Geometry created"
    // RGBA8 out: Image.write's own quantisation (Image.fs:36: clamp, * 255.0, truncate) is applied on the device, so the frame
    // that crosses the bus is 4 bytes per pixel instead of 24 and no Colour list is built.  A GC-pinned array is ordinary
    // pageable memory to CUDA; the library downloads the frame band by band behind the rendering of the later bands.
    let rgba : byte[] = Array.zeroCreate (4 * resH * resV)
    let jh = GCHandle.Alloc (jitter, GCHandleType.Pinned)
    let ph = GCHandle.Alloc (rgba, GCHandleType.Pinned)
    try
        let mutable p =
            FtbRenderParams (width = resH, height = resV, spp = spp, sampling = mode, jitterXy = jh.AddrOfPinnedObject (),
                             recursionLimit = 8, precision = 0, seed = uint64 DateTime.Now.Ticks, outFormat = 2,
                             shardIndex = 0, shardCount = 1, nGpus = 0, collectStats = 0, bandIndex = 0, bandCount = 0)
        check (ftb_render (handle, &camera, &p, ph.AddrOfPinnedObject (), 0n, 0n))
    finally
        jh.Free (); ph.Free (); ftb_scene_destroy handle
    eprintfn "Shaded scene %ims" timer.ElapsedMilliseconds
    eprintfn "Writing output %ims" timer.ElapsedMilliseconds
    // Image.write (Image.fs:35-44) minus its per-pixel loop: the bytes are already Rgba32 in row-major order
    use image = SixLabors.ImageSharp.Image.LoadPixelData<SixLabors.ImageSharp.PixelFormats.Rgba32> (rgba, resH, resV)
    image.Save (output, SixLabors.ImageSharp.Formats.Png.PngEncoder ())
    0

/// The same call with the un-quantised frame (outFormat = 0: 3 doubles per pixel = Bitmap.pixels, Image.fs:30) for callers
/// that want to go on through Image.write unchanged; 6x the bytes over the bus (measured on an 8K frame: +8 ms).
let runF64 (options: Scene.SceneOptions) (handle: nativeint) (camera: byref<FtbCamera>) (p: byref<FtbRenderParams>) : Bitmap =
    let resH, resV = Image.resH options.resolution, Image.resV options.resolution
    let pixels : float[] = Array.zeroCreate (3 * resH * resV)
    let ph = GCHandle.Alloc (pixels, GCHandleType.Pinned)
    try
        p.outFormat <- 0
        check (ftb_render (handle, &camera, &p, ph.AddrOfPinnedObject (), 0n, 0n))
    finally
        ph.Free ()
    { resolution = options.resolution
      pixels = List.init (resH * resV) (fun i -> Colour (pixels.[3 * i], pixels.[3 * i + 1], pixels.[3 * i + 2])) }
