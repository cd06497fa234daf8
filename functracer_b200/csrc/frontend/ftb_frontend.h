/* ftb_frontend.h — host-side front end standing in for the parts of FuncTracer that stay in F#.
 *
 * In production the reference's own SceneParser.fs / PlyParser.fs / BspMesh.compile build the
 * scene and a thin F# flattener fills ftb_scene_desc.  No .NET toolchain exists
 * in this image, so this C++ library restates that host side — the `.scene` grammar
 * (FuncTracer/SceneParser.fs:11-366), the ASCII PLY reader (PlyParser.fs:14-70), the BSP build
 * with triangle clipping (BspMesh.fs:30-65, Triangle.fs:8-41) and Transform.matrix
 * (Transform.fs:47-71) — and emits exactly the ftb_scene_desc the F# flattener would.
 * It contains no rendering code and does not depend on CUDA.
 */
#ifndef FTB_FRONTEND_H
#define FTB_FRONTEND_H

#include "../../../include/functracer_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ftbf_scene ftbf_scene;

const char* ftbf_last_error(void);

/* SceneParser.parse (SceneParser.fs:360-366).  asset_dir resolves `mesh` / `bspMesh` /
 * `texture image` file names: the literal path is tried first, then its basename inside
 * asset_dir (images additionally with the extension replaced by .ppm, binary P6 being the one
 * decoded format here; JPEG/PNG decoding stays with ImageSharp in F#).  Returns 0 or a negative
 * ftb_status; on a parse error the message mirrors FParsec's "line/column + expectation". */
int ftbf_parse(const char* text, const char* asset_dir, ftbf_scene** out);
void ftbf_destroy(ftbf_scene* s);

const ftb_scene_desc* ftbf_desc(const ftbf_scene* s);
const ftb_camera* ftbf_camera(const ftbf_scene* s);
/* SceneOptions after folding the option lines over SceneOptions.Default (Scene.fs:61-65). */
void ftbf_options(const ftbf_scene* s, int* width, int* height, int* spp, int* sampling);

/* Jitter.pattern random Jitter.circle spp (Image.fs:101-105, Jitter.fs:15-24) from a seeded
 * generator instead of the unseeded System.Random: 2*spp doubles in the unit disc. */
void ftbf_jitter_pattern(uint64_t seed, int spp, double* xy);

/* Triangle.slice (Triangle.fs:24-41) exposed for the reference's own known-answer tests
 * (FuncTracer.Tests/Geometry/Triangle.Tests.fs).  above/below receive up to 2 triangles each
 * (9 doubles per triangle). */
void ftbf_slice_triangle(const double* plane_p0, const double* plane_n, const double* tri,
                         double* above, int* n_above, double* below, int* n_below);

/* pcolour (SceneParser.fs:69-87) exposed for FuncTracer.Tests/Parser/Colour.fs. */
int ftbf_parse_colour(const char* text, double* rgb);

/* PNG encoder for the CLI (the reference uses ImageSharp, Image.fs:41-44). rgba = W*H*4. */
int ftbf_write_png(const char* path, int width, int height, const uint8_t* rgba);

#ifdef __cplusplus
}
#endif
#endif
