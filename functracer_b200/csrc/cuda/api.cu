// api.cu — the C ABI of include/functracer_b200.h: scene upload, frame set-up, kernel launch,
// tile assembly (blend / quantise), in-process multi-GPU with a P2P gather, statistics.
//
// There is NO CPU fallback in this file: without a CUDA device every entry point that computes
// returns FTB_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/functracer_b200.h"
#include "bvh_build.h"
#include "device_scene.h"
#include "lower.h"

namespace ftb {

// ---- assembly kernels (api.cu owns them; the render kernels live in render_f32/f64.cu) --------------
constexpr int kMaxShards = 64;
struct TilePtrs {
    const void* p[kMaxShards];
};

template <typename R>
__device__ __forceinline__ void loadTilePixel(const TilePtrs& bufs, int shard_count, int tiles_x, int x, int y, R& r, R& g, R& b)
{
    const int tile = (y / FTB_TILE_H) * tiles_x + (x / FTB_TILE_W);
    const int shard = shardOfTile(tile, shard_count), local = tile / shard_count;
    const R* src = static_cast<const R*>(bufs.p[shard]) + 3 * ((long long)local * FTB_TILE_PIXELS + (y % FTB_TILE_H) * FTB_TILE_W + (x % FTB_TILE_W));
    r = src[0]; g = src[1]; b = src[2];
}

// Image.write's toByte (Image.fs:36, Math.fs:12-16): clamp to [0,1], * 255.0, truncate.
template <typename R>
__device__ __forceinline__ unsigned char toByte(R c)
{
    R k = c > R(1) ? R(1) : (c < R(0) ? R(0) : c);  // NaN falls through both tests like Math.clamp
    return (unsigned char)(int)(k * R(255));
}

// Tile-major shard buffers -> row-major frame.  corner = 1: CornerSampling.blendPixels
// (Image.fs:134-144): pixel = average of the sample grid's [TL; TR; BL; BR] corners.
template <typename R>
__global__ void assemble_kernel(TilePtrs bufs, int shard_count, int tiles_x, int W, int y0, int y1, int corner, int out_format, void* out)
{
    const long long n = (long long)W * y1;
    for (long long i = (long long)W * y0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)(i / W);
        R r, g, b;
        if (!corner) {
            loadTilePixel<R>(bufs, shard_count, tiles_x, x, y, r, g, b);
        } else {
            R r1, g1, b1;
            loadTilePixel<R>(bufs, shard_count, tiles_x, x, y, r, g, b);  // Seq.average: Zero + TL + TR + BL + BR, then / 4
            r = R(0) + r; g = R(0) + g; b = R(0) + b;
            loadTilePixel<R>(bufs, shard_count, tiles_x, x + 1, y, r1, g1, b1); r += r1; g += g1; b += b1;
            loadTilePixel<R>(bufs, shard_count, tiles_x, x, y + 1, r1, g1, b1); r += r1; g += g1; b += b1;
            loadTilePixel<R>(bufs, shard_count, tiles_x, x + 1, y + 1, r1, g1, b1); r += r1; g += g1; b += b1;
            r = r / R(4); g = g / R(4); b = b / R(4);
        }
        if (out_format == FTB_OUT_RGB_F64) {
            double* o = static_cast<double*>(out) + 3 * i;
            o[0] = (double)r; o[1] = (double)g; o[2] = (double)b;
        } else if (out_format == FTB_OUT_RGB_F32) {
            float* o = static_cast<float*>(out) + 3 * i;
            o[0] = (float)r; o[1] = (float)g; o[2] = (float)b;
        } else {
            static_cast<uchar4*>(out)[i] = make_uchar4(toByte(r), toByte(g), toByte(b), 255);
        }
    }
}

// ftb_shade_rays: [ray][3] in R -> doubles
template <typename R>
__global__ void widen_kernel(const R* in, double* out, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

}  // namespace ftb

namespace ftb {
static_assert((unsigned)FT_TABLE == kFeatOriginTable, "lower.h and device_scene.h disagree on the table feature bit");
// kernel variants compiled by the Makefile (render_variant.cu), listed through -DFTB_FEAT_LIST_F32 / _F64
#define X(feat) cudaError_t launch_f32_##feat(const DevScene<float>&, const DevFrame<float>&, bool, int, cudaStream_t, int*); \
    cudaError_t launch_wf_f32_##feat(const DevScene<float>&, const DevFrame<float>&, int, int, bool, int, void*, size_t, cudaStream_t, int*);
FTB_FEAT_LIST_F32
#undef X
#define X(feat) cudaError_t launch_f64_##feat(const DevScene<double>&, const DevFrame<double>&, bool, int, cudaStream_t, int*); \
    cudaError_t launch_wf_f64_##feat(const DevScene<double>&, const DevFrame<double>&, int, int, bool, int, void*, size_t, cudaStream_t, int*);
FTB_FEAT_LIST_F64
#undef X
static const Variant<float> kVariantsF32[] = {
#define X(feat) {(unsigned)(feat), UnitCap<float, (unsigned)(feat)>::value, (feat) == FT_ALL, launch_f32_##feat, launch_wf_f32_##feat},
    FTB_FEAT_LIST_F32
#undef X
};
static const Variant<double> kVariantsF64[] = {
#define X(feat) {(unsigned)(feat), UnitCap<double, (unsigned)(feat)>::value, (feat) == FT_ALL, launch_f64_##feat, launch_wf_f64_##feat},
    FTB_FEAT_LIST_F64
#undef X
};
template <typename R> struct VariantTable;
template <> struct VariantTable<float> { static const Variant<float>* begin() { return kVariantsF32; } static int size() { return (int)(sizeof(kVariantsF32) / sizeof(kVariantsF32[0])); } };
template <> struct VariantTable<double> { static const Variant<double>* begin() { return kVariantsF64; } static int size() { return (int)(sizeof(kVariantsF64) / sizeof(kVariantsF64[0])); } };

// the smallest compiled variant that covers `need` (and has the counting kernel if asked for)
template <typename R>
const Variant<R>* pickCover(unsigned need, bool stats)
{
    const Variant<R>* best = nullptr;
    for (int i = 0; i < VariantTable<R>::size(); ++i) {
        const Variant<R>* v = VariantTable<R>::begin() + i;
        if ((v->feat & need) != need || (stats && !v->has_stats)) continue;
        if (!best || __builtin_popcount(v->feat) < __builtin_popcount(best->feat)) best = v;
    }
    return best;
}
// FT_TABLE is an optimisation, not a requirement: a scene that would like the common-origin bound table but whose other
// needs are only covered together with it by the generic kernel runs on the specialised kernel without the table.
template <typename R>
const Variant<R>* pickVariant(unsigned need, bool stats)
{
    const Variant<R>* v = pickCover<R>(need, stats);
    if ((need & FT_TABLE) && (!v || v->feat == FT_ALL)) {
        const Variant<R>* w = pickCover<R>(need & ~(unsigned)FT_TABLE, stats);
        if (w && w->feat != FT_ALL) return w;
    }
    return v;
}
}  // namespace ftb

namespace {

using namespace ftb;

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
int cudaFail(cudaError_t e, const char* what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    (void)cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? FTB_ERR_OOM : FTB_ERR_CUDA;
}
#define CK(expr)                                              \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) return cudaFail(e__, #expr); \
    } while (0)

// A growable device buffer (scratch that survives between calls so steady-state frames do no cudaMalloc).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Control {  // device control block of one launch
    unsigned int tile_counter;
    unsigned int overflow;
    unsigned long long stats[ST_COUNT];
};

template <typename R>
struct SceneStorage {
    DevScene<R> view;
    std::vector<void*> allocs;
    bool ready = false;
    void release() { for (void* p : allocs) cudaFree(p); allocs.clear(); ready = false; }
};

struct PerDevice;
template <typename R> SceneStorage<R>& storageOf(PerDevice* pd);
typedef std::function<int(int)> ChunkDone;  // called after the launches of chunk k have been queued

// Device -> host copies of finished bands.  The destination is whatever the caller owns: usually ordinary pageable memory
// (a GC-pinned .NET array, a std::vector, numpy), sometimes CUDA page-locked memory.  A cudaMemcpyAsync into pageable
// memory does not return before the data is there (the driver stages it through its own page-locked buffers at link
// speed), so the order of work is what matters: every band's render + assembly is queued FIRST, then the bands are
// copied in the order they finish; while the host sits in the copy of band c the GPU is rendering band c + 1.
// Measured (B200, 16 host cores, RGBA8 frame of the 8K moon scene, 132 MB): driver-staged copy +1.3 ms per frame over the
// device-resident time, a library-owned page-locked ring + memcpy on the calling thread +5.7 ms, cudaHostRegister of
// the caller's buffer per call +21 ms; the 1080p hollow-sphere frame: 1.43 / 1.50 / 14.6 ms.  So the driver's path it is.
struct HostCopier {
    std::vector<cudaStream_t> used;  // streams that carry copies this frame (synchronised by finish)
    cudaError_t begin(const void* d_src, void* host_dst, size_t bytes, cudaStream_t stream)
    {
        if (bytes == 0) return cudaSuccess;
        cudaError_t e = cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess && std::find(used.begin(), used.end(), stream) == used.end()) used.push_back(stream);
        return e;
    }
    cudaError_t finish()
    {
        cudaError_t e = cudaSuccess;
        for (cudaStream_t st : used) {
            cudaError_t e2 = cudaStreamSynchronize(st);
            if (e == cudaSuccess) e = e2;
        }
        used.clear();
        return e;
    }
    void release() { used.clear(); }
};

struct PerDevice {
    int device = -1;
    int sm_count = 0;
    SceneStorage<float> f32;
    SceneStorage<double> f64;
    cudaStream_t stream = nullptr;  // owned; used by the host-buffer entry points
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, done = nullptr;
    DevBuf control, jitter, tiles, out, dbg_prim, dbg_sub, dbg_t, rays, order, wavefront;
    std::vector<int> chunk_first;          // first position in the tile order of every chunk (+ end)
    cudaStream_t copy_stream = nullptr;    // assembly / D2H of finished bands while later bands render (highest priority)
    std::vector<cudaEvent_t> band_events;
    HostCopier copier;
    std::vector<unsigned char> order_key;  // what the cached tile order was computed for
    bool order_valid = false;
    bool control_ready = false;      // the control block has been zeroed once
    std::vector<DevBuf> peer_tiles;  // on the gather device: one per remote shard

    PerDevice() = default;
    PerDevice(const PerDevice&) = delete;
    PerDevice& operator=(const PerDevice&) = delete;
    // Releases everything the device holds, also when a create / upload failed half way (the owner switches devices).
    ~PerDevice()
    {
        if (device < 0) return;
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
        if (cudaSetDevice(device) == cudaSuccess) {
            if (stream) cudaStreamSynchronize(stream);
            if (copy_stream) cudaStreamSynchronize(copy_stream);
            f32.release(); f64.release();
            for (DevBuf* b : {&control, &jitter, &tiles, &out, &dbg_prim, &dbg_sub, &dbg_t, &rays, &order, &wavefront}) b->release();
            for (DevBuf& b : peer_tiles) b.release();
            copier.release();
            if (ev0) cudaEventDestroy(ev0);
            if (ev1) cudaEventDestroy(ev1);
            if (done) cudaEventDestroy(done);
            for (cudaEvent_t e : band_events) cudaEventDestroy(e);
            if (stream) cudaStreamDestroy(stream);
            if (copy_stream) cudaStreamDestroy(copy_stream);
        }
        (void)cudaGetLastError();
        if (cur >= 0) cudaSetDevice(cur);
    }
};

template <> SceneStorage<float>& storageOf<float>(PerDevice* pd) { return pd->f32; }
template <> SceneStorage<double>& storageOf<double>(PerDevice* pd) { return pd->f64; }

}  // namespace

struct ftb_scene {
    ftb::Lowered L;
    // host copies of the tables the upload needs
    std::vector<ftb_bsp_node> bsp_nodes;
    std::vector<ftb_bsp_leaf> bsp_leaves;
    std::vector<ftb_mesh> meshes;
    std::vector<double> triangles;
    std::vector<ftb_light> lights;
    struct Img { int w, h; std::vector<uint8_t> rgb; };
    std::vector<Img> images;
    std::map<int, std::unique_ptr<PerDevice>> devices;
    ftb::DeviceBuildStats bvh_stats;  // mesh index built on the device at create time (bvh_build.h)
    bool bvh_on_device = false;
};

namespace {

template <typename R> struct Mk4;
template <> struct Mk4<float> { static float4 make(double a, double b, double c, double d) { return make_float4((float)a, (float)b, (float)c, (float)d); } };
template <> struct Mk4<double> { static double4 make(double a, double b, double c, double d) { return make_double4(a, b, c, d); } };

template <typename T>
int upload(std::vector<void*>& allocs, const std::vector<T>& host, const T*& dev)
{
    dev = nullptr;
    if (host.empty()) return FTB_OK;
    void* p = nullptr;
    CK(cudaMalloc(&p, host.size() * sizeof(T)));
    allocs.push_back(p);
    CK(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    dev = static_cast<const T*>(p);
    return FTB_OK;
}

// Meshes of this many (clipped) triangles or more are walked by the whole warp (render.cuh packetMesh) and their samples are
// dealt one at a time; smaller ones by each lane for itself.  Measured on 960 / 9.6 k / 355 k triangles at 3840x2160x16: private
// walks 2.92 / 3.41 / 9.11 ms, packet walks 3.19 / 3.60 / 6.38 ms.
constexpr size_t kLargeMesh = 32768;
inline bool largeMesh(const Lowered& L) { return L.bvh_tri.size() >= kLargeMesh; }

template <typename R>
int uploadScene(const ftb_scene& sc, SceneStorage<R>& st)
{
    typedef typename V4<R>::type R4;
    const Lowered& L = sc.L;
    DevScene<R>& v = st.view;
    std::memset(&v, 0, sizeof(v));
    int rc;
#define UP(vec, field) if ((rc = upload(st.allocs, vec, field)) != FTB_OK) { st.release(); return rc; }
    {
        std::vector<R4> w2m, p0; std::vector<int4> meta;
        for (const Leaf& lf : L.leaves) {
            for (int r = 0; r < 3; ++r) w2m.push_back(Mk4<R>::make(lf.w2m[4 * r], lf.w2m[4 * r + 1], lf.w2m[4 * r + 2], lf.w2m[4 * r + 3]));
            {  // world point of the model origin: solve A p0 + b = 0 (Cramer; singular matrices keep p0 = 0)
                const double* m = lf.w2m;
                const double c00 = m[5] * m[10] - m[6] * m[9], c01 = m[6] * m[8] - m[4] * m[10], c02 = m[4] * m[9] - m[5] * m[8];
                const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
                double q[3] = {0, 0, 0};
                if (std::fabs(det) > 0 && std::isfinite(det)) {
                    const double id = 1.0 / det, bx = -m[3], by = -m[7], bz = -m[11];
                    q[0] = (bx * c00 + by * (m[2] * m[9] - m[1] * m[10]) + bz * (m[1] * m[6] - m[2] * m[5])) * id;
                    q[1] = (bx * c01 + by * (m[0] * m[10] - m[2] * m[8]) + bz * (m[2] * m[4] - m[0] * m[6])) * id;
                    q[2] = (bx * c02 + by * (m[1] * m[8] - m[0] * m[9]) + bz * (m[0] * m[5] - m[1] * m[4])) * id;
                }
                p0.push_back(Mk4<R>::make(q[0], q[1], q[2], 0));
            }
            meta.push_back(make_int4(lf.kind | (lf.identity << 8) | (lf.top_level << 9), lf.surface, lf.prim, lf.payload));
        }
        UP(w2m, v.leaf_w2m) UP(meta, v.leaf_meta) UP(p0, v.leaf_p0)
        v.n_leaves = (int)L.leaves.size();
    }
    {
        std::vector<int4> items; std::vector<R4> bounds; std::vector<int2> ops, progs;
        for (const Item& it : L.items) {
            // kind | CSG op << 8 | kind word of leaf A << 12 | of leaf B << 21 (kind word = leaf kind | identity << 8, as in leaf_meta.x):
            // the walk branches on a top-level leaf's / a pair operand's kind without waiting for the leaf's own record
            auto kw = [&](int leaf) { const Leaf& lf = L.leaves[(size_t)(leaf & 0xffffff)]; return (lf.kind & 0xff) | (lf.identity ? 0x100 : 0); };
            int kx = it.kind & 0xfff;
            if ((it.kind & 0xff) == ITEM_LEAF) kx |= kw(it.a) << 12;
            else if ((it.kind & 0xff) == ITEM_CSG2) kx |= (kw(it.a) << 12) | (kw(it.b) << 21);
            items.push_back(make_int4(kx, it.a, it.b, it.casts_shadow));
            progs.push_back(make_int2(it.prog_first, it.prog_count));
            // conservative: radius inflated by 0.2 % + 1e-5 so that FP32 rounding of the test cannot cull a true hit
            // unbounded items (planes) carry r^2 = +inf: no line misses them and no origin is outside them
            const double ri = it.bound_r < 0 ? -1.0 : it.bound_r * 1.002 + 1e-5;
            bounds.push_back(Mk4<R>::make(it.bound_c[0], it.bound_c[1], it.bound_c[2], ri < 0 ? (double)INFINITY : ri * ri));
        }
        for (const CsgOp& op : L.ops) ops.push_back(make_int2(op.kind, op.arg));
        std::vector<unsigned> casts((L.items.size() + 31) / 32 + 1, 0u);
        for (size_t i = 0; i < L.items.size(); ++i)
            if (L.items[i].casts_shadow) casts[i >> 5] |= 1u << (i & 31);
        std::vector<unsigned> meshes(casts.size(), 0u);
        for (size_t i = 0; i < L.items.size(); ++i)
            if (L.items[i].kind == ITEM_LEAF && L.leaves[(size_t)L.items[i].a].kind == LEAF_MESH) meshes[i >> 5] |= 1u << (i & 31);
        std::vector<R4> bounds2;  // pairs of neighbouring items interleaved (render.cuh: the packed bound test of the FP32 kernels)
        for (size_t i = 0; i < bounds.size(); i += 2) {
            const R4 a = bounds[i], b = i + 1 < bounds.size() ? bounds[i + 1] : Mk4<R>::make(0, 0, 0, (double)INFINITY);
            bounds2.push_back(Mk4<R>::make(a.x, b.x, a.y, b.y));
            bounds2.push_back(Mk4<R>::make(a.z, b.z, a.w, b.w));
        }
        UP(bounds2, v.item_bound2)
        UP(items, v.items) UP(bounds, v.item_bound) UP(ops, v.ops) UP(casts, v.item_casts) UP(meshes, v.item_mesh) UP(progs, v.item_prog)
        v.mesh_packet = largeMesh(L) ? 1 : 0;
        v.n_items = (int)L.items.size();
    }
    {
        std::vector<R4> a, b; std::vector<int4> si;
        for (const Surface& s : L.surfaces) {
            a.push_back(Mk4<R>::make(s.colour[0], s.colour[1], s.colour[2], s.roughness));
            b.push_back(Mk4<R>::make(s.reflectance, s.shineyness, 0, 0));
            si.push_back(make_int4(s.texture, s.hue, s.apply_lighting, 0));
        }
        UP(a, v.surf_a) UP(b, v.surf_b) UP(si, v.surf_i)
    }
    {
        std::vector<int4> ti; std::vector<R4> c1, c2; std::vector<int> kinds; std::vector<R> ab;
        for (const TexDef& t : L.textures) {
            ti.push_back(make_int4(t.op_first, t.op_count, t.base_kind, t.image));
            c1.push_back(Mk4<R>::make(t.c1[0], t.c1[1], t.c1[2], 0));
            c2.push_back(Mk4<R>::make(t.c2[0], t.c2[1], t.c2[2], 0));
        }
        for (const TexOp& o : L.tex_ops) { kinds.push_back(o.kind); ab.push_back((R)o.a); ab.push_back((R)o.b); }
        UP(ti, v.tex_i) UP(c1, v.tex_c1) UP(c2, v.tex_c2) UP(kinds, v.texop_kind) UP(ab, v.texop_ab)
        std::vector<uchar4> texels; std::vector<int4> ii;
        for (const ftb_scene::Img& im : sc.images) {
            ii.push_back(make_int4((int)texels.size(), im.w, im.h, 0));
            const size_t n = (size_t)im.w * im.h;
            for (size_t k = 0; k < n; ++k) texels.push_back(make_uchar4(im.rgb[3 * k], im.rgb[3 * k + 1], im.rgb[3 * k + 2], 255));
        }
        UP(texels, v.texels) UP(ii, v.img_i)
    }
    {
        std::vector<int> roots(L.mesh_root.begin(), L.mesh_root.end()); std::vector<R4> box, btris, tris;
        for (const BvhNode& n : L.bvh_nodes) {
            for (int k = 0; k < 3; ++k) {  // one row per axis: (L.lo, R.lo, L.hi, R.hi)
                if (sizeof(R) == 4) box.push_back(Mk4<R>::make(n.lo[0][k], n.lo[1][k], n.hi[0][k], n.hi[1][k]));
                else box.push_back(Mk4<R>::make(n.dlo[0][k], n.dlo[1][k], n.dhi[0][k], n.dhi[1][k]));
            }
            R4 lk = Mk4<R>::make(n.child[0], n.child[1], 0, 0);  // the links in the record's fourth row: FP64 as values, FP32 as bits
            if (sizeof(R) == 4) { std::memcpy(&lk.x, &n.child[0], 4); std::memcpy(&lk.y, &n.child[1], 4); }
            box.push_back(lk);
        }
        const size_t nt = sc.triangles.size() / 9;
        auto pushTri = [&](std::vector<R4>& dst, size_t t, double w0, double w1) {  // v0, e1 = v1 - v0, e2 = v2 - v0 (Triangle.fs:45-46), differences taken in double
            const double* q = sc.triangles.data() + 9 * t;
            dst.push_back(Mk4<R>::make(q[0], q[1], q[2], w0));
            dst.push_back(Mk4<R>::make(q[3] - q[0], q[4] - q[1], q[5] - q[2], w1));
            dst.push_back(Mk4<R>::make(q[6] - q[0], q[7] - q[1], q[8] - q[2], 0));
        };
        for (size_t t = 0; t < nt; ++t) pushTri(tris, t, 0, 0);
        // slot ids ride in the w components as reals: exact up to 2^24 in float
        if (sizeof(R) == 4 && (L.bvh_tri.size() >= (1u << 24) || nt >= (1u << 24))) { st.release(); return fail(FTB_ERR_UNSUPPORTED, "more than 16M mesh triangles"); }
        for (size_t k = 0; k < L.bvh_tri.size(); ++k) pushTri(btris, (size_t)L.bvh_tri[k], (double)L.bvh_seq[k], (double)L.bvh_tri[k]);
        UP(roots, v.mesh_root) UP(box, v.bvh_node) UP(btris, v.bvh_tris) UP(tris, v.tris)
    }
    {
        std::vector<int2> li; std::vector<R4> la, lb, lc;
        for (const ftb_light& l : sc.lights) {
            li.push_back(make_int2(l.kind, l.samples));
            la.push_back(Mk4<R>::make(l.v[0], l.v[1], l.v[2], std::tan(l.scatter_rad / 2.0)));  // Jitter.fs:29
            lb.push_back(Mk4<R>::make(l.falloff[0], l.falloff[1], l.falloff[2], 0));
            lc.push_back(Mk4<R>::make(l.colour[0], l.colour[1], l.colour[2], 0));
        }
        UP(li, v.light_i) UP(la, v.light_a) UP(lb, v.light_b) UP(lc, v.light_c)
        v.n_lights = (int)sc.lights.size();
    }
    {  // the inverse of `rotate unitZ 180` (Cylinder.fs:27) exactly as Transform.matrix builds it (Transform.fs:60-69, CommonTypes.fs:98-99)
        const double ang = -(180.0 * 1.0 * (3.14159265358979323846 / 180.0));
        v.cyl_c = (R)std::cos(ang); v.cyl_s = (R)std::sin(ang);
    }
#undef UP
    st.ready = true;
    return FTB_OK;
}

int getDevice(ftb_scene* sc, int device, PerDevice** out)
{
    auto it = sc->devices.find(device);
    if (it == sc->devices.end()) {
        std::unique_ptr<PerDevice> pd(new PerDevice);  // ~PerDevice releases whatever exists if one of the calls below fails
        pd->device = device;
        CK(cudaSetDevice(device));
        CK(cudaDeviceGetAttribute(&pd->sm_count, cudaDevAttrMultiProcessorCount, device));
        int prLow = 0, prHigh = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
        CK(cudaStreamCreateWithFlags(&pd->stream, cudaStreamNonBlocking));
        // the copy stream's kernels (assembly of a finished band) go ahead of the next band's render blocks
        CK(cudaStreamCreateWithPriority(&pd->copy_stream, cudaStreamNonBlocking, prHigh));
        CK(cudaEventCreate(&pd->ev0));
        CK(cudaEventCreate(&pd->ev1));
        CK(cudaEventCreateWithFlags(&pd->done, cudaEventDisableTiming));
        it = sc->devices.emplace(device, std::move(pd)).first;
    }
    *out = it->second.get();
    return FTB_OK;
}

struct FrameGeom {
    int gw, gh, spp, tiles_x, tiles_y, n_tiles, shard_index, shard_count, n_local_tiles;
    int band_index, band_count;  // ftb_render_tiles_device: only the tiles of one band of tile rows (band_count <= 1: all)
    long long n_samples;
    bool corner;
};

int frameGeom(const ftb_render_params* p, FrameGeom& g)
{
    if (!p) return fail(FTB_ERR_BAD_ARG, "null params");
    if (p->width < 1 || p->height < 1) return fail(FTB_ERR_BAD_ARG, "bad resolution");
    if ((long long)p->width * p->height > (1LL << 30)) return fail(FTB_ERR_BAD_ARG, "resolution too large");
    if (p->sampling != FTB_SAMPLING_JITTER && p->sampling != FTB_SAMPLING_CORNER) return fail(FTB_ERR_BAD_ARG, "bad sampling mode");
    if (p->precision != FTB_PRECISION_FP32 && p->precision != FTB_PRECISION_FP64_VERIFY) return fail(FTB_ERR_BAD_ARG, "bad precision");
    if (p->out_format < FTB_OUT_RGB_F64 || p->out_format > FTB_OUT_RGBA8) return fail(FTB_ERR_BAD_ARG, "bad output format");
    g.corner = p->sampling == FTB_SAMPLING_CORNER;
    if (!g.corner && (p->spp < 1 || !p->jitter_xy)) return fail(FTB_ERR_BAD_ARG, "jitter mode needs spp >= 1 and jitter_xy");
    g.gw = g.corner ? p->width + 1 : p->width;
    g.gh = g.corner ? p->height + 1 : p->height;
    g.spp = g.corner ? 1 : p->spp;
    g.tiles_x = (g.gw + FTB_TILE_W - 1) / FTB_TILE_W;
    g.tiles_y = (g.gh + FTB_TILE_H - 1) / FTB_TILE_H;
    g.n_tiles = g.tiles_x * g.tiles_y;
    g.shard_count = p->shard_count > 1 ? p->shard_count : 1;
    g.shard_index = p->shard_count > 1 ? p->shard_index : 0;
    if (g.shard_count > kMaxShards) return fail(FTB_ERR_BAD_ARG, "shard_count exceeds 64");
    if (g.shard_index < 0 || g.shard_index >= g.shard_count) return fail(FTB_ERR_BAD_ARG, "shard_index out of range");
    g.n_local_tiles = (g.n_tiles + g.shard_count - 1) / g.shard_count;  // one per group of shard_count tiles (device_scene.h tileOfLocal); the last group may lack this shard's
    g.n_samples = (long long)g.gw * g.gh * g.spp;
    g.band_count = p->band_count > 1 ? p->band_count : 1;
    g.band_index = p->band_count > 1 ? p->band_index : 0;
    if (g.band_count > 64) return fail(FTB_ERR_BAD_ARG, "band_count exceeds 64");
    if (g.band_index < 0 || g.band_index >= g.band_count) return fail(FTB_ERR_BAD_ARG, "band_index out of range");
    return FTB_OK;
}

inline size_t realSize(int precision) { return precision == FTB_PRECISION_FP64_VERIFY ? sizeof(double) : sizeof(float); }

struct V3 { double x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, b.x * a.z - b.z * a.x, a.x * b.y - a.y * b.x}; }
inline V3 normalise(V3 v)
{
    double l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    if (l < 0.0000001) return v;
    double s = 1.0 / l;
    return {s * v.x, s * v.y, s * v.z};
}

// ImagePlane.create / fromCamera (Image.fs:48-53, 67-81), evaluated in double on the host; the
// FP32 kernels receive the narrowed results.
template <typename R>
void fillCamera(DevFrame<R>& F, const ftb_camera& c, int resH, int resV)
{
    V3 o = {c.o[0], c.o[1], c.o[2]}, la = {c.look_at[0], c.look_at[1], c.look_at[2]}, up = {c.up[0], c.up[1], c.up[2]};
    V3 k = normalise(sub(la, o));
    V3 i = normalise(cross(up, k));
    V3 j = cross(k, i);
    double height = std::tan(c.fov_y_rad / 2.0) * 2.0;
    double width = height * c.aspect_ratio;
    double pixelHeight = height / (double)(resH - 1);  // the reference's swapped axes (Image.fs:71-72)
    double pixelWidth = width / (double)(resV - 1);
    F.cam_o[0] = (R)o.x; F.cam_o[1] = (R)o.y; F.cam_o[2] = (R)o.z;
    F.primary_slack = (R)(2e-4 * std::sqrt(o.x * o.x + o.y * o.y + o.z * o.z));
    F.cam_k[0] = (R)k.x; F.cam_k[1] = (R)k.y; F.cam_k[2] = (R)k.z;
    F.cam_i[0] = (R)i.x; F.cam_i[1] = (R)i.y; F.cam_i[2] = (R)i.z;
    F.cam_j[0] = (R)j.x; F.cam_j[1] = (R)j.y; F.cam_j[2] = (R)j.z;
    F.pw = (R)pixelWidth; F.ph = (R)pixelHeight;
    F.tlx = (R)(-width / 2.0 + pixelWidth / 2.0);
    F.tly = (R)(height / 2.0 - pixelHeight / 2.0);
    F.has_focus = c.has_focus;
    F.focal = (R)c.focal_length;
    F.tan_half_aperture = (R)std::tan(c.aperture_rad / 2.0);
}

void fillStats(const Control& h, const ftb_scene& sc, ftb_stats* s)
{
    const unsigned long long* c = h.stats;
    s->primary_rays += c[ST_PRIMARY]; s->shadow_rays += c[ST_SHADOW]; s->reflection_rays += c[ST_REFLECTION]; s->shaded_hits += c[ST_SHADED];
    // device leaf kinds -> ftb_prim_kind slots.  A cube and a solidCylinder are one fused leaf each.
    static const int slot[kLeafKinds] = {FTB_PRIM_SPHERE, FTB_PRIM_PLANE, FTB_PRIM_SQUARE, FTB_PRIM_CIRCLE, FTB_PRIM_CYLINDER, FTB_PRIM_CONE, FTB_PRIM_CUBE, FTB_PRIM_TRIANGLE, FTB_PRIM_BSPMESH,
                                         FTB_PRIM_SOLIDCYLINDER};
    for (int k = 0; k < kLeafKinds; ++k) s->leaf_tests[slot[k]] += c[ST_LEAF0 + k];
    s->leaf_tests[FTB_PRIM_TRIANGLE] += c[ST_TRI_TESTS_IN_MESH];
    s->transformed_leaf_tests += c[ST_XFORM]; s->bsp_nodes_visited += c[ST_BSP_NODES]; s->bound_tests += c[ST_BOUND_TESTS] + c[ST_BOUND_FAST]; s->csg_ops += c[ST_CSG_OPS];
    // algorithmic flops, SURVEY.md 8(d) table (FMA = 2; compares / selects = 0)
    static const double F[kLeafKinds] = {28, 8, 8, 8, 26, 32, 20, 45, 0, 42};  // solidCylinder: side + 2 caps in one frame
    double f = 0;
    for (int k = 0; k < kLeafKinds; ++k) f += F[k] * (double)c[ST_LEAF0 + k];
    f += 45.0 * (double)c[ST_TRI_TESTS_IN_MESH] + 33.0 * (double)c[ST_XFORM] + 12.0 * (double)c[ST_BSP_NODES] + 17.0 * (double)c[ST_BOUND_TESTS]   // bound test: 3 sub + 2 dot (5 each) + 2 FMA
         + 5.0 * (double)c[ST_BOUND_FAST];  // common-origin form: one dot
    f += (double)c[ST_SHADED] * (60.0 + 110.0 * (double)sc.lights.size()) + 18.0 * (double)c[ST_REFLECTION];
    s->flops += f;
}

// Bands of tile rows for the host-buffer path: band k = tile rows [bandFirstRow(k), bandFirstRow(k + 1)).  The D2H copy
// of band k overlaps the rendering of band k + 1, so only the LAST band's copy is exposed: the bands shrink towards
// the end (30 / 30 / 25 / 15 % for four bands).
inline int bandFirstRow(int k, int n, int tiles_y)
{
    if (k <= 0) return 0;
    if (k >= n) return tiles_y;
    if (n == 4) { static const double cum[5] = {0.0, 0.30, 0.60, 0.85, 1.0}; return (int)(cum[k] * tiles_y + 0.5); }
    return (int)(((long long)k * tiles_y) / n);
}
inline int bandOfRow(int row, int n, int tiles_y)
{
    int k = 0;
    while (k + 1 < n && row >= bandFirstRow(k + 1, n, tiles_y)) ++k;
    return k;
}

// Longest-processing-time-first tile order.  The persistent kernel hands tiles out from an atomic queue; a lane
// keeps its pixel for all samples and all bounce generations, so the last tiles handed out set the tail.  A
// static estimate - the projected bounding spheres of the items, weighted by how much work a hit on them
// causes (CSG programs, reflective surfaces spawn up to recursion_limit more generations) - puts the costly
// tiles first, so the tail consists of cheap ones.  Ordering cannot change any pixel: tiles are independent.
// Returns false when every tile has the same estimate (no order needed).
bool computeTileOrder(const ftb_scene& sc, const ftb_camera& c, const ftb_render_params& p, const FrameGeom& g, int n_chunks, std::vector<int>& order,
                      std::vector<int>& chunk_first)
{
    const double kPiHalf = 1.5707963267948966;
    V3 o = {c.o[0], c.o[1], c.o[2]}, la = {c.look_at[0], c.look_at[1], c.look_at[2]}, up = {c.up[0], c.up[1], c.up[2]};
    V3 k = normalise(sub(la, o));
    V3 i = normalise(cross(up, k));
    V3 j = cross(k, i);
    const double height = std::tan(c.fov_y_rad / 2.0) * 2.0, width = height * c.aspect_ratio;
    const double ph = height / (double)(p.width - 1), pw = width / (double)(p.height - 1);  // Image.fs:71-72 (swapped axes)
    const double tlx = -width / 2.0 + pw / 2.0, tly = height / 2.0 - ph / 2.0;
    std::vector<float> cost((size_t)g.n_tiles, 1.0f);
    bool any = false;
    auto range = [&](double a, double z, double r, double& lo, double& hi) {
        const double d2 = a * a + z * z;
        if (d2 <= r * r) { lo = -1e300; hi = 1e300; return true; }
        const double th = std::atan2(a, z), al = std::asin(r / std::sqrt(d2));
        const double t0 = th - al, t1 = th + al;
        if (t0 >= kPiHalf || t1 <= -kPiHalf) return false;  // entirely behind the image plane
        lo = t0 <= -kPiHalf ? -1e300 : std::tan(t0);
        hi = t1 >= kPiHalf ? 1e300 : std::tan(t1);
        return true;
    };
    for (const Item& it : sc.L.items) {
        if (it.bound_r < 0) continue;  // unbounded (planes): the same everywhere
        double w = 1.0;
        bool reflective = false;
        auto leafWork = [&](int leaf) {
            const Leaf& lf = sc.L.leaves[leaf];
            const Surface& sf = sc.L.surfaces[lf.surface];
            if (sf.apply_lighting && sf.reflectance > 0.0) reflective = true;
        };
        if (it.kind == ITEM_LEAF) leafWork(it.a);
        else {
            w = 0.0;
            for (int q = it.prog_first; q < it.prog_first + it.prog_count; ++q)
                if (sc.L.ops[q].kind == OP_LEAF) { leafWork(sc.L.ops[q].arg); w += 2.0; }
        }
        w *= (1.0 + (double)sc.lights.size());                              // a shaded hit adds one shadow ray per light
        if (reflective) w *= 1.0 + (double)std::max(0, p.recursion_limit);   // and up to `limit` more generations
        V3 v = {it.bound_c[0] - o.x, it.bound_c[1] - o.y, it.bound_c[2] - o.z};
        const double x = v.x * i.x + v.y * i.y + v.z * i.z, y = v.x * j.x + v.y * j.y + v.z * j.z, z = v.x * k.x + v.y * k.y + v.z * k.z;
        double xl, xh, yl, yh;
        if (!range(x, z, it.bound_r, xl, xh) || !range(y, z, it.bound_r, yl, yh)) continue;
        // image-plane coordinates -> sample-grid pixels (rayThroughPixel, Image.fs:83-89), one pixel of slack for jitter
        const double px0 = (xl - tlx) / pw - 1.5, px1 = (xh - tlx) / pw + 1.5, py0 = (tly - yh) / ph - 1.5, py1 = (tly - yl) / ph + 1.5;
        if (px1 < 0 || py1 < 0 || px0 >= g.gw || py0 >= g.gh) continue;
        const int tx0 = (int)std::max(0.0, std::floor(px0 / FTB_TILE_W)), tx1 = (int)std::min((double)g.tiles_x - 1, std::floor(px1 / FTB_TILE_W));
        const int ty0 = (int)std::max(0.0, std::floor(py0 / FTB_TILE_H)), ty1 = (int)std::min((double)g.tiles_y - 1, std::floor(py1 / FTB_TILE_H));
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) cost[(size_t)ty * g.tiles_x + tx] += (float)w;
        any = true;
    }
    // chunks = bands of tile rows (only used unsharded), see bandFirstRow
    auto tileOf = [&](int l) { return tileOfLocal(l, g.shard_index, g.shard_count); };
    auto chunkOf = [&](int l) { return n_chunks <= 1 ? 0 : bandOfRow(std::min(tileOf(l), g.n_tiles - 1) / g.tiles_x, n_chunks, g.tiles_y); };
    chunk_first.assign((size_t)n_chunks + 1, 0);
    for (int l = 0; l < g.n_local_tiles; ++l) chunk_first[(size_t)chunkOf(l) + 1]++;
    for (int k = 0; k < n_chunks; ++k) chunk_first[(size_t)k + 1] += chunk_first[(size_t)k];
    if (!any && n_chunks <= 1) return false;
    order.resize((size_t)g.n_local_tiles);
    for (int l = 0; l < g.n_local_tiles; ++l) order[l] = l;
    auto costOf = [&](int l) { const int t = tileOf(l); return t < g.n_tiles ? cost[(size_t)t] : 0.0f; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const int ca = chunkOf(a), cb = chunkOf(b);
        return ca != cb ? ca < cb : costOf(a) > costOf(b);
    });
    return true;
}

// Renders one shard's tiles (mode 0) on the CURRENT device into d_tiles.  Everything is queued on
// `stream`; nothing here synchronises unless stats are requested.
template <typename R>
int launchFrame(ftb_scene* sc, PerDevice* pd, const ftb_camera* cam, const ftb_render_params* p, const FrameGeom& g, void* d_tiles,
                const ftb_debug_out* d_dbg, ftb_stats* stats, cudaStream_t stream, bool timeKernel, int n_chunks, const ChunkDone* chunkDone, int only_chunk)
{
    SceneStorage<R>& st = storageOf<R>(pd);
    if (!st.ready) { int rc = uploadScene<R>(*sc, st); if (rc != FTB_OK) return rc; }
    CK(pd->control.reserve(sizeof(Control)));
    if (!pd->control_ready) { CK(cudaMemset(pd->control.p, 0, sizeof(Control))); pd->control_ready = true; }
    // The queue counter is reset before every launch, the counters only when they are asked for; the overflow flag stays
    // up until somebody reads it (finishStats / ftb_check_overflow), so that the banded and the device paths cannot lose it.
    if (stats && p->collect_stats) CK(cudaMemsetAsync(static_cast<Control*>(pd->control.p)->stats, 0, sizeof(static_cast<Control*>(pd->control.p)->stats), stream));
    DevFrame<R> F;
    std::memset(&F, 0, sizeof(F));
    F.mode = 0;
    F.gw = g.gw; F.gh = g.gh; F.spp = g.spp; F.tiles_x = g.tiles_x;
    // floor(2^32 / tiles_x) + 1 does not fit 32 bits for tiles_x == 1 (frames <= 16 sample columns wide): those divide
    F.tiles_x_magic = (g.tiles_x > 1 && (unsigned long long)g.n_tiles * (unsigned long long)g.tiles_x < (1ull << 32)) ? (unsigned)((1ull << 32) / (unsigned)g.tiles_x) + 1u : 0u;
    F.n_local_tiles = g.n_local_tiles; F.shard_index = g.shard_index; F.shard_count = g.shard_count;
    fillCamera<R>(F, *cam, p->width, p->height);
    if (!g.corner) {
        std::vector<R> j(2 * (size_t)g.spp);
        for (size_t k = 0; k < j.size(); ++k) j[k] = (R)p->jitter_xy[k];
        CK(pd->jitter.reserve(j.size() * sizeof(R)));
        CK(cudaMemcpyAsync(pd->jitter.p, j.data(), j.size() * sizeof(R), cudaMemcpyHostToDevice, stream));  // pageable source: staged before return
        F.jitter = static_cast<const R*>(pd->jitter.p);
    } else {
        // CornerSampling.generateRays (Image.fs:128-132): every corner ray uses the offset (-0.5, +0.5)
        const R j[2] = {R(-0.5), R(0.5)};
        CK(pd->jitter.reserve(sizeof(j)));
        CK(cudaMemcpyAsync(pd->jitter.p, j, sizeof(j), cudaMemcpyHostToDevice, stream));
        F.jitter = static_cast<const R*>(pd->jitter.p);
    }
    F.recursion_limit = p->recursion_limit;
    F.seed = p->seed;
    F.out = static_cast<R*>(d_tiles);
    if (d_dbg) { F.dbg_prim = d_dbg->prim_id; F.dbg_sub = d_dbg->sub_id; F.dbg_t = d_dbg->t; }
    {  // tile order: cached per (camera, frame geometry, recursion limit); recomputed + uploaded only when they change
        std::vector<unsigned char> key(sizeof(ftb_camera) + 7 * sizeof(int));
        std::memcpy(key.data(), cam, sizeof(ftb_camera));
        const int kk[7] = {p->width, p->height, p->sampling, g.shard_index, g.shard_count, p->recursion_limit, n_chunks};
        std::memcpy(key.data() + sizeof(ftb_camera), kk, sizeof(kk));
        if (key != pd->order_key) {
            std::vector<int> order;
            pd->order_valid = computeTileOrder(*sc, *cam, *p, g, n_chunks, order, pd->chunk_first);
            if (pd->order_valid) {
                CK(pd->order.reserve(order.size() * sizeof(int)));
                CK(cudaMemcpyAsync(pd->order.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
                CK(cudaStreamSynchronize(stream));  // `order` is a pageable local
            }
            pd->order_key.swap(key);
        }
        static const bool noOrder = std::getenv("FTB_NO_TILE_ORDER") != nullptr;  // A/B switch for measurements
        F.tile_order = (pd->order_valid && (!noOrder || n_chunks > 1)) ? static_cast<const int*>(pd->order.p) : nullptr;
    }
    const int* orderBase = F.tile_order;
    Control* ctl = static_cast<Control*>(pd->control.p);
    F.tile_counter = &ctl->tile_counter; F.overflow = &ctl->overflow; F.stats = ctl->stats;
    const bool wantStats = stats && p->collect_stats;
    if (timeKernel) CK(cudaEventRecord(pd->ev0, stream));
    int launches = 0;
    const Variant<R>* var = pickVariant<R>(sc->L.features | (cam->has_focus ? (unsigned)FT_RNG : 0u), wantStats);
    if (!var) return fail(FTB_ERR_UNSUPPORTED, "no kernel variant covers this scene's feature mask");
    // one launch covers at most UnitCap samples per pixel; more samples = more passes over the same tiles, each
    // continuing the left fold of the previous one (so the blend order is still Array.average's)
    // Work-queue granularity.  A lane takes runs of samples, a warp takes blocks of pixels from the atomic queue.  The
    // block is the indivisible quantum of the tail, so it shrinks (32 = 8x4 ... 1 pixel) until there are >= 24 blocks
    // per resident warp, but never below one warp-round of samples per block (and big frames keep big blocks, which
    // also bounds the number of atomics on the queue counter).
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        if (only_chunk >= 0 && chunk != only_chunk) continue;
        const int first = n_chunks > 1 ? pd->chunk_first[(size_t)chunk] : 0;
        const int ntile = n_chunks > 1 ? pd->chunk_first[(size_t)chunk + 1] - first : g.n_local_tiles;
        if (ntile > 0) {
            F.tile_order = orderBase ? orderBase + first : nullptr;
            const long long warps = (long long)pd->sm_count * 5 * (kBlockThreads / 32);
            const long long pixels = (long long)ntile * FTB_TILE_PIXELS;
            int ppb = 32;
            while (ppb > 1 && pixels / ppb < 24 * warps && (long long)(ppb / 2) * g.spp >= 32) ppb /= 2;
            int bw = ppb >= 8 ? 8 : ppb, bh = ppb / bw;
            F.bw_log = bw == 8 ? 3 : (bw == 4 ? 2 : (bw == 2 ? 1 : 0));
            F.bh_log = bh == 4 ? 2 : (bh == 2 ? 1 : 0);
            F.n_blocks = ntile * (FTB_TILE_PIXELS / ppb);
            // A/B arm (FTB_WAVEFRONT=1): the same frame through the wavefront kernels of wavefront.cuh, where the variant has them
            static const bool wantWavefront = std::getenv("FTB_WAVEFRONT") != nullptr;
            if (wantWavefront && !wantStats) {
                int rph = 0;
                for (const ftb_light& l : sc->lights) rph += l.kind == FTB_LIGHT_SOFT_DIRECTIONAL ? std::max(l.samples, 0) : 1;
                rph = std::max(rph, 1);
                const size_t perPath = 4 * sizeof(typename V4<R>::type) + 3 * sizeof(int) + 3 * sizeof(R) + (size_t)rph * (2 * sizeof(typename V4<R>::type) + 2 * sizeof(int));
                const size_t want = std::min<size_t>((size_t)ntile * FTB_TILE_PIXELS * (size_t)g.spp * perPath + 65536, (size_t)8 << 30);
                CK(pd->wavefront.reserve(want));
                F.s_base = 0; F.s_count = g.spp; F.run = 1; F.rpp_magic = 0;
                cudaError_t we = var->launch_wavefront(st.view, F, ntile, rph, sc->L.has_reflection, pd->sm_count, pd->wavefront.p, pd->wavefront.cap, stream, &launches);
                if (we == cudaSuccess) { if (chunkDone) { int rc = (*chunkDone)(chunk); if (rc != FTB_OK) return rc; } continue; }
                if (we != cudaErrorNotSupported) return cudaFail(we, "wavefront launch");
                (void)cudaGetLastError();  // this variant has no wavefront arm: the megakernel below
            }
            for (int s_base = 0; s_base < g.spp; s_base += var->unit_cap) {
                F.s_base = s_base;
                F.s_count = std::min(var->unit_cap, g.spp - s_base);
                // run length: cheap samples amortise the dealing over up to 8 consecutive samples of a pixel, as long as
                // every pixel still splits into >= 8 runs (the chain a lane can be stuck with stays 1/8 of a pixel)
                // Large meshes deal single samples: a traversal's cost varies so much between neighbouring samples that a lane
                // stuck with a run of them holds its warp back (measured -11..-14 % on the 355 k mesh; +3 % on the 960-triangle
                // one and +7 % on moon the other way).
                static const int runEnv = [] { const char* e = std::getenv("FTB_RUN_MAX"); int v = e ? std::atoi(e) : 0; return v > 8 ? 8 : v; }();  // A/B switch
                const int runMax = runEnv > 0 ? runEnv : (largeMesh(sc->L) ? 1 : 8);
                int run = 1;
                while (run < runMax && F.s_count % (run * 2) == 0 && F.s_count / (run * 2) >= 8) run *= 2;
                F.run = run;
                const unsigned rpp = (unsigned)(F.s_count / run);
                F.rpp_magic = rpp <= 1 ? 0u : (unsigned)((1ull << 32) / rpp) + 1u;
                CK(cudaMemsetAsync(&ctl->tile_counter, 0, sizeof(unsigned int), stream));
                CK(var->launch(st.view, F, wantStats, pd->sm_count, stream, &launches));
            }
        }
        if (chunkDone) { int rc = (*chunkDone)(chunk); if (rc != FTB_OK) return rc; }
    }
    if (timeKernel) CK(cudaEventRecord(pd->ev1, stream));
    if (stats) stats->kernel_launches += launches + 1;  // + the control-block memset
    return FTB_OK;
}

int launchFrameAny(ftb_scene* sc, PerDevice* pd, const ftb_camera* cam, const ftb_render_params* p, const FrameGeom& g, void* d_tiles,
                   const ftb_debug_out* d_dbg, ftb_stats* stats, cudaStream_t stream, bool timeKernel, int n_chunks = 1, const ChunkDone* chunkDone = nullptr,
                   int only_chunk = -1)
{
    if (p->precision == FTB_PRECISION_FP64_VERIFY) return launchFrame<double>(sc, pd, cam, p, g, d_tiles, d_dbg, stats, stream, timeKernel, n_chunks, chunkDone, only_chunk);
    return launchFrame<float>(sc, pd, cam, p, g, d_tiles, d_dbg, stats, stream, timeKernel, n_chunks, chunkDone, only_chunk);
}

// Reads the control block back (synchronises the stream) and folds it into stats.
int finishStats(ftb_scene* sc, PerDevice* pd, ftb_stats* stats, cudaStream_t stream, bool timed, bool* overflow)
{
    Control h;
    CK(cudaMemcpyAsync(&h, pd->control.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (h.overflow) {
        *overflow = true;
        CK(cudaMemsetAsync(&static_cast<Control*>(pd->control.p)->overflow, 0, sizeof(unsigned int), stream));
    }
    if (stats) {
        fillStats(h, *sc, stats);
        if (h.overflow) stats->hit_overflow = 1;
        if (timed) { float ms = 0; CK(cudaEventElapsedTime(&ms, pd->ev0, pd->ev1)); if ((double)ms > stats->kernel_ms) stats->kernel_ms = ms; }
    }
    return FTB_OK;
}

int launchAssemble(const ftb_render_params* p, const FrameGeom& g, const void* const* bufs, void* d_out, cudaStream_t stream, int y0 = 0, int y1 = -1)
{
    TilePtrs tp;
    std::memset(&tp, 0, sizeof(tp));
    for (int i = 0; i < g.shard_count; ++i) tp.p[i] = bufs[i];
    if (y1 < 0) y1 = p->height;
    const long long n = (long long)p->width * (y1 - y0);
    if (n <= 0) return FTB_OK;
    int grid = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    if (p->precision == FTB_PRECISION_FP64_VERIFY)
        assemble_kernel<double><<<grid, 256, 0, stream>>>(tp, g.shard_count, g.tiles_x, p->width, y0, y1, g.corner ? 1 : 0, p->out_format, d_out);
    else
        assemble_kernel<float><<<grid, 256, 0, stream>>>(tp, g.shard_count, g.tiles_x, p->width, y0, y1, g.corner ? 1 : 0, p->out_format, d_out);
    CK(cudaGetLastError());
    return FTB_OK;
}

inline size_t outBytes(const ftb_render_params* p)
{
    const size_t n = (size_t)p->width * p->height;
    return p->out_format == FTB_OUT_RGB_F64 ? n * 24 : (p->out_format == FTB_OUT_RGB_F32 ? n * 12 : n * 4);
}

// CUDA ordinals of the devices this library runs on (compute capability 10.x), in ordinal order: what ftb_device_count
// counts and what ftb_render(n_gpus = N) uses, so that the two agree on a box with other GPUs in it.
std::vector<int> b200Devices()
{
    std::vector<int> out;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return out; }
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) out.push_back(d);
    }
    return out;
}

struct DeviceRestore {
    int dev = -1;
    DeviceRestore() { if (cudaGetDevice(&dev) != cudaSuccess) dev = -1; }
    ~DeviceRestore() { if (dev >= 0) cudaSetDevice(dev); }
};

}  // namespace

extern "C" {

int ftb_abi_version(void) { return FTB_ABI_VERSION; }

int ftb_device_count(void) { return (int)b200Devices().size(); }

const char* ftb_last_error(void) { return g_err.c_str(); }

int ftb_scene_create(const ftb_scene_desc* desc, ftb_scene** out)
{
    if (!desc || !out) return fail(FTB_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<ftb_scene> sc(new ftb_scene);
    std::string err;
    const auto tc0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        static const bool verbose = std::getenv("FTB_VERBOSE") != nullptr;
        if (verbose) std::fprintf(stderr, "functracer_b200: scene_create %-14s at %8.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
    };
    int rc = ftb::lower_scene(*desc, sc->L, err, false);  // the mesh index is built below, on the device
    lap("lowered");
    if (rc != FTB_OK) return fail(rc, err);
    if (sc->L.max_csg_lists > ftb::kMaxLists) return fail(FTB_ERR_UNSUPPORTED, "CSG nesting needs more than 12 pending hit lists");
    if (sc->L.leaves.size() >= (1u << 22)) return fail(FTB_ERR_UNSUPPORTED, "more than 4M leaves");
    sc->bsp_nodes.assign(desc->bsp_nodes, desc->bsp_nodes + desc->n_bsp_nodes);
    sc->bsp_leaves.assign(desc->bsp_leaves, desc->bsp_leaves + desc->n_bsp_leaves);
    sc->meshes.assign(desc->meshes, desc->meshes + desc->n_meshes);
    sc->triangles.assign(desc->triangles, desc->triangles + 9 * (size_t)desc->n_triangles);
    sc->lights.assign(desc->lights, desc->lights + desc->n_lights);
    for (int i = 0; i < desc->n_images; ++i) {
        ftb_scene::Img im;
        im.w = desc->images[i].width; im.h = desc->images[i].height;
        if (desc->images[i].rgb24 && im.w > 0 && im.h > 0) im.rgb.assign(desc->images[i].rgb24, desc->images[i].rgb24 + 3 * (size_t)im.w * im.h);
        else { im.w = im.h = 1; im.rgb.assign(3, 0); }
        sc->images.push_back(std::move(im));
    }
    lap("copied");
    // upload to the current device now, so that create fails loudly without a GPU
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) { (void)cudaGetLastError(); return fail(FTB_ERR_NO_DEVICE, "no CUDA device: functracer_b200 has no CPU path"); }
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (sc->L.has_mesh) {
        // BspMesh.fs:30-65 builds the reference's tree; the device's own index over the same triangles is built here: on the
        // GPU for large meshes (bvh_build.cu), on the host for small ones (and as fallback / A/B arm: FTB_HOST_BVH=1)
        static const bool hostBvh = std::getenv("FTB_HOST_BVH") != nullptr;
        static const bool deviceBvh = std::getenv("FTB_DEVICE_BVH") != nullptr;  // A/B: the device build also for small meshes
        static const int radiusEnv = [] { const char* e = std::getenv("FTB_PLOC_RADIUS"); return e ? std::atoi(e) : 0; }();
        std::vector<std::vector<int32_t>> order;
        ftb::enumerateMeshes(*desc, sc->L, order);
        size_t meshTriangles = 0;
        for (const auto& o : order) meshTriangles += o.size();
        bool done = false;
        if (!hostBvh && (deviceBvh || meshTriangles >= kLargeMesh)) {
            const int radii[3] = {radiusEnv > 0 ? radiusEnv : 16, 8, 4};
            for (int k = 0; k < 3 && !done; ++k) {
                std::string why;
                done = ftb::buildMeshIndexDevice(desc->triangles, desc->n_triangles, order, ftb::kBspStack, radii[k], sc->L, sc->bvh_stats, why);
                if (!done && std::getenv("FTB_VERBOSE")) std::fprintf(stderr, "functracer_b200: device BVH build not used (%s)\n", why.c_str());
            }
            sc->bvh_on_device = done;
        }
        if (!done) {
            const auto t0 = std::chrono::steady_clock::now();
            ftb::buildMeshIndexHost(*desc, sc->L);
            sc->bvh_stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        if (sc->L.max_bvh_depth + 2 > ftb::kBspStack) return fail(FTB_ERR_UNSUPPORTED, "mesh index deeper than the 64-entry traversal stack");
        if (largeMesh(sc->L)) sc->L.features |= (unsigned)FT_MESHPK;  // the warp-packet walk (render.cuh)
    }
    lap("mesh index");
    PerDevice* pd = nullptr;
    rc = getDevice(sc.get(), dev, &pd);
    if (rc != FTB_OK) return rc;
    rc = uploadScene<float>(*sc, pd->f32);
    if (rc != FTB_OK) return rc;
    lap("uploaded");
    *out = sc.release();
    return FTB_OK;
}

void ftb_scene_destroy(ftb_scene* sc)
{
    if (!sc) return;
    DeviceRestore restore;
    delete sc;  // ~PerDevice synchronises and releases every device's buffers, streams and events
}

int ftb_scene_build_info(const ftb_scene* scene, int32_t* bvh_on_device, double* bvh_build_ms, double* bvh_total_ms)
{
    if (!scene) return fail(FTB_ERR_BAD_ARG, "null argument");
    if (bvh_on_device) *bvh_on_device = scene->bvh_on_device ? 1 : 0;
    if (bvh_build_ms) *bvh_build_ms = scene->bvh_stats.build_ms;
    if (bvh_total_ms) *bvh_total_ms = scene->bvh_stats.total_ms;
    return FTB_OK;
}

int64_t ftb_tile_buffer_bytes(const ftb_render_params* params)
{
    FrameGeom g;
    int rc = frameGeom(params, g);
    if (rc != FTB_OK) return rc;
    return (int64_t)g.n_local_tiles * FTB_TILE_PIXELS * 3 * (int64_t)realSize(params->precision);
}

int ftb_render_tiles_device(ftb_scene* scene, const ftb_camera* camera, const ftb_render_params* params, void* d_tiles,
                            const ftb_debug_out* d_dbg, ftb_stats* stats, void* stream)
{
    if (!scene || !camera || !params || !d_tiles) return fail(FTB_ERR_BAD_ARG, "null argument");
    FrameGeom g;
    int rc = frameGeom(params, g);
    if (rc != FTB_OK) return rc;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return fail(FTB_ERR_NO_DEVICE, "no CUDA device"); }
    PerDevice* pd = nullptr;
    if ((rc = getDevice(scene, dev, &pd)) != FTB_OK) return rc;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if ((rc = launchFrameAny(scene, pd, camera, params, g, d_tiles, d_dbg, stats, s, stats != nullptr, g.band_count, nullptr, g.band_count > 1 ? g.band_index : -1)) != FTB_OK) return rc;
    if (stats) {  // the only synchronising path of this entry point
        bool overflow = false;
        if ((rc = finishStats(scene, pd, stats, s, true, &overflow)) != FTB_OK) return rc;
        if (overflow) return fail(FTB_ERR_HIT_OVERFLOW, "a CSG operand produced more than 32 crossings on one ray");
    }
    return FTB_OK;
}

int ftb_assemble_device(const ftb_render_params* params, const void* const* d_tile_buffers, void* d_out, void* stream)
{
    if (!params || !d_tile_buffers || !d_out) return fail(FTB_ERR_BAD_ARG, "null argument");
    FrameGeom g;
    int rc = frameGeom(params, g);
    if (rc != FTB_OK) return rc;
    for (int i = 0; i < g.shard_count; ++i)
        if (!d_tile_buffers[i]) return fail(FTB_ERR_BAD_ARG, "null tile buffer");
    return launchAssemble(params, g, d_tile_buffers, d_out, static_cast<cudaStream_t>(stream));
}

int ftb_band_rows(const ftb_render_params* params, int band_index, int band_count, int* y0, int* y1)
{
    if (!params || !y0 || !y1) return fail(FTB_ERR_BAD_ARG, "null argument");
    FrameGeom g;
    int rc = frameGeom(params, g);
    if (rc != FTB_OK) return rc;
    if (band_count < 1 || band_count > 64 || band_index < 0 || band_index >= band_count) return fail(FTB_ERR_BAD_ARG, "bad band");
    if (g.corner && band_count > 1) return fail(FTB_ERR_UNSUPPORTED, "corner sampling blends across tile rows: render it as one band");
    *y0 = std::min(params->height, bandFirstRow(band_index, band_count, g.tiles_y) * FTB_TILE_H);
    *y1 = std::min(params->height, bandFirstRow(band_index + 1, band_count, g.tiles_y) * FTB_TILE_H);
    return FTB_OK;
}

int ftb_assemble_rows_device(const ftb_render_params* params, const void* const* d_tile_buffers, void* d_out, int y0, int y1, void* stream)
{
    if (!params || !d_tile_buffers || !d_out) return fail(FTB_ERR_BAD_ARG, "null argument");
    FrameGeom g;
    int rc = frameGeom(params, g);
    if (rc != FTB_OK) return rc;
    if (y0 < 0 || y1 > params->height || y0 > y1) return fail(FTB_ERR_BAD_ARG, "bad row range");
    for (int i = 0; i < g.shard_count; ++i)
        if (!d_tile_buffers[i]) return fail(FTB_ERR_BAD_ARG, "null tile buffer");
    return launchAssemble(params, g, d_tile_buffers, d_out, static_cast<cudaStream_t>(stream), y0, y1);
}

int ftb_host_copy_begin(ftb_scene* scene, const void* d_src, void* host_dst, int64_t bytes, void* stream)
{
    if (!scene || bytes < 0 || (bytes > 0 && (!d_src || !host_dst))) return fail(FTB_ERR_BAD_ARG, "null argument");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return fail(FTB_ERR_NO_DEVICE, "no CUDA device"); }
    PerDevice* pd = nullptr;
    int rc = getDevice(scene, dev, &pd);
    if (rc != FTB_OK) return rc;
    CK(pd->copier.begin(d_src, host_dst, (size_t)bytes, static_cast<cudaStream_t>(stream)));
    return FTB_OK;
}

int ftb_host_copy_finish(ftb_scene* scene)
{
    if (!scene) return fail(FTB_ERR_BAD_ARG, "null argument");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return fail(FTB_ERR_NO_DEVICE, "no CUDA device"); }
    PerDevice* pd = nullptr;
    int rc = getDevice(scene, dev, &pd);
    if (rc != FTB_OK) return rc;
    CK(pd->copier.finish());
    return FTB_OK;
}

int ftb_check_overflow(ftb_scene* scene, void* stream)
{
    if (!scene) return fail(FTB_ERR_BAD_ARG, "null argument");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return fail(FTB_ERR_NO_DEVICE, "no CUDA device"); }
    PerDevice* pd = nullptr;
    int rc = getDevice(scene, dev, &pd);
    if (rc != FTB_OK) return rc;
    if (!pd->control.p) { CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); return FTB_OK; }  // nothing rendered yet
    bool overflow = false;
    if ((rc = finishStats(scene, pd, nullptr, static_cast<cudaStream_t>(stream), false, &overflow)) != FTB_OK) return rc;
    if (overflow) return fail(FTB_ERR_HIT_OVERFLOW, "a CSG operand produced more than 32 crossings on one ray");
    return FTB_OK;
}

int ftb_render(ftb_scene* scene, const ftb_camera* camera, const ftb_render_params* params, void* out, const ftb_debug_out* dbg, ftb_stats* stats)
{
    const auto t0 = std::chrono::steady_clock::now();
    if (!scene || !camera || !params || !out) return fail(FTB_ERR_BAD_ARG, "null argument");
    FrameGeom full;
    ftb_render_params p = *params;
    p.band_index = 0; p.band_count = 0;  // banding is this call's own business
    int rc = frameGeom(&p, full);
    if (rc != FTB_OK) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { (void)cudaGetLastError(); return fail(FTB_ERR_NO_DEVICE, "no CUDA device: functracer_b200 has no CPU path"); }
    if (stats) std::memset(stats, 0, sizeof(*stats));
    DeviceRestore restore;
    const int n_gpus = p.n_gpus > 1 ? p.n_gpus : 1;
    std::vector<int> devs;
    if (n_gpus > 1) {
        devs = b200Devices();
        if (n_gpus > (int)devs.size()) return fail(FTB_ERR_NO_DEVICE, "n_gpus exceeds the visible sm_100 devices (ftb_device_count)");
        devs.resize((size_t)n_gpus);
    } else devs.push_back(restore.dev);
    if (n_gpus > 1 && p.shard_count > 1) return fail(FTB_ERR_BAD_ARG, "n_gpus > 1 cannot be combined with an external shard");
    if (n_gpus > 1 && dbg) return fail(FTB_ERR_UNSUPPORTED, "debug planes need n_gpus <= 1");
    if (full.shard_count > 1 && n_gpus == 1) {
        // a single external shard cannot be assembled into a frame: the caller wants ftb_render_tiles_device
        return fail(FTB_ERR_BAD_ARG, "ftb_render renders whole frames; use ftb_render_tiles_device for one shard");
    }
    const size_t rs = realSize(p.precision);
    const size_t frameBytes = outBytes(&p);
    const size_t rowBytes = frameBytes / (size_t)p.height;
    // The frame is rendered in bands of tile rows, on every device; a finished band is assembled and sent to the host while
    // the next bands render, so only the last, smallest band's copy is exposed.  Four bands also for small downloads: a 1080p
    // RGBA8 frame (8 MB) measured 1.44 ms with four bands against 1.71 ms in one piece (hollow-sphere, 1.17 ms device-resident).
    static const int bandsEnv = [] { const char* e = std::getenv("FTB_BANDS"); return e ? std::atoi(e) : 0; }();  // A/B switch for measurements
    const bool bandable = !dbg && !stats && !full.corner && (long long)p.width * p.height >= 262144 && full.tiles_y >= 8;
    const int kBands = bandable ? (bandsEnv > 0 ? std::min(bandsEnv, 4) : 4) : 1;
    const int primary = devs[0];

    struct Shard { PerDevice* pd; FrameGeom g; ftb_render_params ps; void* target; bool direct; };
    std::vector<Shard> sh((size_t)n_gpus);
    std::vector<const void*> bufs((size_t)n_gpus);
    // ---- render: one shard per device, every band of every device queued before anything is waited on ----------------
    for (int k = 0; k < n_gpus; ++k) {
        const int dev = devs[(size_t)k];
        CK(cudaSetDevice(dev));
        Shard& S = sh[(size_t)k];
        if ((rc = getDevice(scene, dev, &S.pd)) != FTB_OK) return rc;
        PerDevice* pd = S.pd;
        S.ps = p;
        S.ps.shard_index = n_gpus > 1 ? k : 0; S.ps.shard_count = n_gpus;
        if ((rc = frameGeom(&S.ps, S.g)) != FTB_OK) return rc;
        const size_t tileBytes = (size_t)S.g.n_local_tiles * FTB_TILE_PIXELS * 3 * rs;
        CK(pd->tiles.reserve(tileBytes));
        while ((int)pd->band_events.size() < kBands) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            pd->band_events.push_back(e);
        }
        ftb_debug_out ddbg = {nullptr, nullptr, nullptr};
        if (dbg && dbg->prim_id) {
            CK(pd->dbg_prim.reserve((size_t)full.n_samples * 4));
            CK(cudaMemsetAsync(pd->dbg_prim.p, 0xff, (size_t)full.n_samples * 4, pd->stream));
            ddbg.prim_id = static_cast<int32_t*>(pd->dbg_prim.p);
            if (dbg->sub_id) { CK(pd->dbg_sub.reserve((size_t)full.n_samples * 4)); CK(cudaMemsetAsync(pd->dbg_sub.p, 0, (size_t)full.n_samples * 4, pd->stream)); ddbg.sub_id = static_cast<int32_t*>(pd->dbg_sub.p); }
            if (dbg->t) { CK(pd->dbg_t.reserve((size_t)full.n_samples * 8)); CK(cudaMemsetAsync(pd->dbg_t.p, 0, (size_t)full.n_samples * 8, pd->stream)); ddbg.t = static_cast<double*>(pd->dbg_t.p); }
        }
        // Shards of the other devices: the render kernel's fold stores finished pixels straight into a buffer on the
        // primary device over NVLink (peer access), so there is no gather step; without peer access the shard is
        // rendered locally and every finished band is copied with cudaMemcpyPeerAsync.
        S.target = pd->tiles.p;
        S.direct = false;
        if (k > 0) {
            PerDevice* p0 = sh[0].pd;
            if ((int)p0->peer_tiles.size() < n_gpus) p0->peer_tiles.resize((size_t)n_gpus);
            CK(cudaSetDevice(primary));
            CK(p0->peer_tiles[(size_t)k].reserve(tileBytes));
            CK(cudaSetDevice(dev));
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev, primary) == cudaSuccess && can) {
                cudaError_t pe = cudaDeviceEnablePeerAccess(primary, 0);
                if (pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled) S.direct = true;
                (void)cudaGetLastError();
            }
            static const bool noP2P = std::getenv("FTB_NO_P2P") != nullptr;  // A/B switch
            if (noP2P) S.direct = false;
            if (S.direct) S.target = p0->peer_tiles[(size_t)k].p;
            bufs[(size_t)k] = p0->peer_tiles[(size_t)k].p;
        } else bufs[0] = pd->tiles.p;
        ChunkDone bandDone = [&, k, dev](int c) -> int {
            Shard& Sk = sh[(size_t)k];
            if (k > 0 && !Sk.direct) {  // the band's tiles are a contiguous run of local tiles (tile rows grow with the local index)
                const size_t l0 = kBands > 1 ? (size_t)Sk.pd->chunk_first[(size_t)c] : 0, l1 = kBands > 1 ? (size_t)Sk.pd->chunk_first[(size_t)c + 1] : (size_t)Sk.g.n_local_tiles;
                const size_t off = l0 * FTB_TILE_PIXELS * 3 * rs, n = (l1 - l0) * FTB_TILE_PIXELS * 3 * rs;
                if (n) CK(cudaMemcpyPeerAsync(static_cast<char*>(sh[0].pd->peer_tiles[(size_t)k].p) + off, primary, static_cast<const char*>(Sk.pd->tiles.p) + off, dev, n, Sk.pd->stream));
                if (stats) stats->kernel_launches += 1;
            }
            CK(cudaEventRecord(Sk.pd->band_events[(size_t)c], Sk.pd->stream));
            return FTB_OK;
        };
        if ((rc = launchFrameAny(scene, pd, camera, &S.ps, S.g, S.target, ddbg.prim_id ? &ddbg : nullptr, stats, pd->stream, stats != nullptr, kBands, &bandDone)) != FTB_OK) return rc;
    }
    // ---- assemble + download on the primary device, band by band, behind the rendering ------------------------------------
    CK(cudaSetDevice(primary));
    PerDevice* p0 = sh[0].pd;
    CK(p0->out.reserve(frameBytes));
    ftb_render_params pa = p;
    pa.shard_index = 0; pa.shard_count = n_gpus;
    FrameGeom ga;
    if ((rc = frameGeom(&pa, ga)) != FTB_OK) return rc;
    for (int c = 0; c < kBands; ++c) {
        const int r0 = bandFirstRow(c, kBands, full.tiles_y), r1 = bandFirstRow(c + 1, kBands, full.tiles_y);
        const int y0 = kBands > 1 ? std::min(p.height, r0 * FTB_TILE_H) : 0, y1 = kBands > 1 ? std::min(p.height, r1 * FTB_TILE_H) : p.height;
        for (int k = 0; k < n_gpus; ++k) CK(cudaStreamWaitEvent(p0->copy_stream, sh[(size_t)k].pd->band_events[(size_t)c], 0));
        if (y1 <= y0) continue;
        if ((rc = launchAssemble(&pa, ga, bufs.data(), p0->out.p, p0->copy_stream, y0, y1)) != FTB_OK) return rc;
        if (stats) stats->kernel_launches += 1;
        CK(p0->copier.begin(static_cast<const char*>(p0->out.p) + (size_t)y0 * rowBytes, static_cast<char*>(out) + (size_t)y0 * rowBytes, (size_t)(y1 - y0) * rowBytes, p0->copy_stream));
    }
    CK(p0->copier.finish());
    CK(cudaStreamSynchronize(p0->copy_stream));
    bool overflow = false;
    for (int k = 0; k < n_gpus; ++k) {
        CK(cudaSetDevice(sh[(size_t)k].pd->device));
        if ((rc = finishStats(scene, sh[(size_t)k].pd, stats, sh[(size_t)k].pd->stream, stats != nullptr, &overflow)) != FTB_OK) return rc;
    }
    CK(cudaSetDevice(primary));
    if (dbg && dbg->prim_id) {
        CK(cudaMemcpyAsync(dbg->prim_id, p0->dbg_prim.p, (size_t)full.n_samples * 4, cudaMemcpyDeviceToHost, p0->stream));
        if (dbg->sub_id) CK(cudaMemcpyAsync(dbg->sub_id, p0->dbg_sub.p, (size_t)full.n_samples * 4, cudaMemcpyDeviceToHost, p0->stream));
        if (dbg->t) CK(cudaMemcpyAsync(dbg->t, p0->dbg_t.p, (size_t)full.n_samples * 8, cudaMemcpyDeviceToHost, p0->stream));
    }
    CK(cudaStreamSynchronize(p0->stream));
    if (stats) stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (overflow) return fail(FTB_ERR_HIT_OVERFLOW, "a CSG operand produced more than 32 crossings on one ray");
    return FTB_OK;
}

int ftb_shade_rays(ftb_scene* scene, const double* rays_od, int64_t n, const ftb_render_params* params, double* out_rgb,
                   const ftb_debug_out* dbg, ftb_stats* stats)
{
    const auto t0 = std::chrono::steady_clock::now();
    if (!scene || !params || (n > 0 && (!rays_od || !out_rgb)) || n < 0) return fail(FTB_ERR_BAD_ARG, "null argument");
    if (params->precision != FTB_PRECISION_FP32 && params->precision != FTB_PRECISION_FP64_VERIFY) return fail(FTB_ERR_BAD_ARG, "bad precision");
    if (n > (int64_t)FTB_TILE_PIXELS * 0x7fffff00LL / 2) return fail(FTB_ERR_BAD_ARG, "too many rays");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { (void)cudaGetLastError(); return fail(FTB_ERR_NO_DEVICE, "no CUDA device: functracer_b200 has no CPU path"); }
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n == 0) return FTB_OK;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    PerDevice* pd = nullptr;
    int rc = getDevice(scene, dev, &pd);
    if (rc != FTB_OK) return rc;
    const bool f64 = params->precision == FTB_PRECISION_FP64_VERIFY;
    const size_t rs = f64 ? 8 : 4;
    cudaStream_t s = pd->stream;
    CK(pd->rays.reserve((size_t)n * 48));
    CK(cudaMemcpyAsync(pd->rays.p, rays_od, (size_t)n * 48, cudaMemcpyHostToDevice, s));
    CK(pd->tiles.reserve((size_t)n * 3 * rs));
    CK(pd->out.reserve((size_t)n * 24));
    CK(pd->control.reserve(sizeof(Control)));
    CK(cudaMemsetAsync(pd->control.p, 0, sizeof(Control), s));
    ftb_debug_out ddbg = {nullptr, nullptr, nullptr};
    if (dbg && dbg->prim_id) {
        CK(pd->dbg_prim.reserve((size_t)n * 4)); ddbg.prim_id = static_cast<int32_t*>(pd->dbg_prim.p);
        if (dbg->sub_id) { CK(pd->dbg_sub.reserve((size_t)n * 4)); ddbg.sub_id = static_cast<int32_t*>(pd->dbg_sub.p); }
        if (dbg->t) { CK(pd->dbg_t.reserve((size_t)n * 8)); ddbg.t = static_cast<double*>(pd->dbg_t.p); }
    }
    Control* ctl = static_cast<Control*>(pd->control.p);
    const bool wantStats = stats && params->collect_stats;
    int launches = 0;
    auto run = [&](auto tag) -> int {
        typedef decltype(tag) R;
        SceneStorage<R>& st = storageOf<R>(pd);
        if (!st.ready) { int r2 = uploadScene<R>(*scene, st); if (r2 != FTB_OK) return r2; }
        DevFrame<R> F;
        std::memset(&F, 0, sizeof(F));
        F.mode = 1; F.spp = 1; F.s_base = 0; F.s_count = 1; F.run = 1; F.rpp_magic = 0;
        F.n_blocks = (int)((n + 31) / 32);
        F.n_rays = n;
        F.n_local_tiles = (int)((n + FTB_TILE_PIXELS - 1) / FTB_TILE_PIXELS);
        F.shard_count = 1;
        F.rays = static_cast<const double*>(pd->rays.p);
        F.recursion_limit = params->recursion_limit;
        F.seed = params->seed;
        F.out = static_cast<R*>(pd->tiles.p);
        F.dbg_prim = ddbg.prim_id; F.dbg_sub = ddbg.sub_id; F.dbg_t = ddbg.t;
        F.tile_counter = &ctl->tile_counter; F.overflow = &ctl->overflow; F.stats = ctl->stats;
        CK(cudaEventRecord(pd->ev0, s));
        const Variant<R>* var = pickVariant<R>(scene->L.features, wantStats);
        if (!var) return fail(FTB_ERR_UNSUPPORTED, "no kernel variant covers this scene's feature mask");
        CK(var->launch(st.view, F, wantStats, pd->sm_count, s, &launches));
        CK(cudaEventRecord(pd->ev1, s));
        const long long m = 3 * (long long)n;
        widen_kernel<R><<<(int)std::min<long long>((m + 255) / 256, 148 * 16), 256, 0, s>>>(static_cast<const R*>(pd->tiles.p), static_cast<double*>(pd->out.p), m);
        CK(cudaGetLastError());
        return FTB_OK;
    };
    rc = f64 ? run(double()) : run(float());
    if (rc != FTB_OK) return rc;
    CK(cudaMemcpyAsync(out_rgb, pd->out.p, (size_t)n * 24, cudaMemcpyDeviceToHost, s));
    if (ddbg.prim_id) {
        CK(cudaMemcpyAsync(dbg->prim_id, ddbg.prim_id, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (ddbg.sub_id) CK(cudaMemcpyAsync(dbg->sub_id, ddbg.sub_id, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (ddbg.t) CK(cudaMemcpyAsync(dbg->t, ddbg.t, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    }
    bool overflow = false;
    if (stats) stats->kernel_launches = launches + 2;
    if ((rc = finishStats(scene, pd, stats, s, true, &overflow)) != FTB_OK) return rc;
    if (stats) stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (overflow) return fail(FTB_ERR_HIT_OVERFLOW, "a CSG operand produced more than 32 crossings on one ray");
    return FTB_OK;
}

}  // extern "C"
