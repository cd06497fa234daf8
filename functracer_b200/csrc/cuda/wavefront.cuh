// wavefront.cuh — the render loop in the formulation BASELINE.json's north_star prescribes, as the A/B arm of render.cuh's
// persistent megakernel (selected at run time with FTB_WAVEFRONT=1; DESIGN.md §8, BENCH.md §4).
//
// The same device functions (traceScene, finalise, shadeLight, primaryRay, jitterVector) in the same order per sample - so a frame
// equals the megakernel's up to the last bits of FP32 (primary hits identical; the compiler contracts multiply-adds differently
// in different kernels, and the megakernel folds long sample runs in segments) - but split into one kernel per stage, with every path's state and every ray living in
// structure-of-arrays records in HBM between the stages:
//   generate   primary rays of a wave of samples (all samples of a run of tiles)          -> path records
//   nearest    nearest-hit query of every live path (Scene.closest, Scene.fs:112-118)      -> hit records
//   setup      winner finalisation + one shadow ray per light [x soft sample]             -> shadow-ray records (fixed stride per path)
//   shadow     any-hit query of every shadow ray (Scene.lightIsBocked, Scene.fs:119-121)   -> blocked flags
//   accumulate per-light shading with the shadow results, running weight W <- W L r,
//              reflection ray of the next generation; live paths are compacted with ballot + popc + one atomic per warp
//   blend      per pixel, samples folded in sample order (Array.average, Image.fs:112-116) -> the tile-major frame
// F#'s recursion (getColourForRay, limit 8) is the loop over generations on the host.
#pragma once

#include <algorithm>
#include <utility>

#include "render.cuh"

namespace ftb {

constexpr int kWfThreads = 128;

template <typename R>
struct WfState {
    typedef typename V4<R>::type R4;
    // per path of the wave
    R4* ro;        // path ray origin (un-offset), w = running weight
    R4* rd;        // path ray direction, w unused
    R4* acc;       // the sample's colour so far, w = bounces left (as a real)
    int* planar;   // planar leaf the current ray leaves (FP32 self-intersection guard), else -1
    R4* hit;       // t, leaf, sub, flip (ints as bits / reals)
    int* listA;    // live paths of this generation
    int* listB;    // live paths of the next
    unsigned* counts;  // [0] live paths of this generation, [1] of the next
    // per shadow ray: stride = rays_per_hit entries per live path
    R4* so;        // origin, w = tmax
    R4* sd;        // direction, w unused
    int* sinfo;    // light | sample << 8 | (skipLeaf + 1) << 16, or -1: no ray needed
    int* sres;     // 1 = blocked
    R* col;        // finished sample colours [path][3]
    int rays_per_hit;
    // the wave
    int tile_first, tile_count;  // positions in the (ordered) local tile list
    long long n_paths;           // tile_count * 256 * spp
};

template <typename R>
FTB_DEV int asInt(R v);
template <>
FTB_DEV int asInt<float>(float v) { return __float_as_int(v); }
template <>
FTB_DEV int asInt<double>(double v) { return (int)v; }
template <typename R>
FTB_DEV R fromInt(int v);
template <>
FTB_DEV float fromInt<float>(int v) { return __int_as_float(v); }
template <>
FTB_DEV double fromInt<double>(int v) { return (double)v; }

// path id of a wave -> pixel / sample (false: a padding pixel of an edge tile)
template <typename R>
FTB_DEV bool wfLocate(const DevFrame<R>& F, const WfState<R>& W, long long pid, int& px, int& py, int& sj, int& slot)
{
    const int spp = F.spp;
    const long long pixel = pid / spp;
    sj = (int)(pid - pixel * spp);
    const int k = (int)(pixel / FTB_TILE_PIXELS), pix = (int)(pixel - (long long)k * FTB_TILE_PIXELS);
    const int ltile = F.tile_order ? __ldg(F.tile_order + W.tile_first + k) : W.tile_first + k;
    const int tile = tileOfLocal(ltile, F.shard_index, F.shard_count);
    const int ty = tile / F.tiles_x, tx = tile - ty * F.tiles_x;
    px = tx * FTB_TILE_W + (pix & (FTB_TILE_W - 1));
    py = ty * FTB_TILE_H + (pix / FTB_TILE_W);
    slot = ltile * FTB_TILE_PIXELS + pix;
    return px < F.gw && py < F.gh;
}

// the common-origin bound table of render.cuh's prologue, one copy per CTA (see the derivation there)
template <typename R, unsigned FEAT>
FTB_DEV bool wfOriginTable(const DevScene<R>& S, const DevFrame<R>& F, typename V4<R>::type* origin_tab)
{
    const bool fastBounds = (FEAT & FT_TABLE) != 0 && originTableFits(S);
    if (fastBounds) buildOriginTable<R>(S, F, origin_tab);
    return fastBounds;
}

// live-path compaction: ballot + popc (warp scan) and one atomic per warp
FTB_DEV void wfAppend(bool take, int value, int* list, unsigned* counter)
{
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (take) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

template <typename R, unsigned FEAT>
__global__ void __launch_bounds__(kWfThreads) wf_generate(const __grid_constant__ DevScene<R> S, const __grid_constant__ DevFrame<R> F, const __grid_constant__ WfState<R> W)
{
    typedef typename V4<R>::type R4;
    const long long n = (W.n_paths + 31) / 32 * 32;
    for (long long pid = blockIdx.x * (long long)blockDim.x + threadIdx.x; pid < n; pid += (long long)gridDim.x * blockDim.x) {
        bool live = false;
        if (pid < W.n_paths) {
            int px, py, sj, slot;
            live = wfLocate(F, W, pid, px, py, sj, slot);
            if (live) {
                const unsigned long long sampleIndex = ((unsigned long long)py * (unsigned)F.gw + (unsigned)px) * (unsigned)F.spp + (unsigned)sj;
                const Ray<R> ray = primaryRay<R, FEAT>(F, px, py, F.mode == 0 && F.spp > 0 ? sj : 0, sampleIndex);
                R4 o, d, a;
                o.x = ray.o.x; o.y = ray.o.y; o.z = ray.o.z; o.w = R(1);
                d.x = ray.d.x; d.y = ray.d.y; d.z = ray.d.z; d.w = R(0);
                a.x = a.y = a.z = R(0); a.w = (R)F.recursion_limit;
                W.ro[pid] = o; W.rd[pid] = d; W.acc[pid] = a;
                W.planar[pid] = -1;
            }
        }
        wfAppend(live, (int)pid, W.listA, W.counts);
    }
}

template <typename R, unsigned FEAT>
__global__ void __launch_bounds__(kWfThreads) wf_nearest(const __grid_constant__ DevScene<R> S, const __grid_constant__ DevFrame<R> F, const __grid_constant__ WfState<R> W, const int* list,
                                                          const unsigned* count, int generation)
{
    typedef typename V4<R>::type R4;
    __shared__ R4 origin_tab[(FEAT & FT_TABLE) != 0 ? kOriginCap : 1];
    const bool fastBounds = wfOriginTable<R, FEAT>(S, F, origin_tab);
    const bool fastPrimary = fastBounds && generation == 0 && !((FEAT & FT_RNG) != 0 && F.has_focus);
    const unsigned m = *count;
    Counters<false> cn;
    bool overflow = false;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const int pid = list[i];
        const R4 o = W.ro[pid], d = W.rd[pid];
        Ray<R> tr;
        tr.d = mk<R>(d.x, d.y, d.z);
        tr.o = mk<R>(o.x, o.y, o.z) + R(0.0001) * tr.d;  // slightOffset (Shading.fs:129)
        const int skipLeaf = generation > 0 ? W.planar[pid] : -1;
        const HitInfo<R> h = traceScene<R, FEAT, false>(S, tr, inf_<R>(), false, skipLeaf, fastPrimary ? origin_tab : nullptr, F.primary_slack, overflow, cn, 0xffffffffu, nullptr);
        R4 rec;
        rec.x = h.t; rec.y = fromInt<R>(h.leaf); rec.z = fromInt<R>(h.sub); rec.w = fromInt<R>(h.flip);
        W.hit[pid] = rec;
        if (generation == 0 && F.dbg_prim) {
            int px, py, sj, slot;
            wfLocate(F, W, pid, px, py, sj, slot);
            int prim = -1, sub = 0;
            if (h.leaf >= 0) {
                const int4 meta = __ldg(S.leaf_meta + h.leaf);
                const int kind = meta.x & 0xff;
                prim = meta.z;
                sub = (kind == LEAF_CUBE || kind == LEAF_MESH || kind == LEAF_SOLIDCYL) ? h.sub : ((kind == LEAF_TRIANGLE) ? 0 : meta.w);
            }
            const unsigned long long at = ((unsigned long long)py * (unsigned)F.gw + (unsigned)px) * (unsigned)F.spp + (unsigned)sj;
            F.dbg_prim[at] = prim;
            if (F.dbg_sub) F.dbg_sub[at] = sub;
            if (F.dbg_t) F.dbg_t[at] = h.leaf >= 0 ? (double)h.t : -1.0;
        }
    }
    if (overflow) atomicExch(F.overflow, 1u);
}

// the hit of a path back from its record
template <typename R>
FTB_DEV HitInfo<R> wfHit(const WfState<R>& W, int pid)
{
    const typename V4<R>::type rec = W.hit[pid];
    HitInfo<R> h;
    h.t = rec.x; h.leaf = asInt<R>(rec.y); h.sub = asInt<R>(rec.z); h.flip = asInt<R>(rec.w);
    return h;
}

template <typename R, unsigned FEAT>
__global__ void __launch_bounds__(kWfThreads) wf_setup(const __grid_constant__ DevScene<R> S, const __grid_constant__ DevFrame<R> F, const __grid_constant__ WfState<R> W, const int* list,
                                                        const unsigned* count, int generation)
{
    typedef typename V4<R>::type R4;
    const unsigned m = *count;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const int pid = list[i];
        const HitInfo<R> h = wfHit(W, pid);
        const size_t base = (size_t)i * W.rays_per_hit;
        int k = 0;
        if (h.leaf >= 0 && S.n_lights > 0) {
            const R4 o = W.ro[pid], d = W.rd[pid];
            Ray<R> tr;
            tr.d = mk<R>(d.x, d.y, d.z);
            tr.o = mk<R>(o.x, o.y, o.z) + R(0.0001) * tr.d;
            const Fragment<R> f = finalise<R, FEAT>(S, tr, h);
            const Vec<R> origin = f.p + R(0.0001) * f.n;  // shadowRayOrigin (Shading.fs:111)
            bool needShadow = f.applyLighting;
            if constexpr ((FEAT & FT_ROUGH) != 0) needShadow = needShadow && !(f.roughness != R(0) && !(f.shineyness > R(0)));
            unsigned long long sampleIndex = 0;
            if constexpr ((FEAT & FT_RNG) != 0) {
                int px, py, sj, slot;
                wfLocate(F, W, pid, px, py, sj, slot);
                sampleIndex = ((unsigned long long)py * (unsigned)F.gw + (unsigned)px) * (unsigned)F.spp + (unsigned)sj;
            }
            for (int li = 0; li < S.n_lights; ++li) {
                const int2 lk = __ldg(S.light_i + li);
                const R4 la = ldg4<R>(S.light_a + li);
                const Vec<R> lv = mk<R>(la.x, la.y, la.z);
                const int nk = lk.x == FTB_LIGHT_SOFT_DIRECTIONAL ? max(lk.y, 0) : 1;
                for (int sk = 0; sk < nk; ++sk, ++k) {
                    int info = -1;
                    R4 so, sd;
                    so.x = so.y = so.z = so.w = R(0); sd = so;
                    if (needShadow) {
                        Vec<R> dir;
                        R tmax;
                        if (lk.x == FTB_LIGHT_POINT) {  // shadowLightIntensity (Shading.fs:33-42)
                            const Vec<R> dvec = lv - origin;
                            tmax = length(dvec);
                            dir = normalise(dvec);
                        } else if (lk.x == FTB_LIGHT_SOFT_DIRECTIONAL) {  // softShadowLightIntensity (Shading.fs:24-31)
                            tmax = realmax_<R>();
                            dir = -lv;
                            if constexpr ((FEAT & FT_RNG) != 0)
                                dir = jitterVector<R>(F.seed, sampleIndex, (unsigned)generation, (unsigned)li, (unsigned)sk, la.w, -lv);
                        } else {
                            tmax = realmax_<R>();
                            dir = -lv;
                        }
                        int skip = -1;
                        if constexpr (sizeof(R) == 4 && (FEAT & FT_PLANAR) != 0) { if (f.planarLeaf >= 0 && dot(dir, f.n) >= R(0)) skip = f.planarLeaf; }
                        so.x = origin.x; so.y = origin.y; so.z = origin.z; so.w = tmax;
                        sd.x = dir.x; sd.y = dir.y; sd.z = dir.z;
                        info = li | (sk << 8) | ((skip + 1) << 16);
                    }
                    W.so[base + k] = so; W.sd[base + k] = sd; W.sinfo[base + k] = info;
                }
            }
        }
        for (; k < W.rays_per_hit; ++k) W.sinfo[base + k] = -1;
    }
}

template <typename R, unsigned FEAT>
__global__ void __launch_bounds__(kWfThreads) wf_shadow(const __grid_constant__ DevScene<R> S, const __grid_constant__ DevFrame<R> F, const __grid_constant__ WfState<R> W, const unsigned* count)
{
    typedef typename V4<R>::type R4;
    __shared__ R4 origin_tab[(FEAT & FT_TABLE) != 0 ? kOriginCap : 1];
    const bool fastBounds = wfOriginTable<R, FEAT>(S, F, origin_tab);
    const size_t n = ((size_t)*count * W.rays_per_hit + 31) / 32 * 32;
    Counters<false> cn;
    bool overflow = false;
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        const int info = j < (size_t)*count * W.rays_per_hit ? W.sinfo[j] : -1;
        const bool on = info >= 0;
        R4 so, sd;
        so.x = so.y = so.z = so.w = R(0); sd = so; sd.x = R(1);
        if (on) { so = W.so[j]; sd = W.sd[j]; }
        // every tracing lane has a row in the common-origin table?  (rays towards point lights have a finite tmax)
        const bool tabled = fastBounds && !__any_sync(0xffffffffu, on && !(so.w < realmax_<R>()));
        if (on) {
            Ray<R> r;
            r.o = mk<R>(so.x, so.y, so.z); r.d = mk<R>(sd.x, sd.y, sd.z);
            const int li = info & 0xff, skip = ((info >> 16) & 0x7fff) - 1;
            const HitInfo<R> h = traceScene<R, FEAT, false>(S, r, so.w, true, skip, tabled ? origin_tab + (1 + li) * tabStride(S.n_items) : nullptr, R(4e-4) * so.w, overflow, cn, 0xffffffffu, nullptr);
            W.sres[j] = h.leaf >= 0 ? 1 : 0;
        }
    }
    if (overflow) atomicExch(F.overflow, 1u);
}

template <typename R, unsigned FEAT>
__global__ void __launch_bounds__(kWfThreads) wf_accumulate(const __grid_constant__ DevScene<R> S, const __grid_constant__ DevFrame<R> F, const __grid_constant__ WfState<R> W, const int* list,
                                                             const unsigned* count, int* next, unsigned* nextCount)
{
    typedef typename V4<R>::type R4;
    const unsigned m = *count;
    const unsigned padded = (m + 31u) / 32u * 32u;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += gridDim.x * blockDim.x) {
        bool goesOn = false;
        int pid = 0;
        if (i < m) {
            pid = list[i];
            const HitInfo<R> h = wfHit(W, pid);
            const R4 o = W.ro[pid], d = W.rd[pid];
            R4 a = W.acc[pid];
            Vec<R> scol = mk<R>(a.x, a.y, a.z);
            const int limit = (int)a.w;
            if (h.leaf >= 0 && S.n_lights > 0) {
                Ray<R> tr;
                tr.d = mk<R>(d.x, d.y, d.z);
                tr.o = mk<R>(o.x, o.y, o.z) + R(0.0001) * tr.d;
                const Fragment<R> f = finalise<R, FEAT>(S, tr, h);
                Vec<R> local = mk<R>(R(0), R(0), R(0));
                const size_t base = (size_t)i * W.rays_per_hit;
                int k = 0;
                bool needShadow = f.applyLighting;  // as in wf_setup: fragments whose colour ignores the light's intensity trace no shadow ray
                if constexpr ((FEAT & FT_ROUGH) != 0) needShadow = needShadow && !(f.roughness != R(0) && !(f.shineyness > R(0)));
                for (int li = 0; li < S.n_lights; ++li) {
                    const int2 lk = __ldg(S.light_i + li);
                    R intensity = R(1);
                    const int nk = lk.x == FTB_LIGHT_SOFT_DIRECTIONAL ? max(lk.y, 0) : 1;
                    if (needShadow) {
                        if (lk.x == FTB_LIGHT_SOFT_DIRECTIONAL) {  // softShadowLightIntensity (Shading.fs:24-31)
                            int occluded = 0;
                            for (int sk = 0; sk < nk; ++sk) occluded += W.sres[base + k + sk];
                            intensity = (R)(lk.y - occluded) / (R)lk.y;  // 0 samples: 0 / 0 = NaN like the reference
                        } else {
                            const bool blocked = W.sres[base + k] != 0;
                            if (lk.x == FTB_LIGHT_POINT) {
                                const R4 lb = ldg4<R>(S.light_b + li);
                                const R tmax = W.so[base + k].w;
                                intensity = blocked ? R(0) : R(1) / (lb.x + tmax * (lb.y + tmax * lb.z));  // Light.attenuate (Light.fs:16-17)
                            } else intensity = blocked ? R(0) : R(1);
                        }
                    }
                    local = local + shadeLight<R, FEAT>(S, f, tr.d, li, intensity);
                    k += nk;
                }
                scol = scol + o.w * local;
                if (f.applyLighting && f.reflectance > R(0) && limit > 0) {  // reflectionShader (Shading.fs:89-98): weight L * reflectance
                    R4 no, nd;
                    const Vec<R> rdir = reflect(f.n, tr.d);
                    no.x = f.p.x; no.y = f.p.y; no.z = f.p.z; no.w = o.w * ((R)S.n_lights * f.reflectance);
                    nd.x = rdir.x; nd.y = rdir.y; nd.z = rdir.z; nd.w = R(0);
                    W.ro[pid] = no; W.rd[pid] = nd;
                    W.planar[pid] = f.planarLeaf;
                    a.x = scol.x; a.y = scol.y; a.z = scol.z; a.w = (R)(limit - 1);
                    W.acc[pid] = a;
                    goesOn = true;
                }
            }
            if (!goesOn) { W.col[3 * (size_t)pid] = scol.x; W.col[3 * (size_t)pid + 1] = scol.y; W.col[3 * (size_t)pid + 2] = scol.z; }
        }
        wfAppend(goesOn, pid, next, nextCount);
    }
}

// Array.average over each pixel's samples, in sample order (Image.fs:112-116): the same left fold as render.cuh's foldUnit
template <typename R>
__global__ void wf_blend(const __grid_constant__ DevFrame<R> F, const __grid_constant__ WfState<R> W)
{
    const long long n = (long long)W.tile_count * FTB_TILE_PIXELS * 3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long pixel = i / 3;
        const int ch = (int)(i - 3 * pixel);
        int px, py, sj, slot;
        if (!wfLocate(F, W, pixel * F.spp, px, py, sj, slot)) continue;
        R acc = R(0);
        const R* c = W.col + 3 * (size_t)(pixel * F.spp) + ch;
        for (int q = 0; q < F.spp; ++q) acc = acc + c[3 * (size_t)q];
        acc = acc / (R)F.spp;
        F.out[3 * (long long)slot + ch] = acc;
    }
}

// Host side: waves of tiles; per wave one generation per bounce.  scratch: one allocation the caller keeps between frames.
template <typename R, unsigned FEAT>
cudaError_t launch_wavefront_impl(const DevScene<R>& s, const DevFrame<R>& f, int n_tiles, int rays_per_hit, bool reflective, int sm_count, void* scratch, size_t scratch_bytes,
                                  cudaStream_t stream, int* launches)
{
    typedef typename V4<R>::type R4;
    if (f.mode != 0 || f.spp < 1) return cudaErrorInvalidValue;
    const size_t per_path = 4 * sizeof(R4) + 3 * sizeof(int) + 3 * sizeof(R) + (size_t)rays_per_hit * (2 * sizeof(R4) + 2 * sizeof(int));
    const long long per_tile = (long long)FTB_TILE_PIXELS * f.spp;
    long long tiles_per_wave = scratch_bytes > 16384 ? (long long)((scratch_bytes - 16384) / per_path) / per_tile : 0;
    if (tiles_per_wave < 1) return cudaErrorMemoryAllocation;
    if (tiles_per_wave > n_tiles) tiles_per_wave = n_tiles;
    const size_t P = (size_t)(tiles_per_wave * per_tile);
    WfState<R> W;
    char* p = static_cast<char*>(scratch);
    auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) & ~(size_t)255; return q; };
    W.counts = reinterpret_cast<unsigned*>(take(256));
    W.ro = reinterpret_cast<R4*>(take(P * sizeof(R4))); W.rd = reinterpret_cast<R4*>(take(P * sizeof(R4)));
    W.acc = reinterpret_cast<R4*>(take(P * sizeof(R4))); W.hit = reinterpret_cast<R4*>(take(P * sizeof(R4)));
    W.planar = reinterpret_cast<int*>(take(P * sizeof(int)));
    W.listA = reinterpret_cast<int*>(take(P * sizeof(int))); W.listB = reinterpret_cast<int*>(take(P * sizeof(int)));
    W.col = reinterpret_cast<R*>(take(3 * P * sizeof(R)));
    W.so = reinterpret_cast<R4*>(take(P * rays_per_hit * sizeof(R4))); W.sd = reinterpret_cast<R4*>(take(P * rays_per_hit * sizeof(R4)));
    W.sinfo = reinterpret_cast<int*>(take(P * rays_per_hit * sizeof(int))); W.sres = reinterpret_cast<int*>(take(P * rays_per_hit * sizeof(int)));
    W.rays_per_hit = rays_per_hit;
    if ((size_t)(p - static_cast<char*>(scratch)) > scratch_bytes) return cudaErrorMemoryAllocation;
    const int grid = sm_count * 8;
    cudaError_t e;
    for (int t0 = 0; t0 < n_tiles; t0 += (int)tiles_per_wave) {
        W.tile_first = t0;
        W.tile_count = (int)std::min<long long>(tiles_per_wave, n_tiles - t0);
        W.n_paths = (long long)W.tile_count * per_tile;
        if ((e = cudaMemsetAsync(W.counts, 0, 256, stream)) != cudaSuccess) return e;
        wf_generate<R, FEAT><<<grid, kWfThreads, 0, stream>>>(s, f, W);
        int* cur = W.listA;
        int* nxt = W.listB;
        for (int gen = 0; gen <= f.recursion_limit; ++gen) {
            unsigned* cnt = W.counts + (gen & 1);
            unsigned* ncnt = W.counts + ((gen + 1) & 1);
            wf_nearest<R, FEAT><<<grid, kWfThreads, 0, stream>>>(s, f, W, cur, cnt, gen);
            wf_setup<R, FEAT><<<grid, kWfThreads, 0, stream>>>(s, f, W, cur, cnt, gen);
            wf_shadow<R, FEAT><<<grid, kWfThreads, 0, stream>>>(s, f, W, cnt);
            if ((e = cudaMemsetAsync(ncnt, 0, sizeof(unsigned), stream)) != cudaSuccess) return e;
            wf_accumulate<R, FEAT><<<grid, kWfThreads, 0, stream>>>(s, f, W, cur, cnt, nxt, ncnt);
            if (launches) *launches += 4;
            if (!reflective) break;  // no surface reflects: every path ends in its first generation
            unsigned live = 0;
            if ((e = cudaMemcpyAsync(&live, ncnt, sizeof(unsigned), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            if (live == 0) break;
            std::swap(cur, nxt);
        }
        wf_blend<R><<<grid, 256, 0, stream>>>(f, W);
        if (launches) *launches += 2;
    }
    return cudaGetLastError();
}

}  // namespace ftb
