// bvh_build.cu — see bvh_build.h.  PLOC (parallel locally-ordered clustering) over Morton-sorted triangles.
#include "bvh_build.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cub/cub.cuh>

namespace ftb {
namespace {

constexpr int kLeafMax = 4;   // triangles per leaf
constexpr int kThreads = 256;

// floats <-> unsigned ints with the same order (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned orderedBits(float f)
{
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
inline float fromOrderedHost(unsigned u)
{
    const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}

struct Arrays {
    const double* tri9;  // the scene's triangles, 9 doubles each
    const int* order;    // this mesh's triangles in the reference's enumeration order
    int n;
    float4 *plo, *phi;   // per triangle (enumeration order): box, rounded outward from the double vertices
    unsigned* cbound;    // [0..2] min, [3..5] max of the box centres (ordered bits)
    unsigned long long *keys, *keys2;
    int *vals, *vals2;
    // the binary tree: nodes 0..n-1 = the sorted triangles, n.. = merges
    float4 *nlo, *nhi;
    int *left, *right, *parent, *count;
    int *ref, *seqv;     // per leaf node: triangle index, enumeration rank
    int *clusterA, *clusterB, *nn, *merged, *valid, *pos;
    int* counters;       // 0: next tree node, 1: clusters after compaction, 2: out nodes, 3: out slots, 4: max depth
    // output
    float4 *olo0, *ohi0, *olo1, *ohi1;
    int2* ochild;
    int *outIndex, *leafCode;  // per tree node
    int *slotTri, *slotSeq;
};

__global__ void primKernel(Arrays A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {0, 0, 0};
    const bool on = i < A.n;
    if (on) {
        const double* q = A.tri9 + 9 * (size_t)A.order[i];
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            const double l = fmin(q[k], fmin(q[3 + k], q[6 + k])), h = fmax(q[k], fmax(q[3 + k], q[6 + k]));
            lo[k] = __double2float_rd(l);  // rounded outward: the box may only grow
            hi[k] = __double2float_ru(h);
            c[k] = 0.5f * lo[k] + 0.5f * hi[k];
        }
        A.plo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        A.phi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    const unsigned mask = __activemask();
    for (int k = 0; k < 3; ++k) {
        const unsigned lo = __reduce_min_sync(mask, on ? orderedBits(c[k]) : 0xffffffffu);
        const unsigned hi = __reduce_max_sync(mask, on ? orderedBits(c[k]) : 0u);
        if ((threadIdx.x & 31) == 0) { atomicMin(&A.cbound[k], lo); atomicMax(&A.cbound[3 + k], hi); }
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned v)  // 21 bits -> every third bit of 63
{
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void mortonKernel(Arrays A, float3 cmin, float3 scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const float4 lo = A.plo[i], hi = A.phi[i];
    const float cx = 0.5f * lo.x + 0.5f * hi.x, cy = 0.5f * lo.y + 0.5f * hi.y, cz = 0.5f * lo.z + 0.5f * hi.z;
    const float m = 2097151.0f;
    const unsigned x = (unsigned)fminf(fmaxf((cx - cmin.x) * scale.x, 0.f), m), y = (unsigned)fminf(fmaxf((cy - cmin.y) * scale.y, 0.f), m),
                   z = (unsigned)fminf(fmaxf((cz - cmin.z) * scale.z, 0.f), m);
    A.keys[i] = spread21(x) << 2 | spread21(y) << 1 | spread21(z);
    A.vals[i] = i;
}

__global__ void leafKernel(Arrays A)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n) return;
    const int src = A.vals2[p];
    A.nlo[p] = A.plo[src]; A.nhi[p] = A.phi[src];
    A.ref[p] = A.order[src]; A.seqv[p] = src;
    A.count[p] = 1; A.left[p] = -1; A.right[p] = -1; A.parent[p] = -1;
    A.clusterA[p] = p;
}

__device__ __forceinline__ float unionArea(float4 alo, float4 ahi, float4 blo, float4 bhi)
{
    const float ex = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), ey = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), ez = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return ex * ey + ey * ez + ez * ex;
}

// nearest neighbour of every cluster within the window: smallest joint surface area, the smaller position on ties
__global__ void nnKernel(Arrays A, const int* C, int m, int radius)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int a = C[i];
    const float4 alo = A.nlo[a], ahi = A.nhi[a];
    float best = 3.4e38f;
    int bestj = -1;
    const int j0 = max(0, i - radius), j1 = min(m - 1, i + radius);
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const int b = C[j];
        const float d = unionArea(alo, ahi, A.nlo[b], A.nhi[b]);
        if (d < best || bestj < 0) { best = d; bestj = j; }
    }
    A.nn[i] = bestj;
}

// clusters that are each other's nearest neighbour merge into a new tree node (kept at the smaller position)
__global__ void mergeKernel(Arrays A, const int* C, int m)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = A.nn[i];
    if (j >= 0 && A.nn[j] == i) {
        if (i < j) {
            const int a = C[i], b = C[j];
            const int k = atomicAdd(&A.counters[0], 1);
            A.left[k] = a; A.right[k] = b; A.parent[k] = -1;
            A.parent[a] = k; A.parent[b] = k;
            const float4 alo = A.nlo[a], ahi = A.nhi[a], blo = A.nlo[b], bhi = A.nhi[b];
            A.nlo[k] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
            A.nhi[k] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
            A.count[k] = A.count[a] + A.count[b];
            A.merged[i] = k; A.valid[i] = 1;
        } else {
            A.merged[i] = -1; A.valid[i] = 0;
        }
    } else {
        A.merged[i] = C[i]; A.valid[i] = 1;
    }
}

__global__ void compactKernel(Arrays A, int* Cnext, int m)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (A.valid[i]) Cnext[A.pos[i]] = A.merged[i];
    if (i == m - 1) A.counters[1] = A.pos[i] + A.valid[i];
}

// leaves = maximal subtrees of <= kLeafMax triangles; inner nodes get their output index, leaves their run of slots
__global__ void emitKernel(Arrays A, int nodes, int root)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nodes) return;
    const int cnt = A.count[v];
    const int par = A.parent[v];
    A.outIndex[v] = -1;
    A.leafCode[v] = 0;
    if (cnt > kLeafMax) {
        A.outIndex[v] = atomicAdd(&A.counters[2], 1);
        int d = 1;  // postponed siblings on the way down to this node's children
        for (int p = par; p >= 0; p = A.parent[p]) ++d;
        atomicMax(&A.counters[4], d);
    } else if (v == root || A.count[par] > kLeafMax) {
        const int first = atomicAdd(&A.counters[3], cnt);
        A.leafCode[v] = ~((first << 3) | cnt);
        int stack[8], sp = 0, k = 0;
        stack[sp++] = v;
        while (sp > 0) {  // <= 4 leaves below: a handful of steps
            const int u = stack[--sp];
            if (A.left[u] < 0) { A.slotTri[first + k] = A.ref[u]; A.slotSeq[first + k] = A.seqv[u]; ++k; }
            else { stack[sp++] = A.right[u]; stack[sp++] = A.left[u]; }
        }
    }
}

__global__ void nodeKernel(Arrays A, int nodes)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nodes) return;
    const int o = A.outIndex[v];
    if (o < 0) return;
    const int l = A.left[v], r = A.right[v];
    A.olo0[o] = A.nlo[l]; A.ohi0[o] = A.nhi[l];
    A.olo1[o] = A.nlo[r]; A.ohi1[o] = A.nhi[r];
    A.ochild[o] = make_int2(A.outIndex[l] >= 0 ? A.outIndex[l] : A.leafCode[l], A.outIndex[r] >= 0 ? A.outIndex[r] : A.leafCode[r]);
}

struct Arena {
    char* base = nullptr;
    size_t size = 0, used = 0;
    template <typename T>
    T* take(size_t n)
    {
        used = (used + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + used);
        used += n * sizeof(T);
        return p;
    }
};

#define BCK(expr)                                                                      \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e__); (void)cudaGetLastError(); goto fail; } \
    } while (0)

}  // namespace

bool buildMeshIndexDevice(const double* triangles, int n_triangles, const std::vector<std::vector<int32_t>>& order, int max_depth_allowed, int radius, Lowered& L,
                          DeviceBuildStats& stats, std::string& err)
{
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        static const bool verbose = std::getenv("FTB_VERBOSE") != nullptr;
        if (verbose) std::fprintf(stderr, "functracer_b200:   bvh build %-12s at %8.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    size_t n_max = 0;
    for (const auto& o : order) n_max = std::max(n_max, o.size());
    std::vector<BvhNode> outNodes;
    std::vector<int32_t> outTri, outSeq, outRoot(order.size(), ~0);
    int outDepth = 0;
    if (n_max == 0) { L.mesh_root.assign(order.size(), ~0); return true; }
    if (n_max >= (1u << 27)) { err = "mesh too large for the device build"; return false; }

    double* d_tri = nullptr;
    int* d_order = nullptr;
    char* d_arena = nullptr;
    void* d_temp = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t s = nullptr;
    {
        const size_t n = n_max, nodes = 2 * n;
        size_t sortBytes = 0, scanBytes = 0;
        Arrays A;
        std::memset(&A, 0, sizeof(A));
        Arena ar;
        float buildMs = 0;
        BCK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        BCK(cudaEventCreate(&ev0));
        BCK(cudaEventCreate(&ev1));
        BCK(cudaMallocAsync((void**)&d_tri, 9 * sizeof(double) * (size_t)n_triangles, s));  // stream-ordered: no device-wide synchronisation when freed (cudaFree measured up to 500 ms here)
        BCK(cudaMemcpyAsync(d_tri, triangles, 9 * sizeof(double) * (size_t)n_triangles, cudaMemcpyHostToDevice, s));
        BCK(cudaMallocAsync((void**)&d_order, sizeof(int) * n, s));
        BCK(cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n, 0, 63, s));
        BCK(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, (const int*)nullptr, (int*)nullptr, (int)n, s));
        BCK(cudaMallocAsync(&d_temp, std::max(sortBytes, scanBytes) + 256, s));
        ar.size = n * (2 * sizeof(float4) + 2 * sizeof(unsigned long long) + 2 * sizeof(int) + 2 * sizeof(int) + 6 * sizeof(int) + 2 * sizeof(int)) +
                  nodes * (2 * sizeof(float4) + 4 * sizeof(int) + 2 * sizeof(int)) + n * (4 * sizeof(float4) + sizeof(int2)) + 64 * 256;
        BCK(cudaMallocAsync((void**)&d_arena, ar.size, s));
        ar.base = d_arena;
        A.tri9 = d_tri; A.order = d_order;
        A.plo = ar.take<float4>(n); A.phi = ar.take<float4>(n);
        A.cbound = ar.take<unsigned>(8);
        A.keys = ar.take<unsigned long long>(n); A.keys2 = ar.take<unsigned long long>(n);
        A.vals = ar.take<int>(n); A.vals2 = ar.take<int>(n);
        A.nlo = ar.take<float4>(nodes); A.nhi = ar.take<float4>(nodes);
        A.left = ar.take<int>(nodes); A.right = ar.take<int>(nodes); A.parent = ar.take<int>(nodes); A.count = ar.take<int>(nodes);
        A.ref = ar.take<int>(n); A.seqv = ar.take<int>(n);
        A.clusterA = ar.take<int>(n); A.clusterB = ar.take<int>(n); A.nn = ar.take<int>(n); A.merged = ar.take<int>(n); A.valid = ar.take<int>(n); A.pos = ar.take<int>(n);
        A.counters = ar.take<int>(8);
        A.olo0 = ar.take<float4>(n); A.ohi0 = ar.take<float4>(n); A.olo1 = ar.take<float4>(n); A.ohi1 = ar.take<float4>(n);
        A.ochild = ar.take<int2>(n);
        A.outIndex = ar.take<int>(nodes); A.leafCode = ar.take<int>(nodes);
        A.slotTri = ar.take<int>(n); A.slotSeq = ar.take<int>(n);
        if (ar.used > ar.size) { err = "internal: build arena too small"; goto fail; }
        lap("allocated");

        for (size_t mesh = 0; mesh < order.size(); ++mesh) {
            const int nm = (int)order[mesh].size();
            if (nm == 0) continue;
            A.n = nm;
            const int grid = (nm + kThreads - 1) / kThreads;
            BCK(cudaMemcpyAsync(d_order, order[mesh].data(), sizeof(int) * (size_t)nm, cudaMemcpyHostToDevice, s));
            BCK(cudaEventRecord(ev0, s));
            {
                const unsigned init[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
                BCK(cudaMemcpyAsync(A.cbound, init, sizeof(init), cudaMemcpyHostToDevice, s));
            }
            primKernel<<<grid, kThreads, 0, s>>>(A);
            unsigned cb[8];
            BCK(cudaMemcpyAsync(cb, A.cbound, sizeof(cb), cudaMemcpyDeviceToHost, s));
            BCK(cudaStreamSynchronize(s));
            float3 cmin, scale;
            {
                const float lo[3] = {fromOrderedHost(cb[0]), fromOrderedHost(cb[1]), fromOrderedHost(cb[2])}, hi[3] = {fromOrderedHost(cb[3]), fromOrderedHost(cb[4]), fromOrderedHost(cb[5])};
                float sc[3];
                for (int k = 0; k < 3; ++k) { const float e = hi[k] - lo[k]; sc[k] = (e > 0.f && std::isfinite(e)) ? 2097151.0f / e : 0.f; }
                cmin = make_float3(lo[0], lo[1], lo[2]); scale = make_float3(sc[0], sc[1], sc[2]);
            }
            mortonKernel<<<grid, kThreads, 0, s>>>(A, cmin, scale);
            BCK(cub::DeviceRadixSort::SortPairs(d_temp, sortBytes, A.keys, A.keys2, A.vals, A.vals2, nm, 0, 63, s));
            leafKernel<<<grid, kThreads, 0, s>>>(A);
            {
                const int init[8] = {nm, nm, 0, 0, 0, 0, 0, 0};
                BCK(cudaMemcpyAsync(A.counters, init, sizeof(init), cudaMemcpyHostToDevice, s));
            }
            int m = nm;
            int* C = A.clusterA;
            int* Cn = A.clusterB;
            int guard = 0;
            while (m > 1) {
                const int g = (m + kThreads - 1) / kThreads;
                nnKernel<<<g, kThreads, 0, s>>>(A, C, m, radius);
                mergeKernel<<<g, kThreads, 0, s>>>(A, C, m);
                BCK(cub::DeviceScan::ExclusiveSum(d_temp, scanBytes, A.valid, A.pos, m, s));
                compactKernel<<<g, kThreads, 0, s>>>(A, Cn, m);
                int m2 = 0;
                BCK(cudaMemcpyAsync(&m2, A.counters + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
                BCK(cudaStreamSynchronize(s));
                if (m2 >= m || m2 < 1 || ++guard > 4096) { err = "internal: clustering made no progress"; goto fail; }
                m = m2;
                std::swap(C, Cn);
            }
            int root = 0;
            BCK(cudaMemcpyAsync(&root, C, sizeof(int), cudaMemcpyDeviceToHost, s));
            int ctr[8];
            BCK(cudaMemcpyAsync(ctr, A.counters, sizeof(ctr), cudaMemcpyDeviceToHost, s));
            BCK(cudaStreamSynchronize(s));
            const int treeNodes = ctr[0];
            if (treeNodes != 2 * nm - 1) { err = "internal: tree has the wrong number of nodes"; goto fail; }
            const int g2 = (treeNodes + kThreads - 1) / kThreads;
            emitKernel<<<g2, kThreads, 0, s>>>(A, treeNodes, root);
            nodeKernel<<<g2, kThreads, 0, s>>>(A, treeNodes);
            BCK(cudaEventRecord(ev1, s));
            lap("tree built");
            BCK(cudaMemcpyAsync(ctr, A.counters, sizeof(ctr), cudaMemcpyDeviceToHost, s));
            int rootOut = -1, rootLeaf = 0;
            BCK(cudaMemcpyAsync(&rootOut, A.outIndex + root, sizeof(int), cudaMemcpyDeviceToHost, s));
            BCK(cudaMemcpyAsync(&rootLeaf, A.leafCode + root, sizeof(int), cudaMemcpyDeviceToHost, s));
            BCK(cudaStreamSynchronize(s));
            BCK(cudaGetLastError());
            const int nOut = ctr[2], nSlots = ctr[3], depth = ctr[4];
            if (nSlots != nm) { err = "internal: leaves do not cover the mesh"; goto fail; }
            if (depth + 2 > max_depth_allowed) { err = "device-built tree is deeper than the traversal stack (depth " + std::to_string(depth) + ", window " + std::to_string(radius) + ")"; goto fail; }
            float ms = 0;
            BCK(cudaEventElapsedTime(&ms, ev0, ev1));
            buildMs += ms;
            outDepth = std::max(outDepth, depth);
            // ---- back to the host form (lower.h BvhNode): the scene upload and the other devices share it
            std::vector<float4> lo0((size_t)nOut), hi0((size_t)nOut), lo1((size_t)nOut), hi1((size_t)nOut);
            std::vector<int2> ch((size_t)nOut);
            std::vector<int> st((size_t)nm), sq((size_t)nm);
            if (nOut > 0) {
                BCK(cudaMemcpyAsync(lo0.data(), A.olo0, sizeof(float4) * (size_t)nOut, cudaMemcpyDeviceToHost, s));
                BCK(cudaMemcpyAsync(hi0.data(), A.ohi0, sizeof(float4) * (size_t)nOut, cudaMemcpyDeviceToHost, s));
                BCK(cudaMemcpyAsync(lo1.data(), A.olo1, sizeof(float4) * (size_t)nOut, cudaMemcpyDeviceToHost, s));
                BCK(cudaMemcpyAsync(hi1.data(), A.ohi1, sizeof(float4) * (size_t)nOut, cudaMemcpyDeviceToHost, s));
                BCK(cudaMemcpyAsync(ch.data(), A.ochild, sizeof(int2) * (size_t)nOut, cudaMemcpyDeviceToHost, s));
            }
            BCK(cudaMemcpyAsync(st.data(), A.slotTri, sizeof(int) * (size_t)nm, cudaMemcpyDeviceToHost, s));
            BCK(cudaMemcpyAsync(sq.data(), A.slotSeq, sizeof(int) * (size_t)nm, cudaMemcpyDeviceToHost, s));
            BCK(cudaStreamSynchronize(s));
            const int nodeBase = (int)(L.bvh_nodes.size() + outNodes.size()), slotBase = (int)(L.bvh_tri.size() + outTri.size());
            auto fixLink = [&](int link) {
                if (link >= 0) return link + nodeBase;
                const int code = ~link;
                return ~((((code >> 3) + slotBase) << 3) | (code & 7));
            };
            for (int o = 0; o < nOut; ++o) {
                BvhNode nd;
                const float4* lo[2] = {&lo0[(size_t)o], &lo1[(size_t)o]};
                const float4* hi[2] = {&hi0[(size_t)o], &hi1[(size_t)o]};
                for (int c = 0; c < 2; ++c) {
                    nd.lo[c][0] = lo[c]->x; nd.lo[c][1] = lo[c]->y; nd.lo[c][2] = lo[c]->z;
                    nd.hi[c][0] = hi[c]->x; nd.hi[c][1] = hi[c]->y; nd.hi[c][2] = hi[c]->z;
                    for (int k = 0; k < 3; ++k) { nd.dlo[c][k] = nd.lo[c][k]; nd.dhi[c][k] = nd.hi[c][k]; }  // conservative boxes serve the FP64 build too
                }
                nd.child[0] = fixLink(ch[(size_t)o].x); nd.child[1] = fixLink(ch[(size_t)o].y);
                outNodes.push_back(nd);
            }
            outTri.insert(outTri.end(), st.begin(), st.end());
            outSeq.insert(outSeq.end(), sq.begin(), sq.end());
            outRoot[mesh] = fixLink(rootOut >= 0 ? rootOut : rootLeaf);
            lap("read back");
        }
        stats.build_ms = buildMs;
    }
    cudaFreeAsync(d_tri, s); cudaFreeAsync(d_order, s); cudaFreeAsync(d_arena, s); cudaFreeAsync(d_temp, s);
    cudaStreamSynchronize(s);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaStreamDestroy(s);
    lap("freed");
    L.bvh_nodes.insert(L.bvh_nodes.end(), outNodes.begin(), outNodes.end());
    L.bvh_tri.insert(L.bvh_tri.end(), outTri.begin(), outTri.end());
    L.bvh_seq.insert(L.bvh_seq.end(), outSeq.begin(), outSeq.end());
    L.mesh_root = outRoot;
    L.max_bvh_depth = std::max(L.max_bvh_depth, outDepth);
    stats.max_depth = outDepth;
    stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return true;
fail:
    if (s) {
        if (d_tri) cudaFreeAsync(d_tri, s);
        if (d_order) cudaFreeAsync(d_order, s);
        if (d_arena) cudaFreeAsync(d_arena, s);
        if (d_temp) cudaFreeAsync(d_temp, s);
        cudaStreamSynchronize(s);
    }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (s) cudaStreamDestroy(s);
    (void)cudaGetLastError();
    return false;
}

}  // namespace ftb
