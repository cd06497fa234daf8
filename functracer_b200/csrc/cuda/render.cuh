// render.cuh — the per-pixel render loop of FuncTracer (Shading.shade and everything below it,
// FuncTracer/Shading.fs:131-147) as ONE persistent CUDA kernel for sm_100a.
//
// Execution model (DESIGN.md "kernel"):
//   * one CTA set per SM, every warp autonomous; no block-level barrier anywhere in the loop;
//   * work = 16x16-pixel tiles handed out by a per-GPU atomic tile counter; inside a warp the
//     pixels of the current tile are dealt to lanes with ballot + popc (a warp-level scan), so a
//     lane whose path ended is immediately re-armed with the next sample / pixel instead of
//     idling ("compaction by regeneration": the 32 lanes stay full until the frame runs out);
//   * F#'s recursion (getColourForRay, Shading.fs:131-139) is an iterative, depth-bounded loop
//     per lane: the reflection ray and its running weight live in registers, so the bounce
//     "queue" costs zero bytes of HBM traffic; shadow rays are traced inline;
//   * a lane owns a pixel for all of its samples, so the per-pixel blend happens in sample
//     order exactly like Array.average (Image.fs:112-116) with no atomics.
//   * the whole scene is brute-forced per ray in the reference's enumeration order (there is no
//     object-level index in the reference, Ray.fs:34): all lanes of a warp read the same leaf
//     record at the same time (one broadcast transaction), and tie-breaks fall out of the order.
//
// Templated on R: float = product, double = FP64 verification build (compiled --fmad=false).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../../include/ftb_rng.h"
#include "../../../include/functracer_b200.h"
#include "device_scene.h"
#include "lower.h"

namespace ftb {

#define FTB_DEV __device__ __forceinline__

// ---- scalar helpers -------------------------------------------------------------------------------
FTB_DEV float sqrt_(float x) { return sqrtf(x); }
FTB_DEV double sqrt_(double x) { return sqrt(x); }
FTB_DEV float abs_(float x) { return fabsf(x); }
FTB_DEV double abs_(double x) { return fabs(x); }
FTB_DEV float floor_(float x) { return floorf(x); }
FTB_DEV double floor_(double x) { return floor(x); }
FTB_DEV float pow_(float x, float y) { return powf(x, y); }
FTB_DEV double pow_(double x, double y) { return pow(x, y); }
FTB_DEV float acos_(float x) { return acosf(x); }
FTB_DEV double acos_(double x) { return acos(x); }
FTB_DEV float asin_(float x) { return asinf(x); }
FTB_DEV double asin_(double x) { return asin(x); }
FTB_DEV float atan2_(float y, float x) { return atan2f(y, x); }
FTB_DEV double atan2_(double y, double x) { return atan2(y, x); }
FTB_DEV float sin_(float x) { return sinf(x); }
FTB_DEV double sin_(double x) { return sin(x); }
FTB_DEV float cos_(float x) { return cosf(x); }
FTB_DEV double cos_(double x) { return cos(x); }
FTB_DEV float tan_(float x) { return tanf(x); }
FTB_DEV double tan_(double x) { return tan(x); }
template <typename R>
FTB_DEV R inf_();
template <>
FTB_DEV float inf_<float>() { return CUDART_INF_F; }
template <>
FTB_DEV double inf_<double>() { return CUDART_INF; }
template <typename R>
FTB_DEV R nan_();
template <>
FTB_DEV float nan_<float>() { return CUDART_NAN_F; }
template <>
FTB_DEV double nan_<double>() { return CUDART_NAN; }
template <typename R>
FTB_DEV R realmax_();  // System.Double.MaxValue of Shading.fs:25,36 in the working precision
template <>
FTB_DEV float realmax_<float>() { return 3.402823466e+38f; }
template <>
FTB_DEV double realmax_<double>() { return 1.7976931348623157e+308; }
// F# max/min on floats propagate NaN
template <typename R>
FTB_DEV R fsmax(R a, R b) { return (a != a || b != b) ? nan_<R>() : (a < b ? b : a); }
template <typename R>
FTB_DEV R fsmin(R a, R b) { return (a != a || b != b) ? nan_<R>() : (a < b ? a : b); }

// ---- CommonTypes.fs ---------------------------------------------------------------------------------
template <typename R>
struct Vec {
    R x, y, z;
};
template <typename R>
FTB_DEV Vec<R> mk(R x, R y, R z) { Vec<R> v; v.x = x; v.y = y; v.z = z; return v; }
template <typename R>
FTB_DEV Vec<R> operator+(Vec<R> a, Vec<R> b) { return mk<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename R>
FTB_DEV Vec<R> operator-(Vec<R> a, Vec<R> b) { return mk<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename R>
FTB_DEV Vec<R> operator-(Vec<R> a) { return mk<R>(-a.x, -a.y, -a.z); }
template <typename R>
FTB_DEV Vec<R> operator*(R s, Vec<R> v) { return mk<R>(s * v.x, s * v.y, s * v.z); }
template <typename R>
FTB_DEV R dot(Vec<R> a, Vec<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename R>
FTB_DEV Vec<R> cross(Vec<R> a, Vec<R> b) { return mk<R>(a.y * b.z - a.z * b.y, b.x * a.z - b.z * a.x, a.x * b.y - a.y * b.x); }  // CommonTypes.fs:17-18
template <typename R>
FTB_DEV R length(Vec<R> v) { return sqrt_(dot(v, v)); }
template <typename R>
FTB_DEV Vec<R> normalise(Vec<R> v)  // CommonTypes.fs:63-67
{
    R l = length(v);
    if (l < R(0.0000001)) return v;
    return (R(1) / l) * v;
}
template <typename R>
FTB_DEV Vec<R> reflect(Vec<R> n, Vec<R> v) { return v - (R(2) * dot(v, n)) * n; }  // :72
template <typename R>
FTB_DEV R angleBetween(Vec<R> a, Vec<R> b) { return acos_(dot(normalise(a), normalise(b))); }  // :74-75
template <typename R>
FTB_DEV Vec<R> perpendicularComponent(Vec<R> a, Vec<R> b) { Vec<R> na = normalise(a); return b - dot(b, na) * na; }  // :77-79

template <typename R>
struct Ray {
    Vec<R> o, d;
};

template <typename R>
FTB_DEV typename V4<R>::type ldg4(const typename V4<R>::type* p) { return __ldg(p); }
template <>
FTB_DEV double4 ldg4<double>(const double4* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// world -> model with the leaf's composed matrix (Transform.fs:85, matrices pre-multiplied on the host)
template <typename R>
FTB_DEV Ray<R> toModel(const DevScene<R>& S, int leaf, bool identity, const Ray<R>& r)
{
    if (identity) return r;
    typedef typename V4<R>::type R4;
    R4 r0 = ldg4<R>(S.leaf_w2m + 3 * leaf), r1 = ldg4<R>(S.leaf_w2m + 3 * leaf + 1), r2 = ldg4<R>(S.leaf_w2m + 3 * leaf + 2);
    Ray<R> m;
    m.o = mk<R>(r0.x * r.o.x + r0.y * r.o.y + r0.z * r.o.z + r0.w, r1.x * r.o.x + r1.y * r.o.y + r1.z * r.o.z + r1.w,
                r2.x * r.o.x + r2.y * r.o.y + r2.z * r.o.z + r2.w);
    m.d = mk<R>(r0.x * r.d.x + r0.y * r.d.y + r0.z * r.d.z, r1.x * r.d.x + r1.y * r.d.y + r1.z * r.d.z,
                r2.x * r.d.x + r2.y * r.d.y + r2.z * r.d.z);
    return m;
}

// ---- per-lane work counters (stats kernels only) -----------------------------------------------------
template <bool STATS>
struct Counters {
    FTB_DEV void add(int, unsigned = 1) {}
    FTB_DEV unsigned get(int) const { return 0; }
};
template <>
struct Counters<true> {
    unsigned int c[ST_COUNT];
    FTB_DEV Counters()
    {
#pragma unroll
        for (int i = 0; i < ST_COUNT; ++i) c[i] = 0;
    }
    FTB_DEV void add(int slot, unsigned n = 1) { c[slot] += n; }
    FTB_DEV unsigned get(int slot) const { return c[slot]; }
};

// ---- Math.quadratic (Math.fs:4-10): far ("+") root first ----------------------------------------------
template <typename R>
FTB_DEV bool quadratic(R a, R b, R c, R& t0, R& t1)
{
    R discriminant = b * b - R(4) * a * c;
    if (discriminant < R(0)) return false;
    R sq = sqrt_(discriminant);
    R twoa = R(2) * a;
    t0 = (-b + sq) / twoa;
    t1 = (-b - sq) / twoa;
    return true;
}

// Plane.intersect for Plane(Zero, unitY) in the leaf's frame (Plane.fs:9-20).  Returns false if no hit.
template <typename R>
FTB_DEV bool planeT(const Ray<R>& r, R& t, Vec<R>& p)
{
    const R eps = R(0.0000001);
    R num = -r.o.y;
    R denom = r.d.y;
    if (abs_(denom) < eps) {
        if (num < eps) { t = R(0); p = r.o; return true; }
        return false;
    }
    t = num / denom;
    p = mk<R>(r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z);
    return true;
}

// BoundingBox.intersects (BoundingBox.fs:32-58)
template <typename R>
FTB_DEV bool aabbIntersects(const R* __restrict__ bb, const Ray<R>& ray, Vec<R> inv)
{
    R bx0 = __ldg(bb + 0), by0 = __ldg(bb + 1), bz0 = __ldg(bb + 2), bx1 = __ldg(bb + 3), by1 = __ldg(bb + 4), bz1 = __ldg(bb + 5);
    bool sx = inv.x < R(0), sy = inv.y < R(0), sz = inv.z < R(0);
    R tmin = ((sx ? bx1 : bx0) - ray.o.x) * inv.x;
    R tmax = ((sx ? bx0 : bx1) - ray.o.x) * inv.x;
    R tymin = ((sy ? by1 : by0) - ray.o.y) * inv.y;
    R tymax = ((sy ? by0 : by1) - ray.o.y) * inv.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    tmin = fsmax(tymin, tmin);
    tmax = fsmin(tymax, tmax);
    R tzmin = ((sz ? bz1 : bz0) - ray.o.z) * inv.z;
    R tzmax = ((sz ? bz0 : bz1) - ray.o.z) * inv.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    tmin = fsmax(tzmin, tmin);
    tmax = fsmin(tzmax, tmax);
    return (tmin < inf_<R>()) && (tmax > -inf_<R>());
}

// Triangle.fs:43-66 with e1, e2 precomputed; returns t or false.
template <typename R>
FTB_DEV bool triangleT(const DevScene<R>& S, int tri, const Ray<R>& ray, R& t)
{
    typedef typename V4<R>::type R4;
    const R epsilon = R(0.0000001);
    R4 a0 = ldg4<R>(S.tris + 3 * tri), a1 = ldg4<R>(S.tris + 3 * tri + 1), a2 = ldg4<R>(S.tris + 3 * tri + 2);
    Vec<R> v0 = mk<R>(a0.x, a0.y, a0.z), edge1 = mk<R>(a1.x, a1.y, a1.z), edge2 = mk<R>(a2.x, a2.y, a2.z);
    Vec<R> h = cross(ray.d, edge2);
    R a = dot(edge1, h);
    if (a > -epsilon && a < epsilon) return false;
    R f = R(1) / a;
    Vec<R> s = ray.o - v0;
    R u = f * dot(s, h);
    if (u < R(0) || u > R(1)) return false;
    Vec<R> q = cross(s, edge1);
    R v = dot(f * ray.d, q);
    if (v < R(0) || u + v > R(1)) return false;
    t = dot(f * edge2, q);
    return t > epsilon;
}

// ---- leaf intersection: calls sink.hit(t, sub) for every crossing, in the reference's order ---------
// sub: cube face 0..5 (Cube.fs:24), triangle index for meshes, else the leaf's payload.
template <typename R, bool STATS, class Sink>
FTB_DEV void intersectLeaf(const DevScene<R>& S, int leaf, const Ray<R>& wr, Sink& sink, Counters<STATS>& cn)
{
    const int4 meta = __ldg(S.leaf_meta + leaf);
    const int kind = meta.x & 0xff;
    const bool identity = (meta.x >> 8) & 1;
    const Ray<R> r = toModel(S, leaf, identity, wr);
    cn.add(ST_LEAF0 + kind);
    if (!identity) cn.add(ST_XFORM);
    switch (kind) {
    case LEAF_SPHERE: {  // Sphere.fs:11-21
        R a = dot(r.d, r.d), b = R(2) * dot(r.o, r.d), c = dot(r.o, r.o) - R(1), t0, t1;
        if (quadratic(a, b, c, t0, t1)) { sink.hit(t0, meta.w); sink.hit(t1, meta.w); }
        break;
    }
    case LEAF_PLANE: {  // Plane.fs:32-33
        R t; Vec<R> p;
        if (planeT(r, t, p)) sink.hit(t, meta.w);
        break;
    }
    case LEAF_SQUARE: {  // Cube.fs:9-15
        R t; Vec<R> p;
        if (planeT(r, t, p) && (p.x >= R(0)) && (p.x <= R(1)) && (p.z >= R(0)) && (p.z <= R(1))) sink.hit(t, meta.w);
        break;
    }
    case LEAF_CIRCLE: {  // Cylinder.fs:22
        R t; Vec<R> p;
        if (planeT(r, t, p) && length(p) < R(1)) sink.hit(t, meta.w);
        break;
    }
    case LEAF_CYLINDER: {  // Cylinder.fs:8-20
        R a = r.d.x * r.d.x + r.d.z * r.d.z, b = R(2) * (r.o.x * r.d.x + r.o.z * r.d.z), c = r.o.x * r.o.x + r.o.z * r.o.z - R(1), t0, t1;
        if (quadratic(a, b, c, t0, t1)) {
            R py = r.o.y + t0 * r.d.y;
            if (py >= R(0) && py <= R(1)) sink.hit(t0, meta.w);
            py = r.o.y + t1 * r.d.y;
            if (py >= R(0) && py <= R(1)) sink.hit(t1, meta.w);
        }
        break;
    }
    case LEAF_CONE: {  // Cone.fs:7-28
        R oy = r.o.y - R(1);
        R a = r.d.x * r.d.x + r.d.z * r.d.z - r.d.y * r.d.y, b = R(2) * (r.o.x * r.d.x + r.o.z * r.d.z - oy * r.d.y),
          c = r.o.x * r.o.x + r.o.z * r.o.z - oy * oy, t0, t1;
        if (quadratic(a, b, c, t0, t1)) {
            R py = (oy + t0 * r.d.y) + R(1);
            if (py >= R(0) && py <= R(1)) sink.hit(t0, meta.w);
            py = (oy + t1 * r.d.y) + R(1);
            if (py >= R(0) && py <= R(1)) sink.hit(t1, meta.w);
        }
        break;
    }
    case LEAF_CUBE: {  // Cube.fs:17-25, the six squares in the cube's centred frame shifted by +.5
        const R eps = R(0.0000001);
        const R ox = r.o.x + R(0.5), oy = r.o.y + R(0.5), oz = r.o.z + R(0.5);
        // face f: plane coordinate w (origin wo, direction wd), offset k in {0,1}, in-face coords (a, b)
        // bottom/top: w = y, (a,b) = (x,z); left/right: w = x, (a,b) = (y,z); front/back: w = z, (a,b) = (x,y).
        // num/denom signs follow each square's own frame (DESIGN.md "cube"): bottom/top: num = k - w, denom = +wd;
        // left/right/front/back (rotated frames): num = w - k, denom = -wd.
#define FTB_CUBE_FACE(face, wo, wd, ao, ad, bo, bd, k, rotated)                                   \
        {                                                                                         \
            R num = (rotated) ? ((wo) - R(k)) : (R(k) - (wo));                                    \
            R denom = (rotated) ? -(wd) : (wd);                                                   \
            R t; bool ok = true;                                                                  \
            if (abs_(denom) < eps) { if (num < eps) t = R(0); else ok = false; }                  \
            else t = num / denom;                                                                 \
            if (ok) {                                                                             \
                R pa = (ao) + t * (ad), pb = (bo) + t * (bd);                                     \
                if ((pa >= R(0)) && (pa <= R(1)) && (pb >= R(0)) && (pb <= R(1))) sink.hit(t, face); \
            }                                                                                     \
        }
        FTB_CUBE_FACE(0, oy, r.d.y, ox, r.d.x, oz, r.d.z, 0, false)
        FTB_CUBE_FACE(1, oy, r.d.y, ox, r.d.x, oz, r.d.z, 1, false)
        FTB_CUBE_FACE(2, ox, r.d.x, oy, r.d.y, oz, r.d.z, 0, true)
        FTB_CUBE_FACE(3, ox, r.d.x, oy, r.d.y, oz, r.d.z, 1, true)
        FTB_CUBE_FACE(4, oz, r.d.z, ox, r.d.x, oy, r.d.y, 0, true)
        FTB_CUBE_FACE(5, oz, r.d.z, ox, r.d.x, oy, r.d.y, 1, true)
#undef FTB_CUBE_FACE
        break;
    }
    case LEAF_TRIANGLE: {
        R t;
        if (triangleT(S, meta.w, r, t)) sink.hit(t, 0);
        break;
    }
    case LEAF_MESH: {  // BspMesh.intersect (BspMesh.fs:67-76): AABB gate, right subtree, then left
        int stack[kBspStack];
        int sp = 0;
        stack[sp++] = __ldg(S.mesh_root + meta.w);
        const Vec<R> inv = mk<R>(R(1) / r.d.x, R(1) / r.d.y, R(1) / r.d.z);
        while (sp > 0) {
            int link = stack[--sp];
            if (link < 0) {
                const int2 lf = __ldg(S.bsp_leaves + (~link));
                for (int i = 0; i < lf.y; ++i) {
                    R t;
                    cn.add(ST_TRI_TESTS_IN_MESH);
                    if (triangleT(S, lf.x + i, r, t)) sink.hit(t, lf.x + i);
                    if (sink.done()) { sp = 0; break; }
                }
            } else {
                cn.add(ST_BSP_NODES);
                if (aabbIntersects(S.bsp_aabb + 6 * link, r, inv)) {
                    const int2 ln = __ldg(S.bsp_links + link);
                    if (sp + 2 <= kBspStack) { stack[sp++] = ln.x; stack[sp++] = ln.y; }  // right pops first
                }
            }
        }
        break;
    }
    }
}

// ---- sinks ----------------------------------------------------------------------------------------------
// Scene.closest (Scene.fs:112-116): smallest t >= 0, first in enumeration order on ties.
template <typename R>
struct NearestSink {
    R t;
    int leaf, sub, flip;
    int cur;  // leaf being intersected
    FTB_DEV void hit(R ht, int hsub)
    {
        if (ht >= R(0) && ht < t) { t = ht; leaf = cur; sub = hsub; flip = 0; }
    }
    FTB_DEV bool done() const { return false; }
};
// Scene.lightIsBocked (Scene.fs:119-121) for a leaf whose surface has applyLighting = true.
template <typename R>
struct AnySink {
    R maxDistance;
    bool blocked;
    FTB_DEV void hit(R ht, int) { if (ht >= R(0) && ht < maxDistance) blocked = true; }
    FTB_DEV bool done() const { return blocked; }
};
// CSG operand: append to the per-ray hit stack.
template <typename R>
struct HitRec {
    R t;
    unsigned int id;  // leaf (0..21) | sub (22..24) | flip (30) | side B (31)
};
constexpr unsigned kIdFlip = 1u << 30, kIdSideB = 1u << 31, kIdSubShift = 22, kIdLeafMask = (1u << 22) - 1;
template <typename R>
struct ListSink {
    HitRec<R>* stack;
    int top;
    int cur;
    bool overflow;
    FTB_DEV void hit(R ht, int hsub)
    {
        if (top < kHitCap) { stack[top].t = ht; stack[top].id = (unsigned)cur | ((unsigned)(hsub & 7) << kIdSubShift); ++top; }
        else overflow = true;
    }
    FTB_DEV bool done() const { return false; }
};

// Csg.fs:19-72 as a lookup: key = hitB * 4 + inA * 2 + inB -> 0 Take, 1 Discard, 2 Flip (2 bits each).
FTB_DEV unsigned csgRuleTable(int op)
{
    // keys: A-hit (inA,inB): 0 (F,F) OutsideIntoA, 1 (F,T) BIntoAB, 2 (T,F) AIntoOutside, 3 (T,T) ABleaveA
    //       B-hit:           4 (F,F) OutsideIntoB, 5 (F,T) BIntoOutside, 6 (T,F) AIntoAB, 7 (T,T) ABleaveB
#define FTB_RULES(k0, k1, k2, k3, k4, k5, k6, k7) ((k0) | ((k1) << 2) | ((k2) << 4) | ((k3) << 6) | ((k4) << 8) | ((k5) << 10) | ((k6) << 12) | ((k7) << 14))
    switch (op) {
    case OP_UNION: return FTB_RULES(0u, 1u, 0u, 1u, 0u, 0u, 1u, 1u);      // Csg.fs:19-25
    case OP_SUBTRACT: return FTB_RULES(0u, 1u, 0u, 1u, 1u, 1u, 2u, 2u);   // :27-33
    case OP_INTERSECT: return FTB_RULES(1u, 0u, 1u, 0u, 1u, 1u, 0u, 0u);  // :35-44
    default: return FTB_RULES(0u, 2u, 0u, 2u, 0u, 0u, 2u, 2u);            // exclude :46-55
    }
#undef FTB_RULES
}

// Evaluates one CSG program (Csg.constructedSolid, Csg.fs:74-94) on the per-ray hit stack.
// On return stack[0..n) holds the root's crossings sorted by t.
template <typename R, bool STATS>
FTB_DEV int evalCsg(const DevScene<R>& S, int opFirst, int opCount, const Ray<R>& wr, HitRec<R>* stack, bool& overflow, Counters<STATS>& cn)
{
    int counts[kMaxLists];
    int nlists = 0;
    ListSink<R> sink;
    sink.stack = stack; sink.top = 0; sink.overflow = false;
    for (int i = 0; i < opCount; ++i) {
        const int2 op = __ldg(S.ops + opFirst + i);
        if (op.x == OP_LEAF) {
            int start = sink.top;
            sink.cur = op.y;
            intersectLeaf<R, STATS>(S, op.y, wr, sink, cn);
            if (nlists < kMaxLists) counts[nlists++] = sink.top - start; else sink.overflow = true;
        } else if (op.x == OP_GROUP) {
            int c = 0;
            for (int k = 0; k < op.y; ++k) c += counts[nlists - 1 - k];
            nlists -= op.y - 1;
            counts[nlists - 1] = c;
        } else if (op.x == OP_EMPTY) {
            if (nlists < kMaxLists) counts[nlists++] = 0; else sink.overflow = true;
        } else {
            cn.add(ST_CSG_OPS);
            const int nb = counts[nlists - 1], na = counts[nlists - 2];
            const int start = sink.top - na - nb, end = sink.top;
            for (int k = start + na; k < end; ++k) stack[k].id |= kIdSideB;
            // Seq.sortBy (stable): insertion sort, strict '<' so equal keys keep A-before-B order
            for (int k = start + 1; k < end; ++k) {
                HitRec<R> x = stack[k];
                int j = k;
                while (j > start && x.t < stack[j - 1].t) { stack[j] = stack[j - 1]; --j; }
                stack[j] = x;
            }
            const unsigned rules = csgRuleTable(op.x);
            bool inA = false, inB = false;
            int out = start;
            for (int k = start; k < end; ++k) {
                HitRec<R> x = stack[k];
                const bool hitB = (x.id & kIdSideB) != 0;
                const unsigned rule = (rules >> (2 * ((hitB ? 4 : 0) + (inA ? 2 : 0) + (inB ? 1 : 0)))) & 3u;
                if (hitB) inB = !inB; else inA = !inA;
                x.id &= ~kIdSideB;
                if (rule == 2u) x.id ^= kIdFlip;
                if (rule != 1u) stack[out++] = x;
            }
            sink.top = out;
            --nlists;
            counts[nlists - 1] = out - start;
        }
    }
    overflow = overflow || sink.overflow;
    return sink.top;
}

template <typename R>
struct HitInfo {
    R t;
    int leaf, sub, flip;
};

// closest over the whole scene: Scene.intersectScene (Scene.fs:118)
template <typename R, bool STATS>
FTB_DEV HitInfo<R> traceNearest(const DevScene<R>& S, const Ray<R>& wr, bool& overflow, Counters<STATS>& cn)
{
    NearestSink<R> best;
    best.t = inf_<R>(); best.leaf = -1; best.sub = 0; best.flip = 0;
    for (int it = 0; it < S.n_items; ++it) {
        const int4 item = __ldg(S.items + it);
        if (item.x == ITEM_LEAF) {
            best.cur = item.y;
            intersectLeaf<R, STATS>(S, item.y, wr, best, cn);
        } else {
            HitRec<R> stack[kHitCap];
            const int n = evalCsg<R, STATS>(S, item.y, item.z, wr, stack, overflow, cn);
            for (int k = 0; k < n; ++k) {  // sorted by t: the first t >= 0 that beats best wins
                const R ht = stack[k].t;
                if (ht >= R(0)) {
                    if (ht < best.t) {
                        best.t = ht; best.leaf = (int)(stack[k].id & kIdLeafMask); best.sub = (int)((stack[k].id >> kIdSubShift) & 7u);
                        best.flip = (stack[k].id & kIdFlip) ? 1 : 0;
                    }
                    break;
                }
            }
        }
    }
    HitInfo<R> h;
    h.t = best.t; h.leaf = best.leaf; h.sub = best.sub; h.flip = best.flip;
    return h;
}

// lightIsBocked over the whole scene (Scene.fs:119-121)
template <typename R, bool STATS>
FTB_DEV bool traceAny(const DevScene<R>& S, const Ray<R>& wr, R maxDistance, bool& overflow, Counters<STATS>& cn)
{
    cn.add(ST_SHADOW);
    AnySink<R> any;
    any.maxDistance = maxDistance; any.blocked = false;
    for (int it = 0; it < S.n_items && !any.blocked; ++it) {
        const int4 item = __ldg(S.items + it);
        if (!item.w) continue;  // nothing under it has applyLighting
        if (item.x == ITEM_LEAF) {
            intersectLeaf<R, STATS>(S, item.y, wr, any, cn);
        } else {
            HitRec<R> stack[kHitCap];
            const int n = evalCsg<R, STATS>(S, item.y, item.z, wr, stack, overflow, cn);
            for (int k = 0; k < n; ++k) {
                const R ht = stack[k].t;
                if (ht >= R(0) && ht < maxDistance) {
                    const int leaf = (int)(stack[k].id & kIdLeafMask);
                    const int surf = __ldg(S.leaf_meta + leaf).y;
                    if (__ldg(S.surf_i + surf).z) { any.blocked = true; break; }
                }
            }
        }
    }
    return any.blocked;
}

// ---- textures (Textures/Texture.fs, Textures/Image.fs:27-36) ---------------------------------------------
template <typename R>
FTB_DEV R repeatOne(R x)
{
    R a = abs_(x - floor_(x));
    return (a < R(0)) ? R(1) - a : a;
}
template <typename R>
FTB_DEV Vec<R> evalTexture(const DevScene<R>& S, int tex, R u, R v)
{
    typedef typename V4<R>::type R4;
    const int4 ti = __ldg(S.tex_i + tex);
    for (int k = 0; k < ti.y; ++k) {
        const int kind = __ldg(S.texop_kind + ti.x + k);
        const R a = __ldg(S.texop_ab + 2 * (ti.x + k)), b = __ldg(S.texop_ab + 2 * (ti.x + k) + 1);
        if (kind == FTB_TEX_SCALE) { u = u / a; v = v / b; }  // Texture.fs:14-16
        else {                                                  // Texture.fs:18-22, a = cos, b = sin
            R x = a * u + R(0) * R(0) + b * v;
            R z = (-b) * u + R(0) * R(0) + a * v;
            u = x; v = z;
        }
    }
    const R ru = repeatOne(u), rv = repeatOne(v);
    if (ti.z == FTB_TEX_GRID) {  // Texture.fs:24-29
        const R4 c1 = ldg4<R>(S.tex_c1 + tex), c2 = ldg4<R>(S.tex_c2 + tex);
        const bool first = (ru < R(0.5) && rv < R(0.5)) || (!(ru < R(0.5)) && (ru > R(0.5) && rv > R(0.5)));
        return first ? mk<R>(c1.x, c1.y, c1.z) : mk<R>(c2.x, c2.y, c2.z);
    }
    const int4 im = __ldg(S.img_i + ti.w);
    int x = (int)floor_(ru * (R)im.y), y = (int)floor_(rv * (R)im.z);
    x = min(max(x, 0), im.y - 1);  // repeat can return exactly 1.0; clamped (documented)
    y = min(max(y, 0), im.z - 1);
    const uchar4 px = __ldg(S.texels + im.x + y * im.y + x);
    return mk<R>((R)px.x / R(255), (R)px.y / R(255), (R)px.z / R(255));
}

// ---- Jitter.fs on the ftb_rng contract --------------------------------------------------------------------
template <typename R>
FTB_DEV Vec<R> jitterVector(unsigned long long seed, unsigned long long sample, unsigned depth, unsigned light, unsigned idx, R tanHalfAngle, Vec<R> vector)
{
    Vec<R> normalised = normalise(vector);
    Vec<R> generator = (normalised.x > R(0.9)) ? mk<R>(R(0), R(1), R(0)) : mk<R>(R(1), R(0), R(0));
    Vec<R> i = normalise(cross(generator, normalised));
    Vec<R> j = cross(i, normalised);
    R x, y;
    for (unsigned attempt = 0;; ++attempt) {  // Jitter.circle (Jitter.fs:15-21)
        // (2 bits - 2^24) / 2^24: exact in float and double, so every build sees the same offsets
        x = (R)(2 * (int)ftb_rng_bits24(seed, sample, depth, light, idx, attempt, 0) - 16777216) * R(5.9604644775390625e-08);
        y = (R)(2 * (int)ftb_rng_bits24(seed, sample, depth, light, idx, attempt, 1) - 16777216) * R(5.9604644775390625e-08);
        if (!((x * x + y * y) > R(1))) break;
    }
    return normalise((normalised + (tanHalfAngle * x) * i) + (tanHalfAngle * y) * j);
}

// ---- winner finalisation: p, n, uv, material of the nearest hit ---------------------------------------------
template <typename R>
struct Fragment {
    Vec<R> p, n;
    Vec<R> colour;
    R roughness, reflectance, shineyness;
    bool applyLighting;
};

template <typename R>
FTB_DEV Fragment<R> finalise(const DevScene<R>& S, const Ray<R>& wr, const HitInfo<R>& h)
{
    typedef typename V4<R>::type R4;
    const int4 meta = __ldg(S.leaf_meta + h.leaf);
    const int kind = meta.x & 0xff;
    const bool identity = (meta.x >> 8) & 1;
    const Ray<R> r = toModel(S, h.leaf, identity, wr);
    const Vec<R> pm = mk<R>(r.o.x + h.t * r.d.x, r.o.y + h.t * r.d.y, r.o.z + h.t * r.d.z);
    Vec<R> nm = mk<R>(R(0), R(1), R(0));
    R u = R(0), v = R(0);
    switch (kind) {
    case LEAF_SPHERE:
        nm = normalise(pm);
        break;
    case LEAF_PLANE: case LEAF_SQUARE: case LEAF_CIRCLE:
        u = pm.x; v = pm.z;
        break;
    case LEAF_CYLINDER: {
        Vec<R> n = normalise(mk<R>(pm.x, R(0), pm.z));
        nm = (dot(n, r.d) < R(0)) ? n : -n;
        break;
    }
    case LEAF_CONE: {
        Vec<R> n = normalise(mk<R>(pm.x, -(pm.y - R(1)), pm.z));
        nm = (dot(n, r.d) < R(0)) ? n : -n;
        break;
    }
    case LEAF_CUBE: {
        const R x1 = pm.x + R(0.5), y1 = pm.y + R(0.5), z1 = pm.z + R(0.5);
        switch (h.sub) {
        case 0: nm = mk<R>(R(0), R(-1), R(0)); u = x1; v = z1; break;
        case 1: nm = mk<R>(R(0), R(1), R(0)); u = x1; v = z1; break;
        case 2: nm = mk<R>(R(-1), R(0), R(0)); u = y1; v = z1; break;
        case 3: nm = mk<R>(R(1), R(0), R(0)); u = y1; v = z1; break;
        case 4: nm = mk<R>(R(0), R(0), R(-1)); u = x1; v = y1; break;
        default: nm = mk<R>(R(0), R(0), R(1)); u = x1; v = y1; break;
        }
        break;
    }
    case LEAF_TRIANGLE: case LEAF_MESH: {
        const int tri = (kind == LEAF_TRIANGLE) ? meta.w : h.sub;
        R4 a1 = ldg4<R>(S.tris + 3 * tri + 1), a2 = ldg4<R>(S.tris + 3 * tri + 2);
        nm = normalise(cross(mk<R>(a1.x, a1.y, a1.z), mk<R>(a2.x, a2.y, a2.z)));
        break;
    }
    }
    Fragment<R> f;
    const int4 si = __ldg(S.surf_i + meta.y);
    if (si.x >= 0 && kind == LEAF_SPHERE) {  // Sphere.setUV (Sphere.fs:6-10), only needed when textured
        u = R(0.5) + (atan2_(nm.z, nm.x) / (R(2) * R(3.14159265358979323846)));
        v = R(0.5) - asin_(nm.y) / R(3.14159265358979323846);
    }
    // n <- normalise(normalToWorld * n), normalToWorld = transpose(worldToModel) (Transform.fs:83,86)
    Vec<R> nw = nm;
    if (!identity) {
        R4 r0 = ldg4<R>(S.leaf_w2m + 3 * h.leaf), r1 = ldg4<R>(S.leaf_w2m + 3 * h.leaf + 1), r2 = ldg4<R>(S.leaf_w2m + 3 * h.leaf + 2);
        nw = normalise(mk<R>(r0.x * nm.x + r1.x * nm.y + r2.x * nm.z, r0.y * nm.x + r1.y * nm.y + r2.y * nm.z, r0.z * nm.x + r1.z * nm.y + r2.z * nm.z));
    }
    if (h.flip) nw = R(-1) * nw;  // Csg Flip (Csg.fs:87)
    f.n = nw;
    // p = modelToWorld * p_model == o + t d in exact arithmetic (t is invariant); the ray itself is used
    f.p = mk<R>(wr.o.x + h.t * wr.d.x, wr.o.y + h.t * wr.d.y, wr.o.z + h.t * wr.d.z);
    const R4 sa = ldg4<R>(S.surf_a + meta.y), sb = ldg4<R>(S.surf_b + meta.y);
    Vec<R> col = mk<R>(sa.x, sa.y, sa.z);
    if (si.x >= 0) {
        col = evalTexture(S, si.x, u, v);
        for (int k = 0; k < si.y; ++k) col = mk<R>(col.z, col.x, col.y);  // Colour.hueShift (CommonTypes.fs:90)
    }
    f.colour = col;
    f.roughness = sa.w; f.reflectance = sb.x; f.shineyness = sb.y;
    f.applyLighting = si.z != 0;
    return f;
}

// ---- Shading.fs ------------------------------------------------------------------------------------------------
template <typename R>
FTB_DEV Vec<R> roughDiffuse(const Fragment<R>& f, Vec<R> lightDir, Vec<R> viewD)  // Shading.fs:50-63
{
    R roughness = f.roughness * f.roughness;
    R rayAngle = angleBetween(f.n, -viewD);
    R lightAngle = angleBetween(f.n, -lightDir);
    R alpha = fsmax(rayAngle, lightAngle);
    R beta = fsmin(rayAngle, lightAngle);
    R A = R(1) - R(0.5) * roughness / (roughness + R(0.33));
    R B = R(0.45) * roughness / (roughness + R(0.09));
    Vec<R> tangentLight = normalise(perpendicularComponent(f.n, -lightDir));
    Vec<R> tangentRay = normalise(perpendicularComponent(f.n, -viewD));
    R intensity = cos_(lightAngle) * (A + (B * fsmax(R(0), dot(tangentLight, tangentRay)) * sin_(alpha) * tan_(beta)));
    return intensity * f.colour;
}

// One level of getColourForRay (Shading.fs:131-139) minus the recursion: the sum over lights of
// (specular + diffuse), or of material.colour when lighting is off.  `sample`/`depth` key the RNG.
template <typename R, bool STATS>
FTB_DEV Vec<R> shadeLocal(const DevScene<R>& S, const Fragment<R>& f, Vec<R> viewD, unsigned long long seed, unsigned long long sample, unsigned depth,
                          bool& overflow, Counters<STATS>& cn)
{
    typedef typename V4<R>::type R4;
    Vec<R> total = mk<R>(R(0), R(0), R(0));
    const Vec<R> shadowRayOrigin = f.p + R(0.0001) * f.n;  // Shading.fs:111
    const Vec<R> viewDirection = normalise(viewD);
    const Vec<R> normal = normalise(f.n);
    for (int li = 0; li < S.n_lights; ++li) {
        const int2 lk = __ldg(S.light_i + li);
        const R4 la = ldg4<R>(S.light_a + li), lc = ldg4<R>(S.light_c + li);
        const Vec<R> lv = mk<R>(la.x, la.y, la.z);
        R intensity;  // shadowLightIntensity (Shading.fs:33-42)
        Vec<R> ldir;  // lightDirection (Shading.fs:44-48)
        Ray<R> sr;
        sr.o = shadowRayOrigin;
        if (lk.x == FTB_LIGHT_DIRECTIONAL) {
            sr.d = -lv;
            intensity = traceAny<R, STATS>(S, sr, realmax_<R>(), overflow, cn) ? R(0) : R(1);
            ldir = lv;
        } else if (lk.x == FTB_LIGHT_SOFT_DIRECTIONAL) {  // softShadowLightIntensity (Shading.fs:24-31)
            int occluded = 0;
            for (int k = 0; k < lk.y; ++k) {
                sr.d = jitterVector<R>(seed, sample, depth, (unsigned)li, (unsigned)k, la.w, -lv);
                if (traceAny<R, STATS>(S, sr, realmax_<R>(), overflow, cn)) ++occluded;
            }
            intensity = (R)(lk.y - occluded) / (R)lk.y;
            ldir = lv;
        } else {
            const R4 lb = ldg4<R>(S.light_b + li);
            const Vec<R> dvec = lv - shadowRayOrigin;
            const R distance = length(dvec);
            sr.d = normalise(dvec);
            if (traceAny<R, STATS>(S, sr, distance, overflow, cn)) intensity = R(0);
            else intensity = R(1) / (lb.x + distance * (lb.y + distance * lb.z));  // Light.attenuate (Light.fs:16-17)
            ldir = normalise(f.p - lv);
        }
        const Vec<R> lightColour = mk<R>(intensity * lc.x, intensity * lc.y, intensity * lc.z);
        if (!f.applyLighting) {  // shadeIfRequired (Shading.fs:100-104)
            total = total + f.colour;
            continue;
        }
        Vec<R> acc = mk<R>(R(0), R(0), R(0));
        {  // specularShader (Shading.fs:78-87)
            const Vec<R> reflectedLightDirection = normalise(reflect(normal, ldir));
            const R si = pow_(dot(viewDirection, -reflectedLightDirection), f.shineyness);
            if (!(f.shineyness <= R(0) || si <= R(0))) acc = acc + mk<R>(lightColour.x * si, lightColour.y * si, lightColour.z * si);
        }
        if (f.roughness == R(0)) {  // lambertianDiffuse (Shading.fs:65-70)
            const R di = dot(-ldir, f.n);
            acc = acc + mk<R>(di * (f.colour.x * lightColour.x), di * (f.colour.y * lightColour.y), di * (f.colour.z * lightColour.z));
        } else {
            acc = acc + roughDiffuse(f, ldir, viewD);
        }
        total = total + acc;
    }
    return total;
}

// ---- Image.fs (sampling half) -----------------------------------------------------------------------------------
template <typename R>
FTB_DEV Ray<R> primaryRay(const DevFrame<R>& F, int px, int py, int s, unsigned long long sampleIndex)
{
    const R jitterX = (F.spp > 0 && F.jitter) ? __ldg(F.jitter + 2 * s) : R(0);
    const R jitterY = (F.spp > 0 && F.jitter) ? __ldg(F.jitter + 2 * s + 1) : R(0);
    // rayThroughPixel (Image.fs:83-89)
    const R centreX = F.tlx + (R)px * F.pw, centreY = F.tly - (R)py * F.ph;
    const R jx = centreX + jitterX * F.pw, jy = centreY + jitterY * F.ph;
    const Vec<R> k = mk<R>(F.cam_k[0], F.cam_k[1], F.cam_k[2]), i = mk<R>(F.cam_i[0], F.cam_i[1], F.cam_i[2]), j = mk<R>(F.cam_j[0], F.cam_j[1], F.cam_j[2]);
    Ray<R> r;
    r.o = mk<R>(F.cam_o[0], F.cam_o[1], F.cam_o[2]);
    r.d = (k + jx * i) + jy * j;
    if (F.has_focus) {  // depthOfFieldJitter (Image.fs:91-94, Ray.fs:15-18)
        r.o = r.o + F.focal * r.d;
        r.d = jitterVector<R>(F.seed, sampleIndex, 0u, FTB_RNG_STREAM_CAMERA, 0u, F.tan_half_aperture, r.d);
        r.o = r.o + (-F.focal) * r.d;
    }
    return r;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
template <typename R, bool STATS>
__global__ void __launch_bounds__(kBlockThreads) render_kernel(const DevScene<R> S, const DevFrame<R> F)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Counters<STATS> cn;
    bool overflow = false;

    // warp-uniform tile cursor
    int tile_slot = -1;  // local tile index = value of the atomic counter
    int tile_x0 = 0, tile_y0 = 0, tile_w = 0, tile_n = 0, tile_pos = 0;
    bool exhausted = false;

    // per-lane pixel / path state
    bool have_pixel = false, active = false;
    int px = 0, py = 0, slot = 0, s = 0, limit = 0;
    unsigned depth = 0;
    long long unit = 0;  // pixel index in the grid (mode 0) or ray index (mode 1)
    Vec<R> pixsum = mk<R>(R(0), R(0), R(0)), scol = mk<R>(R(0), R(0), R(0));
    R weight = R(1);
    Ray<R> ray;
    ray.o = mk<R>(R(0), R(0), R(0)); ray.d = ray.o;
    const int spp = F.mode == 0 ? F.spp : 1;

    for (;;) {
        // ---- re-arm lanes whose path ended -------------------------------------------------------------
        if (!active && have_pixel) {
            pixsum = pixsum + scol;  // Array.average folds from Zero in sample order (Image.fs:115)
            if (s + 1 < spp) {
                ++s;
            } else {
                // DivideByInt (CommonTypes.fs:43)
                R* o = F.out + 3 * (long long)slot;
                o[0] = pixsum.x / (R)spp; o[1] = pixsum.y / (R)spp; o[2] = pixsum.z / (R)spp;
                have_pixel = false;
            }
        }
        bool need = !active && !have_pixel;
        unsigned m = __ballot_sync(full, need);
        while (m && !exhausted) {
            if (tile_pos >= tile_n) {  // warp grabs the next tile from the per-GPU queue
                unsigned c = 0;
                if (lane == 0) c = atomicAdd(F.tile_counter, 1u);
                c = __shfl_sync(full, c, 0);
                if (c >= (unsigned)F.n_local_tiles) { exhausted = true; break; }
                tile_slot = (int)c;
                tile_pos = 0;
                if (F.mode == 0) {
                    const int tile = (int)c * F.shard_count + F.shard_index;
                    tile_x0 = (tile % F.tiles_x) * FTB_TILE_W;
                    tile_y0 = (tile / F.tiles_x) * FTB_TILE_H;
                    tile_w = min(FTB_TILE_W, F.gw - tile_x0);
                    tile_n = tile_w * min(FTB_TILE_H, F.gh - tile_y0);
                } else {
                    const long long first = (long long)c * FTB_TILE_PIXELS;
                    tile_n = (int)min((long long)FTB_TILE_PIXELS, F.n_rays - first);
                }
            }
            // deal the tile's remaining pixels to the lanes that need one (ballot + popc = warp scan)
            const int rank = __popc(m & lt_mask);
            const int avail = tile_n - tile_pos;
            if (need && rank < avail) {
                const int j = tile_pos + rank;
                if (F.mode == 0) {
                    const int lx = j % tile_w, ly = j / tile_w;
                    px = tile_x0 + lx; py = tile_y0 + ly;
                    slot = tile_slot * FTB_TILE_PIXELS + ly * FTB_TILE_W + lx;
                    unit = (long long)py * F.gw + px;
                } else {
                    unit = (long long)tile_slot * FTB_TILE_PIXELS + j;
                    slot = (int)unit;
                }
                have_pixel = true; need = false;
                s = 0;
                pixsum = mk<R>(R(0), R(0), R(0));
            }
            tile_pos += min(__popc(m), avail);
            m = __ballot_sync(full, need);
        }
        if (!active && have_pixel) {  // start the next primary sample
            const unsigned long long sampleIndex = (unsigned long long)unit * (unsigned)spp + (unsigned)s;
            if (F.mode == 0) ray = primaryRay(F, px, py, s, sampleIndex);
            else {
                const double* q = F.rays + 6 * unit;
                ray.o = mk<R>((R)__ldg(q), (R)__ldg(q + 1), (R)__ldg(q + 2));
                ray.d = mk<R>((R)__ldg(q + 3), (R)__ldg(q + 4), (R)__ldg(q + 5));
            }
            active = true; depth = 0; limit = F.recursion_limit; weight = R(1);
            scol = mk<R>(R(0), R(0), R(0));
            cn.add(ST_PRIMARY);
        }
        if (!__any_sync(full, active)) break;

        // ---- one generation for every active lane ------------------------------------------------------------
        if (active) {
            const unsigned long long sampleIndex = (unsigned long long)unit * (unsigned)spp + (unsigned)s;
            Ray<R> off;  // slightOffset (Shading.fs:129): d is NOT normalised
            off.o = ray.o + R(0.0001) * ray.d;
            off.d = ray.d;
            const HitInfo<R> h = traceNearest<R, STATS>(S, off, overflow, cn);
            if (depth == 0 && F.dbg_prim) {
                int prim = -1, sub = 0;
                if (h.leaf >= 0) {
                    const int4 meta = __ldg(S.leaf_meta + h.leaf);
                    const int kind = meta.x & 0xff;
                    prim = meta.z;
                    sub = (kind == LEAF_CUBE || kind == LEAF_MESH) ? h.sub : ((kind == LEAF_TRIANGLE) ? 0 : meta.w);
                }
                F.dbg_prim[sampleIndex] = prim;
                if (F.dbg_sub) F.dbg_sub[sampleIndex] = sub;
                if (F.dbg_t) F.dbg_t[sampleIndex] = h.leaf >= 0 ? (double)h.t : -1.0;
            }
            bool cont = false;
            if (h.leaf >= 0 && S.n_lights > 0) {
                cn.add(ST_SHADED);
                const Fragment<R> f = finalise(S, off, h);
                const Vec<R> local = shadeLocal<R, STATS>(S, f, ray.d, F.seed, sampleIndex, depth, overflow, cn);
                scol = scol + weight * local;
                // reflectionShader (Shading.fs:89-98) sits inside the per-light sum: L identical re-traces
                if (f.applyLighting && f.reflectance > R(0) && limit > 0) {
                    weight = weight * ((R)S.n_lights * f.reflectance);
                    const Vec<R> rd = reflect(f.n, ray.d);
                    ray.o = f.p; ray.d = rd;
                    --limit; ++depth;
                    cont = true;
                    cn.add(ST_REFLECTION);
                }
            }
            active = cont;
        }
    }
    if (overflow) atomicExch(F.overflow, 1u);
    if constexpr (STATS) {
        for (int i = 0; i < ST_COUNT; ++i) {
            unsigned long long v = cn.get(i);
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
            if (lane == 0 && v) atomicAdd(F.stats + i, v);
        }
    }
}

template <typename R>
cudaError_t launch_render_impl(const DevScene<R>& s, const DevFrame<R>& f, bool stats, int sm_count, cudaStream_t stream, int* launches)
{
    int per_sm = 0;
    cudaError_t e;
    if (stats) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel<R, true>, kBlockThreads, 0);
    else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel<R, false>, kBlockThreads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    // persistent grid: a multiple of the SM count, never more warps than there are tiles to hand out
    long long want = (long long)sm_count * per_sm;
    long long cap = ((long long)f.n_local_tiles + (kBlockThreads / 32) - 1) / (kBlockThreads / 32);
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    if (stats) render_kernel<R, true><<<grid, kBlockThreads, 0, stream>>>(s, f);
    else render_kernel<R, false><<<grid, kBlockThreads, 0, stream>>>(s, f);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace ftb
