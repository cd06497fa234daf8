// ftb_oracle.cpp — CPU restatement of FuncTracer's per-pixel render loop.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/, the smoke test in
// __graft_entry__.py and bench.py's cpu_baseline / `--impl reference` legs may load it.  The
// product (functracer_b200/csrc) never links, imports or calls anything in oracle/.
//
// Parity status: the reference (F#, netcoreapp2.0) cannot be built or run in this image (no
// dotnet/mono/fsc), so the oracle is pinned against (a) every known-answer fact the
// reference's own tests hold for this path (FuncTracer.Tests/Geometry/BoundingBox.fs:11-27,
// Sphere.fs:18-21) and (b) the hand-derived vectors of SURVEY.md Appendix D.  The reference
// has no image-level golden vectors at all, so image-level parity is "unpinned by the
// reference's tests" (SURVEY.md section 8c); see DESIGN.md.
//
// Every function cites the reference file:line it restates.  All arithmetic is IEEE double,
// expression order is the F# source order, and the build uses -ffp-contract=off because
// RyuJIT on .NET Core 2.0 does not contract a*b+c.  `x ** 2.0` is restated as x*x
// (SURVEY.md 8c probe; documented grazing source).
//
// The scene arrives as the reference's own SceneGraph DU serialised by include/functracer_b200.h
// and is evaluated *faithfully*: nested transform nodes are applied level by level
// (Transform.fs:84-86), surface ops are maps over hits, CSG sorts and toggles, nearest hit is
// "stable sort by t, skip t<0, head".
#include "../include/functracer_b200.h"
#include "../include/ftb_rng.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---- CommonTypes.fs -----------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline V3 vadd(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // CommonTypes.fs:5-6
inline V3 vscale(V3 v, double s) { return {s * v.x, s * v.y, s * v.z}; }  // :7-10
inline V3 vneg(V3 v) { return {-v.x, -v.y, -v.z}; }                       // :11-12
inline V3 vsub(V3 a, V3 b) { return vadd(a, vneg(b)); }                   // :13-14  v1 + -v2
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // :15-16
inline V3 cross(V3 a, V3 b)                                                  // :17-18
{
    return {a.y * b.z - a.z * b.y, b.x * a.z - b.z * a.x, a.x * b.y - a.y * b.x};
}
inline double length(V3 v) { return std::sqrt(dot(v, v)); }  // :19
inline V3 psub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // Point - Point :34-35
inline V3 normalise(V3 v)  // :63-67
{
    double l = length(v);
    if (l < 0.0000001) return v;
    return vscale(v, 1.0 / l);
}
inline V3 reflect(V3 n, V3 v) { return vsub(v, vscale(n, 2.0 * dot(v, n))); }  // :72
inline double angleBetween(V3 a, V3 b) { return std::acos(dot(normalise(a), normalise(b))); }  // :74-75
inline V3 perpendicularComponent(V3 a, V3 b)  // :77-79
{
    V3 na = normalise(a);
    return vsub(b, vscale(na, dot(b, na)));
}
const double kPi = 3.14159265358979323846;  // System.Math.PI
inline double degToRad(double d) { return d * 1.0 * (kPi / 180.0); }  // CommonTypes.fs:98-99

struct Col {
    double r, g, b;
};
inline Col cadd(Col a, Col b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }  // :44-45
inline Col cmul(Col a, Col b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }  // :46-47
inline Col scaleColour(double i, Col c) { return {i * c.r, i * c.g, i * c.b}; }  // Image.fs:25-26

// ---- Transform.fs ---------------------------------------------------------------------------
struct M34 {
    double m[12];  // row-major 3x4
};
inline V3 mulVec(const double* m, V3 v)  // Transform.fs:15-18
{
    return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
            m[8] * v.x + m[9] * v.y + m[10] * v.z};
}
inline V3 mulPoint(const double* m, V3 p)  // :19-22
{
    return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
            m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}
// normalToWorld = transpose(worldToModel) applied to a Vector (:73-78, 15-18)
inline V3 mulVecT(const double* m, V3 v)
{
    return {m[0] * v.x + m[4] * v.y + m[8] * v.z, m[1] * v.x + m[5] * v.y + m[9] * v.z,
            m[2] * v.x + m[6] * v.y + m[10] * v.z};
}
// Transform.matrix for the basic transforms (:55-69); used for the primitives' own internal
// transforms (cube faces, solidCylinder caps).  Scene-level matrices come from the host.
M34 matTranslate(double x, double y, double z) { return {{1, 0, 0, x, 0, 1, 0, y, 0, 0, 1, z}}; }
M34 matRotate(V3 axis, double angle)
{
    V3 u = normalise(axis);  // Transform.fs:37-38
    double c = std::cos(angle), invc = 1.0 - c, s = std::sin(angle);
    return {{c + invc * u.x * u.x, invc * u.x * u.y - s * u.z, invc * u.x * u.z + s * u.y, 0.0,
             invc * u.x * u.y + s * u.z, c + invc * u.y * u.y, invc * u.y * u.z - s * u.x, 0.0,
             invc * u.x * u.z - s * u.y, invc * u.y * u.z + s * u.x, c + invc * u.z * u.z, 0.0}};
}
struct Xf {
    M34 m2w, w2m;
};
Xf xfTranslate(double x, double y, double z) { return {matTranslate(x, y, z), matTranslate(-x, -y, -z)}; }  // :48
Xf xfRotate(V3 axis, double angle) { return {matRotate(axis, angle), matRotate(axis, -angle)}; }            // :50

// ---- Ray.fs -----------------------------------------------------------------------------------
struct Material {  // Ray.fs:4-10
    Col colour;
    double roughness, reflectance, shineyness;
    bool applyLighting;
};
const Material mattWhite = {{1.0, 1.0, 1.0}, 0.0, 0.0, 0.0, true};  // :11
struct Ray {
    V3 o, d;
};
struct Hit {  // RayIntersection :21-27 (+ provenance for the debug planes)
    double t;
    V3 p, n;
    Material material;
    double u, v;
    int32_t prim, sub;
};
inline Hit newIntersection() { return {0.0, {0, 0, 0}, {1, 0, 0}, mattWhite, 0.0, 0.0, -1, 0}; }  // :29
typedef std::vector<Hit> Hits;

struct Counters {
    uint64_t primary = 0, shadow = 0, reflection = 0, shaded = 0;
    uint64_t leaf[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t xformed = 0, bsp_nodes = 0, csg_ops = 0;
    void add(const Counters& o)
    {
        primary += o.primary; shadow += o.shadow; reflection += o.reflection; shaded += o.shaded;
        for (int i = 0; i < 10; ++i) leaf[i] += o.leaf[i];
        xformed += o.xformed; bsp_nodes += o.bsp_nodes; csg_ops += o.csg_ops;
    }
};

// ---- Math.fs ----------------------------------------------------------------------------------
// quadratic (Math.fs:4-10): far ("+") root first.
inline int quadratic(double a, double b, double c, double out[2])
{
    double discriminant = b * b - 4.0 * a * c;  // b ** 2.0 restated as b*b
    if (discriminant < 0.0) return 0;
    double sq = std::sqrt(discriminant);
    double twoa = 2.0 * a;
    out[0] = (-b + sq) / twoa;
    out[1] = (-b - sq) / twoa;
    return 2;
}
inline double clamp01(double x) { return x > 1.0 ? 1.0 : (x < 0.0 ? 0.0 : x); }  // Math.fs:12-16

// F# `max`/`min` on floats are NaN-propagating (Math.Max/Min), SURVEY.md A.4.
inline double fsmax(double a, double b) { return (a != a || b != b) ? std::numeric_limits<double>::quiet_NaN() : (a < b ? b : a); }
inline double fsmin(double a, double b) { return (a != a || b != b) ? std::numeric_limits<double>::quiet_NaN() : (a < b ? a : b); }

// ---- primitives ---------------------------------------------------------------------------------
// Plane.intersect (Plane.fs:9-20) + plane's setUV (:28-33), for Plane(Point.Zero, unitY).
inline void planeHits(const Ray& r, Hits& out)
{
    const double eps = 0.0000001;
    const V3 p0 = {0, 0, 0}, n = {0, 1, 0};
    double num = dot(psub(p0, r.o), n);
    double denom = dot(r.d, n);
    Hit h = newIntersection();
    if (std::fabs(denom) < eps) {
        if (num < eps) {
            h.t = 0.0; h.p = r.o; h.n = n;
        } else
            return;
    } else {
        double t = num / denom;
        h.t = t; h.p = vadd(r.o, vscale(r.d, t)); h.n = n;
    }
    h.u = h.p.x; h.v = h.p.z;
    out.push_back(h);
}
// Cube.square (Cube.fs:9-15)
inline void squareHits(const Ray& r, Hits& out)
{
    size_t s = out.size();
    planeHits(r, out);
    if (out.size() > s) {
        const V3& p = out.back().p;
        if (!((p.x >= 0.0) && (p.x <= 1.0) && (p.z >= 0.0) && (p.z <= 1.0))) out.pop_back();
    }
}
// Cylinder.circle (Cylinder.fs:22)
inline void circleHits(const Ray& r, Hits& out)
{
    size_t s = out.size();
    planeHits(r, out);
    if (out.size() > s) {
        if (!(length(psub(out.back().p, V3{0, 0, 0})) < 1.0)) out.pop_back();
    }
}
// Sphere.sphere (Sphere.fs:6-21)
inline void sphereHits(const Ray& r, Hits& out)
{
    V3 ov = r.o;
    double a = dot(r.d, r.d);
    double b = 2.0 * dot(ov, r.d);
    double c = dot(ov, ov) - 1.0;
    double ts[2];
    int n = quadratic(a, b, c, ts);
    for (int i = 0; i < n; ++i) {
        Hit h = newIntersection();
        h.t = ts[i];
        h.p = vadd(r.o, vscale(r.d, ts[i]));
        h.n = normalise(h.p);
        h.u = 0.5 + (std::atan2(h.n.z, h.n.x) / (2.0 * kPi));
        h.v = 0.5 - std::asin(h.n.y) / kPi;
        out.push_back(h);
    }
}
// Cylinder.cylinder (Cylinder.fs:8-20)
inline void cylinderHits(const Ray& r, Hits& out)
{
    double ox = r.o.x, oz = r.o.z, dx = r.d.x, dz = r.d.z;
    double a = dx * dx + dz * dz;
    double b = 2.0 * (ox * dx + oz * dz);
    double c = ox * ox + oz * oz - 1.0;
    double ts[2];
    int n = quadratic(a, b, c, ts);
    for (int i = 0; i < n; ++i) {
        V3 p = vadd(r.o, vscale(r.d, ts[i]));
        V3 nn = normalise(V3{p.x, 0.0, p.z});
        Hit h = newIntersection();
        h.t = ts[i]; h.p = p;
        h.n = (dot(nn, r.d) < 0.0) ? nn : vneg(nn);
        if (p.y >= 0.0 && p.y <= 1.0) out.push_back(h);
    }
}
// Cone.cone (Cone.fs:7-28)
inline void coneHits(const Ray& r, Hits& out)
{
    double ox = r.o.x, oy = r.o.y, oz = r.o.z, dx = r.d.x, dy = r.d.y, dz = r.d.z;
    oy = oy - 1.0;
    double a = dx * dx + dz * dz - dy * dy;
    double b = 2.0 * (ox * dx + oz * dz - oy * dy);
    double c = ox * ox + oz * oz - oy * oy;
    double ts[2];
    int n = quadratic(a, b, c, ts);
    for (int i = 0; i < n; ++i) {
        V3 q = vadd(V3{ox, oy, oz}, vscale(r.d, ts[i]));
        V3 p = {q.x, q.y + 1.0, q.z};
        V3 nn = normalise(V3{q.x, -q.y, q.z});
        Hit h = newIntersection();
        h.t = ts[i]; h.p = p;
        h.n = (dot(nn, r.d) < 0.0) ? nn : vneg(nn);
        if (p.y >= 0.0 && p.y <= 1.0) out.push_back(h);
    }
}
// Triangle.fs:43-66 (Moller-Trumbore as written, including operator precedence of `.*`)
inline void triangleHits(const double* tri, const Ray& ray, Hits& out)
{
    const double epsilon = 0.0000001;
    V3 v0 = {tri[0], tri[1], tri[2]}, v1 = {tri[3], tri[4], tri[5]}, v2 = {tri[6], tri[7], tri[8]};
    V3 edge1 = psub(v1, v0), edge2 = psub(v2, v0);
    V3 h = cross(ray.d, edge2);
    double a = dot(edge1, h);
    if (a > -epsilon && a < epsilon) return;
    double f = 1.0 / a;
    V3 s = psub(ray.o, v0);
    double u = f * dot(s, h);
    if (u < 0.0 || u > 1.0) return;
    V3 q = cross(s, edge1);
    double v = dot(vscale(ray.d, f), q);  // f * ray.d.*q  ==  (f * ray.d) .* q
    if (v < 0.0 || u + v > 1.0) return;
    double t = dot(vscale(edge2, f), q);  // f * edge2.*q
    if (t > epsilon) {
        Hit hit = newIntersection();
        hit.t = t;
        hit.p = vadd(ray.o, vscale(normalise(ray.d), t * length(ray.d)));
        hit.n = normalise(cross(edge1, edge2));
        out.push_back(hit);
    }
}
// BoundingBox.intersects (BoundingBox.fs:32-58)
inline bool aabbIntersects(const double* bmin, const double* bmax, const Ray& ray)
{
    const double inf = std::numeric_limits<double>::infinity();
    double t0 = -inf, t1 = inf;
    const double* bounds[2] = {bmin, bmax};
    V3 inv = {1.0 / ray.d.x, 1.0 / ray.d.y, 1.0 / ray.d.z};
    int sign[3] = {inv.x < 0.0 ? 1 : 0, inv.y < 0.0 ? 1 : 0, inv.z < 0.0 ? 1 : 0};
    double tmin = (bounds[sign[0]][0] - ray.o.x) * inv.x;
    double tmax = (bounds[1 - sign[0]][0] - ray.o.x) * inv.x;
    double tymin = (bounds[sign[1]][1] - ray.o.y) * inv.y;
    double tymax = (bounds[1 - sign[1]][1] - ray.o.y) * inv.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    tmin = fsmax(tymin, tmin);
    tmax = fsmin(tymax, tmax);
    double tzmin = (bounds[sign[2]][2] - ray.o.z) * inv.z;
    double tzmax = (bounds[1 - sign[2]][2] - ray.o.z) * inv.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    tmin = fsmax(tzmin, tmin);
    tmax = fsmin(tzmax, tmax);
    return (tmin < t1) && (tmax > t0);
}

// Transform.transform (Transform.fs:80-87) around a callable producing hits.
template <class F>
inline void withTransform(const double* m2w, const double* w2m, const Ray& r, Hits& out, F&& object)
{
    Ray r2 = {mulPoint(w2m, r.o), mulVec(w2m, r.d)};
    size_t s = out.size();
    object(r2, out);
    for (size_t i = s; i < out.size(); ++i) {
        out[i].p = mulPoint(m2w, out[i].p);
        out[i].n = normalise(mulVecT(w2m, out[i].n));
    }
}
inline void flipNormals(Hits& out, size_t from)  // Ray.fs:36
{
    for (size_t i = from; i < out.size(); ++i) out[i].n = vscale(out[i].n, -1.0);
}

// Internal constant transforms of the composite primitives.
struct Consts {
    Xf top;         // translate (0,1,0)                     Cube.fs:19, Cylinder.fs:26
    Xf left;        // rotate unitZ 90deg                    Cube.fs:20
    Xf rightShift;  // translate unitX                       Cube.fs:21
    Xf front;       // rotate unitX -90deg                   Cube.fs:22
    Xf backShift;   // translate unitZ                       Cube.fs:23
    Xf centre;      // translate (-.5,-.5,-.5)               Cube.fs:25
    Xf bottomCap;   // rotate unitZ 180deg                   Cylinder.fs:27
    Consts()
    {
        top = xfTranslate(0.0, 1.0, 0.0);
        left = xfRotate(V3{0, 0, 1}, degToRad(90.0));
        rightShift = xfTranslate(1.0, 0.0, 0.0);
        front = xfRotate(V3{1, 0, 0}, degToRad(-90.0));
        backShift = xfTranslate(0.0, 0.0, 1.0);
        centre = xfTranslate(-0.5, -0.5, -0.5);
        bottomCap = xfRotate(V3{0, 0, 1}, degToRad(180.0));
    }
};
const Consts K;

inline void tagSub(Hits& out, size_t from, int sub)
{
    for (size_t i = from; i < out.size(); ++i) out[i].sub = sub;
}
// Cube.cube (Cube.fs:17-25)
inline void cubeHits(const Ray& r, Hits& out)
{
    withTransform(K.centre.m2w.m, K.centre.w2m.m, r, out, [](const Ray& r1, Hits& o) {
        size_t s;
        auto left = [](const Ray& rr, Hits& oo) { withTransform(K.left.m2w.m, K.left.w2m.m, rr, oo, squareHits); };
        auto front = [](const Ray& rr, Hits& oo) { withTransform(K.front.m2w.m, K.front.w2m.m, rr, oo, squareHits); };
        s = o.size(); squareHits(r1, o); flipNormals(o, s); tagSub(o, s, 0);                                   // bottom
        s = o.size(); withTransform(K.top.m2w.m, K.top.w2m.m, r1, o, squareHits); tagSub(o, s, 1);             // top
        s = o.size(); left(r1, o); tagSub(o, s, 2);                                                           // left
        s = o.size(); withTransform(K.rightShift.m2w.m, K.rightShift.w2m.m, r1, o, left); flipNormals(o, s); tagSub(o, s, 3);  // right
        s = o.size(); front(r1, o); tagSub(o, s, 4);                                                          // front
        s = o.size(); withTransform(K.backShift.m2w.m, K.backShift.w2m.m, r1, o, front); flipNormals(o, s); tagSub(o, s, 5);   // back
    });
}
// Cylinder.solidCylinder (Cylinder.fs:25-29)
inline void solidCylinderHits(const Ray& r, Hits& out)
{
    size_t s = out.size();
    withTransform(K.top.m2w.m, K.top.w2m.m, r, out, circleHits); tagSub(out, s, 0);
    s = out.size();
    withTransform(K.bottomCap.m2w.m, K.bottomCap.w2m.m, r, out, circleHits); tagSub(out, s, 1);
    s = out.size();
    cylinderHits(r, out); tagSub(out, s, 2);
}

// ---- textures (Textures/Texture.fs, Textures/Image.fs:27-36) ----------------------------------
inline double repeatOne(double x)  // Texture.fs:9-11
{
    double a = std::fabs(x - std::floor(x));
    return (a < 0.0) ? 1.0 - a : a;
}
Col evalTexture(const ftb_scene_desc* d, int tex, double u, double v)
{
    const ftb_texture& t = d->textures[tex];
    switch (t.kind) {
    case FTB_TEX_SCALE:  // Texture.fs:14-16
        return evalTexture(d, t.inner, u / t.p[0], v / t.p[1]);
    case FTB_TEX_ROTATE: {  // Texture.fs:18-22: matrix (rotate unitY angle) * Vector(u,0,v) -> (x, z)
        double c = t.p[1], s = t.p[2];
        double x = c * u + 0.0 * 0.0 + s * v;
        double z = (-s) * u + 0.0 * 0.0 + c * v;
        return evalTexture(d, t.inner, x, z);
    }
    case FTB_TEX_GRID: {  // Texture.fs:24-29
        double ru = repeatOne(u), rv = repeatOne(v);
        Col c1 = {t.p[0], t.p[1], t.p[2]}, c2 = {t.p[3], t.p[4], t.p[5]};
        if (ru < 0.5 && rv < 0.5) return c1;
        if (ru < 0.5) return c2;
        if (ru > 0.5 && rv > 0.5) return c1;
        return c2;
    }
    case FTB_TEX_IMAGE: {  // Textures/Image.fs:27-36
        const ftb_image& im = d->images[t.image];
        double ru = repeatOne(u), rv = repeatOne(v);
        long x = (long)std::floor(ru * (double)im.width);
        long y = (long)std::floor(rv * (double)im.height);
        // A.8: repeat can return exactly 1.0 -> index one past the row/image; the reference reads
        // the next row or throws.  Clamped here and in the kernels; documented grazing case.
        if (x >= im.width) x = im.width - 1;
        if (y >= im.height) y = im.height - 1;
        if (x < 0) x = 0;
        if (y < 0) y = 0;
        long index = y * (3L * im.width) + (3L * x);
        return {(double)im.rgb24[index] / 255.0, (double)im.rgb24[index + 1] / 255.0,
                (double)im.rgb24[index + 2] / 255.0};
    }
    }
    return {0, 0, 0};
}

// ---- Csg.fs ---------------------------------------------------------------------------------------
enum IType { OutsideIntoA, OutsideIntoB, BIntoAB, AIntoAB, ABleaveA, ABleaveB, AIntoOutside, BIntoOutside };
enum Rule { Take, Discard, Flip };
inline Rule unionRules(IType t)  // Csg.fs:19-25
{
    switch (t) { case OutsideIntoA: case OutsideIntoB: case AIntoOutside: case BIntoOutside: return Take; default: return Discard; }
}
inline Rule subtractRules(IType t)  // :27-33
{
    switch (t) { case OutsideIntoA: return Take; case AIntoAB: return Flip; case ABleaveB: return Flip; case AIntoOutside: return Take; default: return Discard; }
}
inline Rule intersectRules(IType t)  // :35-44
{
    switch (t) { case BIntoAB: case AIntoAB: case ABleaveA: case ABleaveB: return Take; default: return Discard; }
}
inline Rule excludeRules(IType t)  // :46-55
{
    switch (t) { case BIntoAB: case AIntoAB: case ABleaveA: case ABleaveB: return Flip; default: return Take; }
}
inline IType getIntersectionType(bool hitA, bool inA, bool inB)  // :59-72
{
    if (hitA) {
        if (inA && inB) return ABleaveA;
        if (!inA && inB) return BIntoAB;
        if (inA && !inB) return AIntoOutside;
        return OutsideIntoA;
    }
    if (inA && inB) return ABleaveB;
    if (!inA && inB) return BIntoOutside;
    if (inA && !inB) return AIntoAB;
    return OutsideIntoB;
}

// ---- Scene.intersect (Scene.fs:67-104) as a recursive evaluator --------------------------------
struct Scene {
    const ftb_scene_desc* d;
    std::vector<int32_t> primCount;  // PRIMITIVE instances under each node (for depth-first ids)
};

int countPrims(const ftb_scene_desc* d, int node, std::vector<int32_t>& memo, int depth)
{
    if (node < 0 || node >= d->n_nodes || depth > 4096) return -1;
    if (memo[node] >= 0) return memo[node];
    const ftb_node& n = d->nodes[node];
    long c = 0;
    switch (n.kind) {
    case FTB_NODE_PRIMITIVE: c = 1; break;
    case FTB_NODE_TRANSFORM: case FTB_NODE_MATERIAL: case FTB_NODE_TEXTURE: case FTB_NODE_HUESHIFT: case FTB_NODE_IGNORELIGHT: {
        int k = countPrims(d, n.b, memo, depth + 1);
        if (k < 0) return -1;
        c = k;
        break;
    }
    case FTB_NODE_GROUP:
        if (n.b < 0 || n.a < 0 || n.a + n.b > d->n_children) return -1;
        for (int i = 0; i < n.b; ++i) { int k = countPrims(d, d->children[n.a + i], memo, depth + 1); if (k < 0) return -1; c += k; }
        break;
    case FTB_NODE_UNION: case FTB_NODE_INTERSECT: case FTB_NODE_SUBTRACT: case FTB_NODE_EXCLUDE: {
        int ka = countPrims(d, n.a, memo, depth + 1), kb = countPrims(d, n.b, memo, depth + 1);
        if (ka < 0 || kb < 0) return -1;
        c = (long)ka + kb;
        break;
    }
    default: return -1;
    }
    if (c > 0x7fffffff) return -1;
    memo[node] = (int32_t)c;
    return (int32_t)c;
}

void bspHits(const Scene& sc, int link, const Ray& r, Hits& out, Counters& cn)
{
    const ftb_scene_desc* d = sc.d;
    if (link < 0) {  // Leaf: group of triangles (BspMesh.fs:52-53)
        const ftb_bsp_leaf& lf = d->bsp_leaves[~link];
        for (int i = 0; i < lf.tri_count; ++i) {
            size_t s = out.size();
            triangleHits(d->triangles + 9 * (size_t)(lf.tri_first + i), r, out);
            cn.leaf[FTB_PRIM_TRIANGLE]++;
            tagSub(out, s, lf.tri_first + i);
        }
        return;
    }
    // BspMesh.intersect (BspMesh.fs:67-76): AABB gate, then right ++ left
    const ftb_bsp_node& n = d->bsp_nodes[link];
    cn.bsp_nodes++;
    if (aabbIntersects(n.aabb_min, n.aabb_max, r)) {
        bspHits(sc, n.right, r, out, cn);
        bspHits(sc, n.left, r, out, cn);
    }
}

void nodeHits(const Scene& sc, int node, int primBase, const Ray& r, Hits& out, Counters& cn)
{
    const ftb_scene_desc* d = sc.d;
    const ftb_node& n = d->nodes[node];
    size_t s = out.size();
    switch (n.kind) {
    case FTB_NODE_PRIMITIVE:  // intersectPrimitive (Scene.fs:20-30)
        cn.leaf[n.a]++;
        switch (n.a) {
        case FTB_PRIM_BSPMESH: cn.leaf[n.a]--; bspHits(sc, d->meshes[n.b].root, r, out, cn); break;
        case FTB_PRIM_CIRCLE: circleHits(r, out); break;
        case FTB_PRIM_SQUARE: squareHits(r, out); break;
        case FTB_PRIM_CUBE: cubeHits(r, out); break;
        case FTB_PRIM_SPHERE: sphereHits(r, out); break;
        case FTB_PRIM_PLANE: planeHits(r, out); break;
        case FTB_PRIM_CONE: coneHits(r, out); break;
        case FTB_PRIM_SOLIDCYLINDER: solidCylinderHits(r, out); break;
        case FTB_PRIM_CYLINDER: cylinderHits(r, out); break;
        case FTB_PRIM_TRIANGLE: triangleHits(d->triangles + 9 * (size_t)n.b, r, out); break;
        }
        for (size_t i = s; i < out.size(); ++i) out[i].prim = primBase;
        break;
    case FTB_NODE_TRANSFORM: {  // Transform.transform
        const ftb_transform& t = d->transforms[n.a];
        cn.xformed++;
        withTransform(t.m2w, t.w2m, r, out, [&](const Ray& r2, Hits& o) { nodeHits(sc, n.b, primBase, r2, o, cn); });
        break;
    }
    case FTB_NODE_MATERIAL: {  // Ray.setMaterial (Ray.fs:49)
        nodeHits(sc, n.b, primBase, r, out, cn);
        const ftb_material& m = d->materials[n.a];
        Material mm = {{m.colour[0], m.colour[1], m.colour[2]}, m.roughness, m.reflectance, m.shineyness, m.apply_lighting != 0};
        for (size_t i = s; i < out.size(); ++i) out[i].material = mm;
        break;
    }
    case FTB_NODE_TEXTURE:  // Ray.textureDiffuse (Ray.fs:57-59)
        nodeHits(sc, n.b, primBase, r, out, cn);
        for (size_t i = s; i < out.size(); ++i) out[i].material.colour = evalTexture(d, n.a, out[i].u, out[i].v);
        break;
    case FTB_NODE_HUESHIFT:  // Ray.hueShift (Ray.fs:51-55) -> Colour.hueShift (CommonTypes.fs:90)
        nodeHits(sc, n.b, primBase, r, out, cn);
        for (size_t i = s; i < out.size(); ++i) {
            Col c = out[i].material.colour;
            out[i].material.colour = {c.b, c.r, c.g};
        }
        break;
    case FTB_NODE_IGNORELIGHT:  // Ray.ignoreLight (Ray.fs:47)
        nodeHits(sc, n.b, primBase, r, out, cn);
        for (size_t i = s; i < out.size(); ++i) out[i].material.applyLighting = false;
        break;
    case FTB_NODE_GROUP: {  // Ray.group (Ray.fs:34)
        int base = primBase;
        for (int i = 0; i < n.b; ++i) {
            int ch = d->children[n.a + i];
            nodeHits(sc, ch, base, r, out, cn);
            base += sc.primCount[ch];
        }
        break;
    }
    default: {  // Csg.constructedSolid (Csg.fs:74-94)
        cn.csg_ops++;
        Hits merged;
        nodeHits(sc, n.a, primBase, r, merged, cn);
        size_t na = merged.size();
        nodeHits(sc, n.b, primBase + sc.primCount[n.a], r, merged, cn);
        std::vector<int> order(merged.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        // Seq.sortBy is a stable sort; insertion sort keeps NaN keys where they are.
        for (size_t i = 1; i < order.size(); ++i) {
            int x = order[i];
            size_t j = i;
            while (j > 0 && merged[x].t < merged[order[j - 1]].t) { order[j] = order[j - 1]; --j; }
            order[j] = x;
        }
        bool inA = false, inB = false;
        for (size_t i = 0; i < order.size(); ++i) {
            Hit h = merged[order[i]];
            bool hitA = (size_t)order[i] < na;
            IType it = getIntersectionType(hitA, inA, inB);
            Rule rule;
            switch (n.kind) {
            case FTB_NODE_UNION: rule = unionRules(it); break;
            case FTB_NODE_SUBTRACT: rule = subtractRules(it); break;
            case FTB_NODE_INTERSECT: rule = intersectRules(it); break;
            default: rule = excludeRules(it); break;
            }
            if (hitA) inA = !inA; else inB = !inB;
            if (rule == Take) out.push_back(h);
            else if (rule == Flip) { h.n = vscale(h.n, -1.0); out.push_back(h); }
        }
        break;
    }
    }
}

// closest (Scene.fs:112-116): stable sort by t, skip t<0, head == first-in-order minimum t>=0.
inline const Hit* closest(const Hits& hs)
{
    const Hit* best = nullptr;
    for (const Hit& h : hs)
        if (h.t >= 0.0 && (!best || h.t < best->t)) best = &h;
    return best;
}
// lightIsBocked (Scene.fs:119-121)
inline bool lightIsBlocked(const Scene& sc, double maxDistance, const Ray& r, Counters& cn)
{
    cn.shadow++;
    Hits hs;
    nodeHits(sc, sc.d->root, 0, r, hs, cn);
    for (const Hit& i : hs)
        if (i.t >= 0.0 && i.t < maxDistance && i.material.applyLighting) return true;
    return false;
}

// ---- Jitter.fs -----------------------------------------------------------------------------------
struct RngKey {
    uint64_t seed, sample;
    uint32_t depth, light;
};
// Jitter.circle (Jitter.fs:15-21) on the ftb_rng contract.
inline void jitterCircle(const RngKey& k, uint32_t idx, double& x, double& y)
{
    for (uint32_t attempt = 0;; ++attempt) {
        x = FTB_RNG_TO_UNIT(ftb_rng_bits24(k.seed, k.sample, k.depth, k.light, idx, attempt, 0));
        y = FTB_RNG_TO_UNIT(ftb_rng_bits24(k.seed, k.sample, k.depth, k.light, idx, attempt, 1));
        bool outsideCircle = (x * x + y * y) > 1.0;
        if (!outsideCircle) return;
    }
}
// Jitter.jitterVector (Jitter.fs:26-39), sample idx of `count`
inline V3 jitterVector(const RngKey& k, uint32_t idx, double maxAngle, V3 vector)
{
    V3 normalised = normalise(vector);
    double maxOffsetMagnitude = std::tan(maxAngle / 2.0);
    V3 generator = (normalised.x > 0.9) ? V3{0, 1, 0} : V3{1, 0, 0};
    V3 i = normalise(cross(generator, normalised));
    V3 j = cross(i, normalised);
    double x, y;
    jitterCircle(k, idx, x, y);
    return normalise(vadd(vadd(normalised, vscale(i, maxOffsetMagnitude * x)), vscale(j, maxOffsetMagnitude * y)));
}

// ---- Shading.fs ------------------------------------------------------------------------------------
inline double attenuate(const double f[3], double distance) { return 1.0 / (f[0] + distance * (f[1] + distance * f[2])); }  // Light.fs:16-17

struct ShadeCtx {
    const Scene* sc;
    uint64_t seed;
    Counters* cn;
};

// shadowLightIntensity / softShadowLightIntensity (Shading.fs:24-42)
double shadowLightIntensity(const ShadeCtx& cx, const ftb_light& L, int lightIdx, V3 point, uint64_t sample, uint32_t depth)
{
    const double dblMax = std::numeric_limits<double>::max();
    V3 v = {L.v[0], L.v[1], L.v[2]};
    switch (L.kind) {
    case FTB_LIGHT_DIRECTIONAL:
        return lightIsBlocked(*cx.sc, dblMax, Ray{point, vneg(v)}, *cx.cn) ? 0.0 : 1.0;
    case FTB_LIGHT_SOFT_DIRECTIONAL: {
        int occluded = 0;
        RngKey key = {cx.seed, sample, depth, (uint32_t)lightIdx};
        for (int k = 0; k < L.samples; ++k) {
            V3 dir = jitterVector(key, (uint32_t)k, L.scatter_rad, vneg(v));
            if (lightIsBlocked(*cx.sc, dblMax, Ray{point, dir}, *cx.cn)) ++occluded;
        }
        return (double)(L.samples - occluded) / (double)L.samples;
    }
    default: {
        V3 dvec = psub(v, point);
        double distance = length(dvec);
        if (lightIsBlocked(*cx.sc, distance, Ray{point, normalise(dvec)}, *cx.cn)) return 0.0;
        return attenuate(L.falloff, distance);
    }
    }
}
inline V3 lightDirection(const ftb_light& L, V3 atPoint)  // Shading.fs:44-48
{
    V3 v = {L.v[0], L.v[1], L.v[2]};
    if (L.kind == FTB_LIGHT_POINT) return normalise(psub(atPoint, v));
    return v;
}
inline Col roughDiffuse(const Hit& ix, V3 lightDir, const Ray& viewRay)  // Shading.fs:50-63
{
    double roughness = ix.material.roughness * ix.material.roughness;  // ** 2.0
    double rayAngle = angleBetween(ix.n, vneg(viewRay.d));
    double lightAngle = angleBetween(ix.n, vneg(lightDir));
    double alpha = fsmax(rayAngle, lightAngle);
    double beta = fsmin(rayAngle, lightAngle);
    double A = 1.0 - 0.5 * roughness / (roughness + 0.33);
    double B = 0.45 * roughness / (roughness + 0.09);
    V3 tangentLight = normalise(perpendicularComponent(ix.n, vneg(lightDir)));
    V3 tangentRay = normalise(perpendicularComponent(ix.n, vneg(viewRay.d)));
    double intensity = std::cos(lightAngle) * (A + (B * fsmax(0.0, dot(tangentLight, tangentRay)) * std::sin(alpha) * std::tan(beta)));
    return scaleColour(intensity, ix.material.colour);
}
inline Col lambertianDiffuse(const Hit& ix, Col lightColour, V3 lightDir)  // :65-70
{
    double intensity = dot(vneg(lightDir), ix.n);
    return scaleColour(intensity, cmul(ix.material.colour, lightColour));
}
inline Col specularShader(const Hit& ix, Col lightColour, V3 lightDir, const Ray& viewRay)  // :78-87
{
    V3 normal = normalise(ix.n);
    double shineyness = ix.material.shineyness;
    V3 reflectedLightDirection = normalise(reflect(normal, lightDir));
    V3 viewDirection = normalise(viewRay.d);
    double intensity = std::pow(dot(viewDirection, vneg(reflectedLightDirection)), shineyness);
    if (shineyness <= 0.0 || intensity <= 0.0) return {0, 0, 0};
    return {lightColour.r * intensity, lightColour.g * intensity, lightColour.b * intensity};
}

struct PrimaryInfo {
    int32_t prim, sub;
    double t;
};

// getColourForRay (Shading.fs:131-139).  The reflection colour is the same for every light's
// fragment (the RNG key does not depend on the light that spawned the re-trace), so it is
// traced once and added once per light, in the reference's summation order.
Col getColourForRay(const ShadeCtx& cx, int recursionLimit, const Ray& ray, uint64_t sample, uint32_t depth, PrimaryInfo* info)
{
    const ftb_scene_desc* d = cx.sc->d;
    Ray off = {vadd(ray.o, vscale(ray.d, 0.0001)), ray.d};  // slightOffset :129
    Hits hs;
    nodeHits(*cx.sc, d->root, 0, off, hs, *cx.cn);
    const Hit* hp = closest(hs);
    if (info) {
        info->prim = hp ? hp->prim : -1;
        info->sub = hp ? hp->sub : 0;
        info->t = hp ? hp->t : -1.0;
    }
    Col total = {0, 0, 0};
    if (!hp) return total;
    Hit ix = *hp;
    if (d->n_lights > 0) cx.cn->shaded++;
    // getLightsOnPoint :109-117
    V3 shadowRayOrigin = vadd(ix.p, vscale(ix.n, 0.0001));
    bool haveRefl = false;
    Col refl = {0, 0, 0};
    for (int li = 0; li < d->n_lights; ++li) {
        const ftb_light& L = d->lights[li];
        double intensity = shadowLightIntensity(cx, L, li, shadowRayOrigin, sample, depth);
        Col lightColour = scaleColour(intensity, Col{L.colour[0], L.colour[1], L.colour[2]});
        V3 ldir = lightDirection(L, ix.p);
        Col frag;
        if (!ix.material.applyLighting) {  // shadeIfRequired :100-104
            frag = ix.material.colour;
        } else {  // multiPartShader [specular; reflection; diffuse] :105-107, Program.fs:59
            Col acc = {0, 0, 0};
            acc = cadd(acc, specularShader(ix, lightColour, ldir, ray));
            Col r = {0, 0, 0};  // reflectionShader :89-98
            if (ix.material.reflectance > 0.0) {
                if (!haveRefl) {
                    haveRefl = true;
                    if (recursionLimit <= 0) refl = {0, 0, 0};  // :133
                    else {
                        V3 reflectedDirection = reflect(ix.n, ray.d);
                        cx.cn->reflection++;
                        refl = getColourForRay(cx, recursionLimit - 1, Ray{ix.p, reflectedDirection}, sample, depth + 1, nullptr);
                    }
                }
                r = {refl.r * ix.material.reflectance, refl.g * ix.material.reflectance, refl.b * ix.material.reflectance};
            }
            acc = cadd(acc, r);
            Col diff = (ix.material.roughness == 0.0) ? lambertianDiffuse(ix, lightColour, ldir)  // :72-76
                                                      : roughDiffuse(ix, ldir, ray);
            acc = cadd(acc, diff);
            frag = acc;
        }
        total = cadd(total, frag);
    }
    return total;
}

// ---- Image.fs (sampling half) ---------------------------------------------------------------------
struct ImagePlane {
    V3 origin, k, i, j;
    double pw, ph, tlx, tly;
};
ImagePlane createImagePlane(const ftb_camera& c, int resH, int resV)  // Image.fs:48-53, 67-81
{
    V3 o = {c.o[0], c.o[1], c.o[2]}, la = {c.look_at[0], c.look_at[1], c.look_at[2]}, up = {c.up[0], c.up[1], c.up[2]};
    V3 k = normalise(psub(la, o));
    V3 i = normalise(cross(up, k));
    V3 j = cross(k, i);
    double height = std::tan(c.fov_y_rad / 2.0) * 2.0;
    double width = height * c.aspect_ratio;
    double pixelHeight = height / (double)(resH - 1);
    double pixelWidth = width / (double)(resV - 1);
    return {o, k, i, j, pixelWidth, pixelHeight, -width / 2.0 + pixelWidth / 2.0, height / 2.0 - pixelHeight / 2.0};
}
inline Ray rayThroughPixel(const ImagePlane& ip, int px, int py, double jitterX, double jitterY)  // :83-89
{
    double centreX = ip.tlx + (double)px * ip.pw, centreY = ip.tly - (double)py * ip.ph;
    double jx = centreX + jitterX * ip.pw, jy = centreY + jitterY * ip.ph;
    V3 via = vadd(vadd(ip.k, vscale(ip.i, jx)), vscale(ip.j, jy));
    return {ip.origin, via};
}
inline Ray depthOfFieldJitter(const ftb_camera& c, const Ray& r0, uint64_t seed, uint64_t sample)  // :91-94, Ray.fs:15-18
{
    Ray r = {vadd(r0.o, vscale(r0.d, c.focal_length)), r0.d};
    RngKey key = {seed, sample, 0, FTB_RNG_STREAM_CAMERA};
    r.d = jitterVector(key, 0, c.aperture_rad, r.d);
    r.o = vadd(r.o, vscale(r.d, -c.focal_length));
    return r;
}

thread_local char g_err[256] = "";
int fail(int code, const char* msg)
{
    std::strncpy(g_err, msg, sizeof(g_err) - 1);
    return code;
}

bool prepare(const ftb_scene_desc* d, Scene& sc)
{
    if (!d || !d->nodes || d->n_nodes <= 0) return false;
    sc.d = d;
    sc.primCount.assign(d->n_nodes, -1);
    return countPrims(d, d->root, sc.primCount, 0) >= 0;
}

// shade (Shading.fs:141-147): 1000-ray chunks handed to a thread pool, results in ray order.
// item(i) yields the i-th ray of the list, its RNG sample key (the ray's index in the
// reference's full-frame ray list) and the slot of the debug planes it reports into.
struct Item {
    Ray r;
    uint64_t key;
    int64_t slot;
};
template <class ItemFn>
void shadeAll(const Scene& sc, const ftb_render_params* p, int64_t n, ItemFn item, double* cols,
              const ftb_debug_out* dbg, Counters& total, int threads)
{
    std::atomic<int64_t> next(0);
    const int64_t chunk = 1000;
    int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::vector<Counters> cns(nthreads);
    auto work = [&](int tid) {
        ShadeCtx cx = {&sc, p->seed, &cns[tid]};
        for (;;) {
            int64_t b = next.fetch_add(chunk);
            if (b >= n) break;
            int64_t e = std::min(n, b + chunk);
            for (int64_t i = b; i < e; ++i) {
                Item it = item(i);
                PrimaryInfo info;
                cns[tid].primary++;
                Col c = getColourForRay(cx, p->recursion_limit, it.r, it.key, 0, &info);
                cols[3 * i] = c.r; cols[3 * i + 1] = c.g; cols[3 * i + 2] = c.b;
                if (dbg) {
                    if (dbg->prim_id) dbg->prim_id[it.slot] = info.prim;
                    if (dbg->sub_id) dbg->sub_id[it.slot] = info.sub;
                    if (dbg->t) dbg->t[it.slot] = info.t;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    for (auto& c : cns) total.add(c);
}

void exportCounters(const Counters& c, ftb_stats* s)
{
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    s->primary_rays = c.primary; s->shadow_rays = c.shadow; s->reflection_rays = c.reflection; s->shaded_hits = c.shaded;
    for (int i = 0; i < 10; ++i) s->leaf_tests[i] = c.leaf[i];
    s->transformed_leaf_tests = c.xformed; s->bsp_nodes_visited = c.bsp_nodes; s->csg_ops = c.csg_ops;
}

}  // namespace

extern "C" {

const char* ftbo_last_error(void) { return g_err; }

// Program.fs:54-64 on the CPU: generateRays (+ depthOfFieldJitter), shade, blendPixels.
// `threads` <= 0 uses every hardware thread.  out_rgb = W*H*3 doubles; dbg planes W*H*spp
// (jitter) or (W+1)*(H+1) (corner).  A pixel window [x0,x1)x[y0,y1) restricts the work (for
// timing a stated sample of a large frame); pixels outside it are left untouched.
int ftbo_render_window(const ftb_scene_desc* desc, const ftb_camera* cam, const ftb_render_params* p,
                       int x0, int y0, int x1, int y1, double* out_rgb, const ftb_debug_out* dbg,
                       ftb_stats* stats, int threads)
{
    Scene sc;
    if (!cam || !p || !out_rgb) return fail(FTB_ERR_BAD_ARG, "null argument");
    if (!prepare(desc, sc)) return fail(FTB_ERR_BAD_SCENE, "malformed scene graph");
    const int W = p->width, H = p->height;
    if (W < 1 || H < 1) return fail(FTB_ERR_BAD_ARG, "bad resolution");
    x0 = std::max(0, x0); y0 = std::max(0, y0); x1 = std::min(W, x1); y1 = std::min(H, y1);
    ImagePlane ip = createImagePlane(*cam, W, H);
    Counters total;
    if (p->sampling == FTB_SAMPLING_CORNER) {  // CornerSampling (Image.fs:125-150)
        const int stride = W + 1;
        const int cw = x1 - x0 + 1, chh = y1 - y0 + 1;  // corners the window needs
        if (x1 <= x0 || y1 <= y0) { exportCounters(total, stats); return FTB_OK; }
        std::vector<double> cols(3 * (size_t)cw * chh);
        auto item = [&](int64_t i) {
            int cx = x0 + (int)(i % cw), cy = y0 + (int)(i / cw);
            int64_t g = (int64_t)cy * stride + cx;  // index in the reference's ray list (:130)
            Ray r = rayThroughPixel(ip, cx, cy, -0.5, 0.5);  // :131
            if (cam->has_focus) r = depthOfFieldJitter(*cam, r, p->seed, (uint64_t)g);
            return Item{r, (uint64_t)g, g};
        };
        shadeAll(sc, p, (int64_t)cw * chh, item, cols.data(), dbg, total, threads);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {  // colourForPixel :138-141, Seq.average
                int lx = x - x0, ly = y - y0;
                int64_t corners[4] = {(int64_t)ly * cw + lx, (int64_t)ly * cw + lx + 1, (int64_t)(ly + 1) * cw + lx, (int64_t)(ly + 1) * cw + lx + 1};
                Col s = {0, 0, 0};
                for (int k = 0; k < 4; ++k) s = cadd(s, Col{cols[3 * corners[k]], cols[3 * corners[k] + 1], cols[3 * corners[k] + 2]});
                int64_t o = (int64_t)y * W + x;
                out_rgb[3 * o] = s.r / (double)4; out_rgb[3 * o + 1] = s.g / (double)4; out_rgb[3 * o + 2] = s.b / (double)4;
            }
    } else {  // JitteredSampling (Image.fs:97-122)
        const int spp = p->spp;
        if (spp < 1 || !p->jitter_xy) return fail(FTB_ERR_BAD_ARG, "jitter mode needs spp >= 1 and jitter_xy");
        const int ww = x1 - x0, wh = y1 - y0;
        if (ww <= 0 || wh <= 0) { exportCounters(total, stats); return FTB_OK; }
        const int64_t n = (int64_t)ww * wh * spp;
        std::vector<double> cols(3 * (size_t)n);
        auto item = [&](int64_t i) {
            int s = (int)(i % spp);
            int64_t pix = i / spp;
            int x = x0 + (int)(pix % ww), y = y0 + (int)(pix / ww);
            int64_t g = ((int64_t)y * W + x) * spp + s;  // index in the reference's ray list (:104-110)
            Ray r = rayThroughPixel(ip, x, y, p->jitter_xy[2 * s], p->jitter_xy[2 * s + 1]);
            if (cam->has_focus) r = depthOfFieldJitter(*cam, r, p->seed, (uint64_t)g);
            return Item{r, (uint64_t)g, g};
        };
        shadeAll(sc, p, n, item, cols.data(), dbg, total, threads);
        // blendPixels: Array.average = (fold (+) Zero) then DivideByInt (Image.fs:112-116)
        for (int64_t pix = 0; pix < (int64_t)ww * wh; ++pix) {
            Col s = {0, 0, 0};
            for (int k = 0; k < spp; ++k) s = cadd(s, Col{cols[3 * (pix * spp + k)], cols[3 * (pix * spp + k) + 1], cols[3 * (pix * spp + k) + 2]});
            int x = x0 + (int)(pix % ww), y = y0 + (int)(pix / ww);
            int64_t o = (int64_t)y * W + x;
            out_rgb[3 * o] = s.r / (double)spp; out_rgb[3 * o + 1] = s.g / (double)spp; out_rgb[3 * o + 2] = s.b / (double)spp;
        }
    }
    exportCounters(total, stats);
    return FTB_OK;
}

// Several pixel windows of one frame in ONE call (one thread pool, one list of 1000-ray chunks across all windows:
// bench.py's bounded CPU samples keep every host thread busy that way), with PACKED outputs: window w's pixels
// follow window w-1's, row-major inside the window; out_rgb holds 3 doubles per window pixel and the dbg planes spp
// entries per window pixel, so a few stripes of an 8K x 64 spp frame do not need full-frame (8.5 GB) planes.  Jitter
// sampling only.  rects = n_windows x (x0, y0, x1, y1).  Ray generation, RNG keys and the blend are those of
// ftbo_render_window, i.e. of the full frame (Image.fs:97-122).
int ftbo_render_windows_packed(const ftb_scene_desc* desc, const ftb_camera* cam, const ftb_render_params* p, int n_windows, const int* rects,
                               double* out_rgb, const ftb_debug_out* dbg, ftb_stats* stats, int threads)
{
    Scene sc;
    if (!cam || !p || !out_rgb || !rects || n_windows < 0) return fail(FTB_ERR_BAD_ARG, "null argument");
    if (!prepare(desc, sc)) return fail(FTB_ERR_BAD_SCENE, "malformed scene graph");
    const int W = p->width, H = p->height, spp = p->spp;
    if (W < 1 || H < 1) return fail(FTB_ERR_BAD_ARG, "bad resolution");
    if (p->sampling != FTB_SAMPLING_JITTER || spp < 1 || !p->jitter_xy) return fail(FTB_ERR_BAD_ARG, "packed windows need jitter sampling");
    ImagePlane ip = createImagePlane(*cam, W, H);
    std::vector<int64_t> first((size_t)n_windows + 1, 0);  // first packed pixel of every window
    std::vector<int> r(rects, rects + 4 * (size_t)n_windows);
    for (int w = 0; w < n_windows; ++w) {
        int& x0 = r[4 * w]; int& y0 = r[4 * w + 1]; int& x1 = r[4 * w + 2]; int& y1 = r[4 * w + 3];
        x0 = std::max(0, x0); y0 = std::max(0, y0); x1 = std::min(W, x1); y1 = std::min(H, y1);
        if (x1 < x0) x1 = x0;
        if (y1 < y0) y1 = y0;
        first[(size_t)w + 1] = first[(size_t)w] + (int64_t)(x1 - x0) * (y1 - y0);
    }
    const int64_t npix = first[(size_t)n_windows], n = npix * spp;
    Counters total;
    std::vector<double> cols(3 * (size_t)n);
    auto item = [&](int64_t i) {
        const int s = (int)(i % spp);
        const int64_t pix = i / spp;
        const int w = (int)(std::upper_bound(first.begin(), first.end(), pix) - first.begin()) - 1;
        const int x0 = r[4 * w], y0 = r[4 * w + 1], ww = r[4 * w + 2] - x0;
        const int64_t lp = pix - first[(size_t)w];
        const int x = x0 + (int)(lp % ww), y = y0 + (int)(lp / ww);
        const int64_t g = ((int64_t)y * W + x) * spp + s;  // index in the reference's ray list (:104-110): the RNG key
        Ray ray = rayThroughPixel(ip, x, y, p->jitter_xy[2 * s], p->jitter_xy[2 * s + 1]);
        if (cam->has_focus) ray = depthOfFieldJitter(*cam, ray, p->seed, (uint64_t)g);
        return Item{ray, (uint64_t)g, i};
    };
    shadeAll(sc, p, n, item, cols.data(), dbg, total, threads);
    for (int64_t pix = 0; pix < npix; ++pix) {  // Array.average (Image.fs:112-116)
        Col sum = {0, 0, 0};
        for (int k = 0; k < spp; ++k) sum = cadd(sum, Col{cols[3 * (pix * spp + k)], cols[3 * (pix * spp + k) + 1], cols[3 * (pix * spp + k) + 2]});
        out_rgb[3 * pix] = sum.r / (double)spp; out_rgb[3 * pix + 1] = sum.g / (double)spp; out_rgb[3 * pix + 2] = sum.b / (double)spp;
    }
    exportCounters(total, stats);
    return FTB_OK;
}

int ftbo_render(const ftb_scene_desc* desc, const ftb_camera* cam, const ftb_render_params* p, double* out_rgb,
                const ftb_debug_out* dbg, ftb_stats* stats, int threads)
{
    if (!p) return fail(FTB_ERR_BAD_ARG, "null params");
    return ftbo_render_window(desc, cam, p, 0, 0, p->width, p->height, out_rgb, dbg, stats, threads);
}

// Shading.shade on explicit rays (Shading.fs:141-147).
int ftbo_shade_rays(const ftb_scene_desc* desc, const double* rays_od, int64_t n, const ftb_render_params* p,
                    double* out_rgb, const ftb_debug_out* dbg, ftb_stats* stats, int threads)
{
    Scene sc;
    if (!rays_od || !p || !out_rgb || n < 0) return fail(FTB_ERR_BAD_ARG, "null argument");
    if (!prepare(desc, sc)) return fail(FTB_ERR_BAD_SCENE, "malformed scene graph");
    Counters total;
    auto item = [&](int64_t i) {
        const double* r = rays_od + 6 * i;
        return Item{Ray{{r[0], r[1], r[2]}, {r[3], r[4], r[5]}}, (uint64_t)i, i};
    };
    shadeAll(sc, p, n, item, out_rgb, dbg, total, threads);
    exportCounters(total, stats);
    return FTB_OK;
}

// ---- unit-level probes for the known-answer tests ---------------------------------------------
typedef struct ftbo_hit {
    double t, p[3], n[3], uv[2], colour[3];
    double roughness, reflectance, shineyness;
    int32_t apply_lighting, prim, sub, reserved;
} ftbo_hit;

// All hits of `node` for one ray, in the reference's sequence order (before `closest`).
int ftbo_node_hits(const ftb_scene_desc* desc, int node, const double* o, const double* d, ftbo_hit* out, int max_hits)
{
    Scene sc;
    if (!prepare(desc, sc)) return fail(FTB_ERR_BAD_SCENE, "malformed scene graph");
    if (node < 0) node = desc->root;
    // prim base of an inner node is not tracked here: ids are relative to `node`
    if (node >= desc->n_nodes) return fail(FTB_ERR_BAD_ARG, "bad node");
    if (countPrims(desc, node, sc.primCount, 0) < 0) return fail(FTB_ERR_BAD_SCENE, "bad node");
    Hits hs;
    Counters cn;
    nodeHits(sc, node, 0, Ray{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}}, hs, cn);
    int k = 0;
    for (const Hit& h : hs) {
        if (k >= max_hits) break;
        ftbo_hit& q = out[k++];
        q.t = h.t; q.p[0] = h.p.x; q.p[1] = h.p.y; q.p[2] = h.p.z; q.n[0] = h.n.x; q.n[1] = h.n.y; q.n[2] = h.n.z;
        q.uv[0] = h.u; q.uv[1] = h.v; q.colour[0] = h.material.colour.r; q.colour[1] = h.material.colour.g; q.colour[2] = h.material.colour.b;
        q.roughness = h.material.roughness; q.reflectance = h.material.reflectance; q.shineyness = h.material.shineyness;
        q.apply_lighting = h.material.applyLighting; q.prim = h.prim; q.sub = h.sub; q.reserved = 0;
    }
    return (int)hs.size();
}
int ftbo_quadratic(double a, double b, double c, double* out) { return quadratic(a, b, c, out); }
int ftbo_aabb_intersects(const double* bmin, const double* bmax, const double* o, const double* d)
{
    return aabbIntersects(bmin, bmax, Ray{{o[0], o[1], o[2]}, {d[0], d[1], d[2]}}) ? 1 : 0;
}
double ftbo_attenuate(const double* falloff, double distance) { return attenuate(falloff, distance); }
void ftbo_texture(const ftb_scene_desc* desc, int tex, double u, double v, double* rgb)
{
    Col c = evalTexture(desc, tex, u, v);
    rgb[0] = c.r; rgb[1] = c.g; rgb[2] = c.b;
}
uint8_t ftbo_to_byte(double c) { return (uint8_t)(clamp01(c) * 255.0); }  // Image.fs:36
void ftbo_hue_shift(const double* in, double* out) { out[0] = in[2]; out[1] = in[0]; out[2] = in[1]; }  // CommonTypes.fs:90
void ftbo_jitter_vector(uint64_t seed, uint64_t sample, uint32_t depth, uint32_t light, uint32_t idx, double max_angle, const double* v, double* out)
{
    V3 r = jitterVector(RngKey{seed, sample, depth, light}, idx, max_angle, V3{v[0], v[1], v[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// lambert / specular on a synthetic fragment (Appendix D 10, 11)
void ftbo_lambert(const double* n, const double* ld, const double* lc, const double* colour, double* out)
{
    Hit h = newIntersection();
    h.n = {n[0], n[1], n[2]};
    h.material.colour = {colour[0], colour[1], colour[2]};
    Col c = lambertianDiffuse(h, Col{lc[0], lc[1], lc[2]}, V3{ld[0], ld[1], ld[2]});
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
void ftbo_specular(const double* n, const double* ld, const double* lc, const double* view_d, double shineyness, double* out)
{
    Hit h = newIntersection();
    h.n = {n[0], n[1], n[2]};
    h.material.shineyness = shineyness;
    Col c = specularShader(h, Col{lc[0], lc[1], lc[2]}, V3{ld[0], ld[1], ld[2]}, Ray{{0, 0, 0}, {view_d[0], view_d[1], view_d[2]}});
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
void ftbo_rough_diffuse(const double* n, const double* ld, const double* view_d, const double* colour, double roughness, double* out)
{
    Hit h = newIntersection();
    h.n = {n[0], n[1], n[2]};
    h.material.colour = {colour[0], colour[1], colour[2]};
    h.material.roughness = roughness;
    Col c = roughDiffuse(h, V3{ld[0], ld[1], ld[2]}, Ray{{0, 0, 0}, {view_d[0], view_d[1], view_d[2]}});
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
// ImagePlane.create + rayThroughPixel (Appendix D 8)
void ftbo_primary_ray(const ftb_camera* cam, int width, int height, int px, int py, double jx, double jy, double* od)
{
    ImagePlane ip = createImagePlane(*cam, width, height);
    Ray r = rayThroughPixel(ip, px, py, jx, jy);
    od[0] = r.o.x; od[1] = r.o.y; od[2] = r.o.z; od[3] = r.d.x; od[4] = r.d.y; od[5] = r.d.z;
}
}
