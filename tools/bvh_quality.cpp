// bvh_quality — CPU-side yardstick for the mesh index that ftb_scene_create builds (csrc/cuda/lower.cpp, BvhBuilder).
// Parses a scene with the host front end, lowers it exactly as the library does, and reports for every mesh:
//   nodes, leaves, depth, the SAH cost of the tree, and — for the scene camera's primary rays at WxH pixel centres —
//   nodes visited and triangles tested per ray with the kernel's traversal (front to back, culled against the best t).
// Any valid tree renders the same picture (ties are broken by the reference's enumeration rank, not by the tree), so the
// builder can be tuned against these numbers without a GPU.  Analysis tool only: nothing here is linked into the product.
//   g++ -O2 -std=c++17 -o /tmp/bvh_quality tools/bvh_quality.cpp functracer_b200/csrc/cuda/lower.cpp \
//       -Lfunctracer_b200 -lftb_frontend -Wl,-rpath,$PWD/functracer_b200
//   /tmp/bvh_quality scene.txt asset_dir [W H]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../functracer_b200/csrc/cuda/lower.h"
#include "../functracer_b200/csrc/frontend/ftb_frontend.h"

using namespace ftb;

struct V { double x, y, z; };
static V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { double l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }

static double area(const double* lo, const double* hi)
{
    double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2 * (dx * dy + dy * dz + dz * dx);
}

struct Stat { long nodes = 0, leaves = 0, tris = 0; int depth = 0; double sah = 0; };

static void walk(const Lowered& L, int link, const double* lo, const double* hi, double rootArea, int depth, Stat& s)
{
    s.depth = std::max(s.depth, depth);
    if (link < 0) {
        int n = (~link) & 7;
        ++s.leaves; s.tris += n;
        s.sah += area(lo, hi) / rootArea * n;  // C_isect = 1 per triangle
        return;
    }
    const BvhNode& nd = L.bvh_nodes[link];
    ++s.nodes;
    s.sah += area(lo, hi) / rootArea * 1.0;  // C_trav = 1 per node (both child boxes are tested in one visit)
    for (int c = 0; c < 2; ++c) walk(L, nd.child[c], nd.dlo[c], nd.dhi[c], rootArea, depth + 1, s);
}

static double boxEntry(const double* lo, const double* hi, V o, V inv, double tmax)
{
    double x0 = (lo[0] - o.x) * inv.x, x1 = (hi[0] - o.x) * inv.x;
    double y0 = (lo[1] - o.y) * inv.y, y1 = (hi[1] - o.y) * inv.y;
    double z0 = (lo[2] - o.z) * inv.z, z1 = (hi[2] - o.z) * inv.z;
    double tn = std::fmax(std::fmax(std::fmin(x0, x1), std::fmin(y0, y1)), std::fmax(std::fmin(z0, z1), 0.0));
    double tf = std::fmin(std::fmin(std::fmax(x0, x1), std::fmax(y0, y1)), std::fmin(std::fmax(z0, z1), tmax));
    return tn <= tf ? tn : INFINITY;
}

static bool triT(const double* t9, V o, V d, double& t)
{
    V v0 = {t9[0], t9[1], t9[2]}, e1 = sub({t9[3], t9[4], t9[5]}, v0), e2 = sub({t9[6], t9[7], t9[8]}, v0);
    V h = cross(d, e2);
    double a = dot(e1, h);
    if (a > -1e-7 && a < 1e-7) return false;
    double f = 1 / a;
    V s = sub(o, v0);
    double u = f * dot(s, h);
    if (u < 0 || u > 1) return false;
    V q = cross(s, e1);
    double v = f * dot(d, q);
    if (v < 0 || u + v > 1) return false;
    t = f * dot(e2, q);
    return t > 1e-7;
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s scene.txt asset_dir [W H]\n", argv[0]); return 2; }
    std::ifstream in(argv[1]);
    std::stringstream ss; ss << in.rdbuf();
    ftbf_scene* fs = nullptr;
    if (ftbf_parse(ss.str().c_str(), argv[2], &fs) != 0) { std::fprintf(stderr, "parse: %s\n", ftbf_last_error()); return 1; }
    const ftb_scene_desc* d = ftbf_desc(fs);
    const ftb_camera* cam = ftbf_camera(fs);
    Lowered L; std::string err;
    if (lower_scene(*d, L, err) != 0) { std::fprintf(stderr, "lower: %s\n", err.c_str()); return 1; }
    int W = argc > 4 ? std::atoi(argv[3]) : 480, H = argc > 4 ? std::atoi(argv[4]) : 270;
    std::printf("triangles %d, bvh nodes %zu, slots %zu\n", d->n_triangles, L.bvh_nodes.size(), L.bvh_tri.size());
    // camera frame (api.cu fillCamera, Image.fs:67-89)
    V o = {cam->o[0], cam->o[1], cam->o[2]};
    V k = norm(sub({cam->look_at[0], cam->look_at[1], cam->look_at[2]}, o));
    V i = norm(cross({cam->up[0], cam->up[1], cam->up[2]}, k));
    V j = cross(k, i);
    double height = std::tan(cam->fov_y_rad / 2) * 2, width = height * cam->aspect_ratio;
    double ph = height / (W - 1), pw = width / (H - 1);  // the reference's swapped axes
    for (size_t li = 0; li < L.leaves.size(); ++li) {
        const Leaf& lf = L.leaves[li];
        if (lf.kind != LEAF_MESH) continue;
        int root = L.mesh_root[lf.payload];
        if (root < 0) { std::printf("mesh %d: a single leaf\n", lf.payload); continue; }
        const BvhNode& rn = L.bvh_nodes[root];
        double lo[3], hi[3];
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(rn.dlo[0][a], rn.dlo[1][a]); hi[a] = std::max(rn.dhi[0][a], rn.dhi[1][a]); }
        Stat s;
        walk(L, root, lo, hi, area(lo, hi), 0, s);
        long visits = 0, tests = 0, rays = 0, hits = 0, maxVisits = 0;
        std::vector<int> steps((size_t)W * H, 0);  // node visits + triangle tests of every ray, for the lock-step estimate below
        std::vector<int> stack(256); std::vector<double> stackT(256);
        const double* m = lf.w2m;
        for (int py = 0; py < H; ++py)
            for (int px = 0; px < W; ++px) {
                double jx = -width / 2 + pw / 2 + px * pw, jy = height / 2 - ph / 2 - py * ph;
                V dw = {k.x + jx * i.x + jy * j.x, k.y + jx * i.y + jy * j.y, k.z + jx * i.z + jy * j.z};
                V ro = {m[0] * o.x + m[1] * o.y + m[2] * o.z + m[3], m[4] * o.x + m[5] * o.y + m[6] * o.z + m[7], m[8] * o.x + m[9] * o.y + m[10] * o.z + m[11]};
                V rd = {m[0] * dw.x + m[1] * dw.y + m[2] * dw.z, m[4] * dw.x + m[5] * dw.y + m[6] * dw.z, m[8] * dw.x + m[9] * dw.y + m[10] * dw.z};
                V inv = {1 / rd.x, 1 / rd.y, 1 / rd.z};
                double bt = INFINITY; bool hit = false;
                int sp = 0, link = root; long v0 = visits, t0 = tests;
                for (;;) {
                    while (link >= 0) {
                        ++visits;
                        const BvhNode& nd = L.bvh_nodes[link];
                        double tl = boxEntry(nd.dlo[0], nd.dhi[0], ro, inv, bt), tr = boxEntry(nd.dlo[1], nd.dhi[1], ro, inv, bt);
                        bool hl = tl < INFINITY, hr = tr < INFINITY;
                        if (hl && hr) {
                            bool lf1 = tl <= tr;
                            stack[sp] = lf1 ? nd.child[1] : nd.child[0]; stackT[sp] = lf1 ? tr : tl; ++sp;
                            link = lf1 ? nd.child[0] : nd.child[1];
                        } else if (hl || hr) link = hl ? nd.child[0] : nd.child[1];
                        else { link = 0x7fffffff; break; }
                    }
                    if (link < 0) {
                        int code = ~link, first = code >> 3, count = code & 7;
                        for (int q = 0; q < count; ++q) {
                            ++tests;
                            double t;
                            if (triT(d->triangles + 9 * (size_t)L.bvh_tri[first + q], ro, rd, t) && t < bt) { bt = t; hit = true; }
                        }
                    }
                    link = 0x7fffffff;
                    while (sp > 0) { --sp; if (stackT[sp] <= bt) { link = stack[sp]; break; } }
                    if (link == 0x7fffffff) break;
                }
                ++rays; hits += hit; maxVisits = std::max(maxVisits, visits - v0);
                steps[(size_t)py * W + px] = (int)(visits - v0) + (int)(tests - t0);
            }
        // a warp traces the 32 pixels of an 8x4 block in lock step: it runs as long as its longest ray
        long sum = 0, lock = 0;
        for (int by = 0; by < H; by += 4)
            for (int bx = 0; bx < W; bx += 8) {
                int mx = 0;
                for (int y = by; y < std::min(H, by + 4); ++y)
                    for (int x = bx; x < std::min(W, bx + 8); ++x) { sum += steps[(size_t)y * W + x]; mx = std::max(mx, steps[(size_t)y * W + x]); }
                lock += 32L * mx;
            }
        std::printf("mesh %d: lock-step efficiency of 8x4 pixel blocks %.1f%% (sum of steps / 32 x longest ray of each block)\n", lf.payload, 100.0 * sum / std::max(1L, lock));
        std::printf("mesh %d: nodes %ld leaves %ld tris/leaf %.2f depth %d SAH %.2f | %dx%d primary rays: hit %.1f%%, nodes/ray %.2f (max %ld), tris/ray %.2f\n",
                    lf.payload, s.nodes, s.leaves, (double)s.tris / std::max(1L, s.leaves), s.depth, s.sah, W, H, 100.0 * hits / rays,
                    (double)visits / rays, maxVisits, (double)tests / rays);
    }
    ftbf_destroy(fs);
    return 0;
}
