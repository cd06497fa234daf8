// bvh_lockstep — how much of a warp's time the mesh traversal loop wastes on lock-step execution, on the CPU.
// Groups the scene camera's primary rays into warps of 32 (8x4 pixel blocks) and replays intersectMesh (render.cuh) for
// the 32 lanes together, charging every warp-level step its instruction cost whether 1 or 32 lanes take part:
//   A  the kernel's loop: all lanes descend inner nodes until each has reached a leaf (or run dry), then all test their
//      leaf's triangles, then pop ("while-while");
//   B  one loop in which a lane does either one node step or one triangle test per iteration ("if-if");
//   D  loop A with one postponed leaf per lane ("speculative traversal"): a lane that reaches a leaf parks it and keeps
//      descending until it reaches a second one; the triangle phase then serves both;
//   C  loop A fed from a pool: POOL consecutive warps' worth of rays (default 4 x 32) are traversed by one warp whose lanes
//      fetch the next pooled ray the moment their own ends (what a per-warp ray pool in shared memory would do).
// Reports useful lane-steps / (32 x warp-steps) and the warp instruction estimate of each (C_node, C_tri from the SASS).
// Analysis tool only.  Build like tools/bvh_quality.cpp; usage: bvh_lockstep scene.txt asset_dir [W H [spp]]
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

static const int kEmpty = 0x7fffffff;
struct Lane {
    V o, d, inv;
    double bt;
    int link, sp, tri_i, tri_n, tri_first;
    int stack[128]; double stackT[128];
    bool done;
};

struct Sim {
    const Lowered& L; const ftb_scene_desc* d;
    long nodeSteps = 0, triSteps = 0;  // lane-level useful steps
    void nodeStep(Lane& l)
    {
        ++nodeSteps;
        const BvhNode& nd = L.bvh_nodes[l.link];
        double tl = boxEntry(nd.dlo[0], nd.dhi[0], l.o, l.inv, l.bt), tr = boxEntry(nd.dlo[1], nd.dhi[1], l.o, l.inv, l.bt);
        bool hl = tl < INFINITY, hr = tr < INFINITY;
        if (hl && hr) {
            bool lf = tl <= tr;
            l.stack[l.sp] = lf ? nd.child[1] : nd.child[0]; l.stackT[l.sp] = lf ? tr : tl; ++l.sp;
            l.link = lf ? nd.child[0] : nd.child[1];
        } else if (hl || hr) l.link = hl ? nd.child[0] : nd.child[1];
        else l.link = kEmpty;
    }
    void enterLeaf(Lane& l) { int code = ~l.link; l.tri_first = code >> 3; l.tri_n = code & 7; l.tri_i = 0; }
    void triStep(Lane& l)
    {
        ++triSteps;
        double t;
        if (triT(d->triangles + 9 * (size_t)L.bvh_tri[l.tri_first + l.tri_i], l.o, l.d, t) && t < l.bt) l.bt = t;
        ++l.tri_i;
    }
    void pop(Lane& l)
    {
        l.link = kEmpty;
        while (l.sp > 0) { --l.sp; if (l.stackT[l.sp] <= l.bt) { l.link = l.stack[l.sp]; break; } }
        if (l.link == kEmpty) l.done = true;
    }
};

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
    int W = argc > 4 ? std::atoi(argv[3]) : 1920, H = argc > 4 ? std::atoi(argv[4]) : 1080;
    // spp > 1: lanes take SAMPLES like the kernel does (32 lanes = 32 / spp neighbouring pixels x spp jittered samples)
    const int spp = argc > 5 ? std::atoi(argv[5]) : 1;
    const int bw = spp >= 32 ? 1 : (spp >= 16 ? 2 : (spp >= 8 ? 2 : (spp >= 4 ? 4 : 8))), bh = std::max(1, 32 / (spp * bw));
    unsigned rngState = 12345u;
    auto rnd = [&]() { rngState = rngState * 1664525u + 1013904223u; return (rngState >> 8) * (1.0 / 16777216.0) - 0.5; };
    const double Cn = 45, Ct = 40, Cp = 8;  // warp instructions per node step / triangle test / pop (render.cuh SASS, 0x004 variant)
    V o = {cam->o[0], cam->o[1], cam->o[2]};
    V k = norm(sub({cam->look_at[0], cam->look_at[1], cam->look_at[2]}, o));
    V i = norm(cross({cam->up[0], cam->up[1], cam->up[2]}, k));
    V j = cross(k, i);
    double height = std::tan(cam->fov_y_rad / 2) * 2, width = height * cam->aspect_ratio;
    double ph = height / (W - 1), pw = width / (H - 1);
    for (size_t li = 0; li < L.leaves.size(); ++li) {
        const Leaf& lf = L.leaves[li];
        if (lf.kind != LEAF_MESH) continue;
        int root = L.mesh_root[lf.payload];
        if (root < 0) continue;
        const double* m = lf.w2m;
        double costA = 0, costB = 0, costC = 0, costD = 0, ideal = 0;
        std::vector<Lane> lanesD(32);
        Sim simD{L, d};
        const int POOL = 4;
        std::vector<Lane> pool; pool.reserve(32 * POOL);
        Sim simC{L, d};
        long warps = 0;
        std::vector<Lane> lanes(32), lanesB(32);
        Sim simA{L, d}, simB{L, d};
        for (int by = 0; by < H; by += (spp > 1 ? bh : 4))
            for (int bx = 0; bx < W; bx += (spp > 1 ? bw : 8)) {
                int n = 0;
                for (int y = by; y < std::min(H, by + (spp > 1 ? bh : 4)); ++y)
                    for (int x = bx; x < std::min(W, bx + (spp > 1 ? bw : 8)); ++x)
                      for (int sIdx = 0; sIdx < spp && n < 32; ++sIdx) {
                        const double ox = spp > 1 ? rnd() : 0.0, oy = spp > 1 ? rnd() : 0.0;
                        double jx = -width / 2 + pw / 2 + (x + ox) * pw, jy = height / 2 - ph / 2 - (y + oy) * ph;
                        V dw = {k.x + jx * i.x + jy * j.x, k.y + jx * i.y + jy * j.y, k.z + jx * i.z + jy * j.z};
                        Lane& l = lanes[n++];
                        l.o = {m[0] * o.x + m[1] * o.y + m[2] * o.z + m[3], m[4] * o.x + m[5] * o.y + m[6] * o.z + m[7], m[8] * o.x + m[9] * o.y + m[10] * o.z + m[11]};
                        l.d = {m[0] * dw.x + m[1] * dw.y + m[2] * dw.z, m[4] * dw.x + m[5] * dw.y + m[6] * dw.z, m[8] * dw.x + m[9] * dw.y + m[10] * dw.z};
                        l.inv = {1 / l.d.x, 1 / l.d.y, 1 / l.d.z};
                        l.bt = INFINITY; l.link = root; l.sp = 0; l.tri_i = l.tri_n = 0; l.done = false;
                    }
                for (int q = 0; q < n; ++q) lanesB[q] = lanes[q];
                for (int q = 0; q < n; ++q) pool.push_back(lanes[q]);
                for (int q = 0; q < n; ++q) lanesD[q] = lanes[q];
                {   // ---- D: while-while with a postponed leaf ----
                    std::vector<int> parked(n, kEmpty);
                    for (;;) {
                        bool anyAlive = false;
                        for (int q = 0; q < n; ++q) anyAlive |= !lanesD[q].done || parked[q] != kEmpty;
                        if (!anyAlive) break;
                        for (;;) {  // descend; a lane holding a parked leaf that reaches another leaf stops here
                            bool any = false;
                            for (int q = 0; q < n; ++q) {
                                Lane& l = lanesD[q];
                                if (l.done) continue;
                                if (l.link >= 0 && l.link != kEmpty) { simD.nodeStep(l); any = true; }
                                if (l.link < 0 && parked[q] == kEmpty) { parked[q] = l.link; simD.pop(l); }       // park it, carry on
                                else if (l.link == kEmpty) simD.pop(l);
                            }
                            if (!any) break;
                            costD += Cn + 4;
                        }
                        // triangle phase: the parked leaf, then the one the lane stopped at (if any)
                        for (int round = 0; round < 2; ++round) {
                            int mx = 0;
                            for (int q = 0; q < n; ++q) {
                                Lane& l = lanesD[q];
                                int leaf = round == 0 ? parked[q] : ((!l.done && l.link < 0) ? l.link : kEmpty);
                                if (leaf == kEmpty) { l.tri_n = 0; l.tri_i = 0; continue; }
                                int code = ~leaf; l.tri_first = code >> 3; l.tri_n = code & 7; l.tri_i = 0;
                                mx = std::max(mx, l.tri_n);
                            }
                            for (int s2 = 0; s2 < mx; ++s2) {
                                for (int q = 0; q < n; ++q) if (lanesD[q].tri_i < lanesD[q].tri_n) simD.triStep(lanesD[q]);
                                costD += Ct;
                            }
                            if (round == 0) for (int q = 0; q < n; ++q) parked[q] = kEmpty;
                            else for (int q = 0; q < n; ++q) if (!lanesD[q].done && lanesD[q].link < 0) simD.pop(lanesD[q]);
                        }
                        // entries popped while a closer hit was still parked may be stale: re-check the top against the new best t
                        for (int q = 0; q < n; ++q) {
                            Lane& l = lanesD[q];
                            if (!l.done && l.link >= 0 && l.link != kEmpty) continue;
                            if (!l.done && l.link == kEmpty) simD.pop(l);
                        }
                        costD += Cp;
                    }
                }
                ++warps;
                if (pool.size() >= (size_t)32 * POOL) {
                    // ---- C: while-while over a pool with dynamic fetch ----
                    size_t next = 0;
                    std::vector<Lane> cur(32);
                    int nl = 0;
                    for (; nl < 32 && next < pool.size(); ++nl) cur[nl] = pool[next++];
                    for (;;) {
                        bool anyAlive = false;
                        for (int q = 0; q < nl; ++q) {
                            if (cur[q].done && next < pool.size()) cur[q] = pool[next++];  // fetch
                            anyAlive |= !cur[q].done;
                        }
                        if (!anyAlive) break;
                        for (;;) {
                            bool any = false;
                            for (int q = 0; q < nl; ++q) if (!cur[q].done && cur[q].link >= 0 && cur[q].link != kEmpty) { simC.nodeStep(cur[q]); any = true; }
                            if (!any) break;
                            costC += Cn;
                        }
                        int mx = 0;
                        for (int q = 0; q < nl; ++q) if (!cur[q].done && cur[q].link < 0) { simC.enterLeaf(cur[q]); mx = std::max(mx, cur[q].tri_n); }
                        for (int s2 = 0; s2 < mx; ++s2) {
                            for (int q = 0; q < nl; ++q) if (!cur[q].done && cur[q].link < 0 && cur[q].tri_i < cur[q].tri_n) simC.triStep(cur[q]);
                            costC += Ct;
                        }
                        for (int q = 0; q < nl; ++q) if (!cur[q].done) simC.pop(cur[q]);
                        costC += Cp + 6;  // + the fetch bookkeeping
                    }
                    pool.clear();
                }
                long n0 = simA.nodeSteps, t0 = simA.triSteps;
                // ---- A: while-while (the kernel) ----
                for (;;) {
                    bool anyAlive = false;
                    for (int q = 0; q < n; ++q) anyAlive |= !lanes[q].done;
                    if (!anyAlive) break;
                    for (;;) {  // descend
                        bool any = false;
                        for (int q = 0; q < n; ++q) if (!lanes[q].done && lanes[q].link >= 0 && lanes[q].link != kEmpty) { simA.nodeStep(lanes[q]); any = true; }
                        if (!any) break;
                        costA += Cn;
                    }
                    int mx = 0;
                    for (int q = 0; q < n; ++q) if (!lanes[q].done && lanes[q].link < 0) { simA.enterLeaf(lanes[q]); mx = std::max(mx, lanes[q].tri_n); }
                    for (int s = 0; s < mx; ++s) {
                        for (int q = 0; q < n; ++q) if (!lanes[q].done && lanes[q].link < 0 && lanes[q].tri_i < lanes[q].tri_n) simA.triStep(lanes[q]);
                        costA += Ct;
                    }
                    for (int q = 0; q < n; ++q) if (!lanes[q].done) simA.pop(lanes[q]);
                    costA += Cp;
                }
                ideal += ((simA.nodeSteps - n0) * Cn + (simA.triSteps - t0) * Ct) / 32.0;
                // ---- B: if-if ----
                for (;;) {
                    bool anyNode = false, anyTri = false, anyAlive = false;
                    for (int q = 0; q < n; ++q) {
                        Lane& l = lanesB[q];
                        if (l.done) continue;
                        anyAlive = true;
                        if (l.link >= 0 && l.link != kEmpty) { simB.nodeStep(l); anyNode = true; if (l.link < 0) simB.enterLeaf(l); else if (l.link == kEmpty) simB.pop(l), (void)0; }
                        else if (l.link < 0 && l.tri_i < l.tri_n) { simB.triStep(l); anyTri = true; if (l.tri_i >= l.tri_n) { simB.pop(l); if (!l.done && l.link < 0) simB.enterLeaf(l); } }
                        else { simB.pop(l); if (!l.done && l.link < 0) simB.enterLeaf(l); }
                    }
                    if (!anyAlive) break;
                    costB += (anyNode ? Cn : 0) + (anyTri ? Ct : 0) + Cp;
                }
            }
        std::printf("mesh %d, %dx%d x %d spp primary rays in %ld warps: lane steps %ld nodes + %ld triangles\n", lf.payload, W, H, spp, warps, simA.nodeSteps, simA.triSteps);
        std::printf("  ideal (perfectly packed)      %8.1f M warp instructions\n", ideal / 1e6);
        std::printf("  A while-while (the kernel)    %8.1f M  = %.1f%% lock-step efficiency\n", costA / 1e6, 100 * ideal / costA);
        std::printf("  B if-if                       %8.1f M  = %.1f%%\n", costB / 1e6, 100 * ideal / costB);
        std::printf("  D loop A, one postponed leaf  %8.1f M  = %.1f%% (lane steps %ld nodes + %ld triangles)\n", costD / 1e6, 100 * ideal / costD, simD.nodeSteps, simD.triSteps);
        std::printf("  C loop A over a pool of %d rays %8.1f M  = %.1f%% (rays still pooled at the end of the image are not counted)\n", 32 * POOL, costC / 1e6, 100 * ideal / costC);
    }
    ftbf_destroy(fs);
    return 0;
}
