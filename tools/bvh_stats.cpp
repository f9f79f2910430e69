// bvh_stats — development tool: how much traversal work does the host-built BVH ask of a ray?
//   g++ -O2 -std=c++17 -Iinclude tools/bvh_stats.cpp -Limproved-path-tracer_b200 -lipt_b200 -Wl,-rpath,$PWD/improved-path-tracer_b200 -o /tmp/bvh_stats
//   /tmp/bvh_stats scene.json [leaf_size=4] [rays=20000] [check]
// Random rays inside the scene's bounding box, nearest hit by ordered (near child first) traversal with the same
// pruning as k_extend_bvh (ipt_kernels.cuh); prints node visits, leaf visits and primitive tests per ray.  With
// `check` every ray is also tested against every primitive: the tree must report the same nearest distance (exit 1
// otherwise) - the builder's boxes are conservative and every primitive sits in exactly one leaf.  CPU only.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
#include "ipt_host.h"

struct Ray { double o[3], d[3]; };

static double hit_prim(const ipt_scene* s, uint32_t prim, const Ray& r, double best)
{
    if (!(prim & 0x80000000u)) {
        const double* sp = s->sphere_cxyzr + 4 * (size_t)prim;
        const double op[3] = {r.o[0] - sp[0], r.o[1] - sp[1], r.o[2] - sp[2]};
        const double b = op[0] * r.d[0] + op[1] * r.d[1] + op[2] * r.d[2];
        const double delta = b * b - (op[0] * op[0] + op[1] * op[1] + op[2] * op[2]) + sp[3] * sp[3];
        if (delta < 0) return best;
        const double sq = std::sqrt(delta), t1 = -b - sq, t2 = -b + sq;
        const double t = t1 > 1e-4 ? t1 : (t2 > 1e-4 ? t2 : 0.0);
        return (t != 0.0 && t < best) ? t : best;
    }
    const size_t j = prim & 0x7FFFFFFFu;
    const double *pl = s->rect_plane + 4 * j, *u = s->rect_u + 4 * j, *v = s->rect_v + 4 * j, *bd = s->rect_bounds + 4 * j;
    const double den = pl[0] * r.d[0] + pl[1] * r.d[1] + pl[2] * r.d[2];
    if (den == 0) return best;
    const double t = (pl[3] - (pl[0] * r.o[0] + pl[1] * r.o[1] + pl[2] * r.o[2])) / den;
    if (!(t > 1e-4) || !(t < best)) return best;
    const double P[3] = {r.o[0] + r.d[0] * t, r.o[1] + r.d[1] * t, r.o[2] + r.d[2] * t};
    const double su = std::fabs(u[0] * P[0] + u[1] * P[1] + u[2] * P[2] - u[3]), sv = std::fabs(v[0] * P[0] + v[1] * P[1] + v[2] * P[2] - v[3]);
    return (su >= bd[0] && su <= bd[1] && sv >= bd[2] && sv <= bd[3]) ? t : best;
}

static bool slab(const float* lo, const float* hi, const Ray& r, const double* inv, double best, double& tn)
{
    double n = 0, f = best;
    for (int k = 0; k < 3; k++) {
        const double a = (lo[k] - r.o[k]) * inv[k], b = (hi[k] - r.o[k]) * inv[k];
        n = std::fmax(n, std::fmin(a, b)); f = std::fmin(f, std::fmax(a, b));
    }
    tn = n;
    return n <= f * 1.0000004;
}

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: bvh_stats scene.json [leaf_size] [rays]\n"); return 2; }
    const uint32_t leaf = argc > 2 ? (uint32_t)std::atoi(argv[2]) : IPT_DEFAULT_LEAF_SIZE;
    const int n_rays = argc > 3 ? std::atoi(argv[3]) : 20000;
    const bool check = argc > 4 && std::string(argv[4]) == "check";
    char msg[256];
    ipt_host_scene* hs = ipt_host_load_scene(argv[1], msg, sizeof msg);
    if (!hs) { std::fprintf(stderr, "%s\n", msg); return 1; }
    if (ipt_host_build_bvh(hs, leaf, 0) < 0) { std::fprintf(stderr, "BVH build failed\n"); return 1; }
    const ipt_scene* s = ipt_host_scene_view(hs);
    if (!s->n_bvh_nodes) { std::fprintf(stderr, "no BVH built\n"); return 1; }
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> N(0.0, 1.0);
    const double lo[3] = {30, -480, 30}, hi[3] = {1250, 680, 690};   // inside the synthetic room (scripts/make_synthetic_scene.py)
    unsigned long long nodes = 0, leaves = 0, prims = 0, hits = 0, wrong = 0;
    std::vector<int> stack(256);
    for (int i = 0; i < n_rays; i++) {
        Ray r;
        double len = 0;
        for (int k = 0; k < 3; k++) { r.o[k] = lo[k] + U(rng) * (hi[k] - lo[k]); r.d[k] = N(rng); len += r.d[k] * r.d[k]; }
        len = std::sqrt(len);
        double inv[3];
        for (int k = 0; k < 3; k++) { r.d[k] /= len; inv[k] = 1.0 / r.d[k]; }
        double best = 1e20;
        int sp = 0, node = 0;
        for (;;) {
            if (node >= 0) {
                nodes++;
                const ipt_bvh_node& b = s->bvh_nodes[node];
                double n0, n1;
                const bool h0 = slab(b.lo0, b.hi0, r, inv, best, n0), h1 = slab(b.lo1, b.hi1, r, inv, best, n1);
                // leaves are encoded as -(first*32 + count) - 1 on this private stack
                auto enc = [&](int k) { return b.child[k] >= 0 ? b.child[k] : -(int)(((uint32_t)~b.child[k]) * 32u + b.count[k]) - 1; };
                if (h0 && h1) { const bool sw = n1 < n0; stack[sp++] = sw ? enc(0) : enc(1); node = sw ? enc(1) : enc(0); }
                else if (h0) node = enc(0);
                else if (h1) node = enc(1);
                else if (sp) node = stack[--sp];
                else break;
            } else {
                leaves++;
                const uint32_t code = (uint32_t)(-(node + 1)), first = code / 32u, cnt = code % 32u;
                for (uint32_t k = 0; k < cnt; k++) { prims++; best = hit_prim(s, s->bvh_slot_prim[first + k], r, best); }
                if (sp) node = stack[--sp]; else break;
            }
        }
        hits += best < 1e20;
        if (check) {
            double lin = 1e20;
            for (uint32_t k = 0; k < s->n_spheres; k++) lin = hit_prim(s, k, r, lin);
            for (uint32_t k = 0; k < s->n_rects; k++) lin = hit_prim(s, 0x80000000u | k, r, lin);
            wrong += lin != best;
        }
    }
    std::printf("leaf_size %u: %u nodes; per ray: %.1f node visits, %.1f leaf visits, %.1f primitive tests; %.1f %% of the rays hit\n", leaf,
                s->n_bvh_nodes, (double)nodes / n_rays, (double)leaves / n_rays, (double)prims / n_rays, 100.0 * hits / n_rays);
    if (check) std::printf("check: %llu of %d rays differ from the linear scan\n", wrong, n_rays);
    ipt_host_free_scene(hs);
    return wrong ? 1 : 0;
}
