// wide_stats — development and test tool (CPU only): the 8-wide quantised tree of csrc/ipt_wide.h walked the way
// k_extend_cw walks it (fp32 slab arithmetic on the quantised planes; one stack entry per node holding the hit inner
// children as a bit mask, visited in increasing slot ^ octant; the hit leaf children as a bit mask over the node's
// consecutive primitive slots; implicit child and primitive addressing), against the linear scan of Renderer.cu:227-243.
//   g++ -O2 -std=c++17 -Iinclude tools/wide_stats.cpp -Limproved-path-tracer_b200 -lipt_b200 -Wl,-rpath,$PWD/improved-path-tracer_b200 -o /tmp/wide_stats
//   /tmp/wide_stats scene.json [leaf2=4] [leaf_max=4] [rays=20000] [check] [greedy] [cnode=4]
// Prints node steps, leaf children entered and primitive tests per ray and the deepest stack seen; with `check` every ray
// is also tested against every primitive and the nearest distance must agree exactly (exit 1 otherwise).  The links of a
// node (WideNode::link) are cross-checked against the implicit addressing on every step.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
#include "ipt_host.h"
#include "../improved-path-tracer_b200/csrc/ipt_wide.h"

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

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: wide_stats scene.json [leaf2] [leaf_max] [rays] [check]\n"); return 2; }
    const uint32_t leaf2 = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 4, leaf_max = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 4;
    const int n_rays = argc > 4 ? std::atoi(argv[4]) : 20000;
    bool check = false, greedy = false;
    double c_node = 4.0;
    for (int i = 5; i < argc; i++) {
        check = check || std::string(argv[i]) == "check"; greedy = greedy || std::string(argv[i]) == "greedy";
        if (std::string(argv[i]).rfind("cnode=", 0) == 0) c_node = std::atof(argv[i] + 6);
    }
    char msg[256];
    ipt_host_scene* hs = ipt_host_load_scene(argv[1], msg, sizeof msg);
    if (!hs) { std::fprintf(stderr, "%s\n", msg); return 1; }
    if (ipt_host_build_bvh(hs, leaf2, 0) < 0) { std::fprintf(stderr, "BVH build failed\n"); return 1; }
    const ipt_scene* s = ipt_host_scene_view(hs);
    ipt::WideTree wt;
    if (const char* e = ipt::wide_collapse(s->bvh_nodes, s->n_bvh_nodes, s->n_bvh_slots, leaf_max, wt, greedy, c_node)) { std::fprintf(stderr, "%s\n", e); return 1; }
    std::printf("2-wide: %u nodes (leaf %u)  ->  8-wide: %zu nodes, %.2f children per node, depth %u\n", s->n_bvh_nodes, leaf2,
                wt.nodes.size(), wt.sum_children / wt.nodes.size(), wt.depth);
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> N(0.0, 1.0);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};   // ray origins: inside the box of the sphere centres
    for (uint32_t k = 0; k < s->n_spheres; k++) for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], s->sphere_cxyzr[4 * (size_t)k + a]); hi[a] = std::max(hi[a], s->sphere_cxyzr[4 * (size_t)k + a]); }
    if (!s->n_spheres) for (int a = 0; a < 3; a++) { lo[a] = -100; hi[a] = 100; }
    unsigned long long steps = 0, leaves = 0, prims = 0, wrong = 0, hits = 0, bad_links = 0;
    uint32_t max_sp = 0;
    struct Group { uint32_t base_imask, hits; };
    std::vector<Group> stack(256);
    for (int i = 0; i < n_rays; i++) {
        Ray r;
        double len = 0;
        for (int k = 0; k < 3; k++) { r.o[k] = lo[k] + U(rng) * (hi[k] - lo[k]); r.d[k] = N(rng); len += r.d[k] * r.d[k]; }
        len = std::sqrt(len);
        for (int k = 0; k < 3; k++) r.d[k] /= len;
        // device arithmetic: fp32, 1/d clamped away from infinity, planes as q * (scale / d) + (origin / d - o / d)
        float o[3], bi[3], oi[3];
        int sgn[3];
        for (int k = 0; k < 3; k++) {
            o[k] = (float)r.o[k];
            const float d = (float)r.d[k];
            bi[k] = 1.f / (std::fabs(d) > 1e-18f ? d : std::copysign(1e-18f, d));
            oi[k] = o[k] * bi[k];
            sgn[k] = std::signbit(d) ? 1 : 0;
        }
        const uint32_t oct = (uint32_t)(sgn[0] | sgn[1] << 1 | sgn[2] << 2), octx = oct ^ 7u;
        double best = 1e20;
        uint32_t sp = 0;
        Group g{1u << 24, 1u << octx};     // the root: slot 0 of a node whose only inner child is node 0
        for (;;) {
            if (g.hits == 0) {
                if (sp == 0) break;
                g = stack[--sp];
            }
            uint32_t qb = 31;
            while (!(g.hits >> qb & 1u)) qb--;
            const uint32_t slot = qb ^ octx;
            g.hits &= ~(1u << qb);
            const uint32_t node = (g.base_imask & 0xFFFFFFu) + (uint32_t)__builtin_popcount((g.base_imask >> 24) & ((1u << slot) - 1u));
            if (g.hits) { stack[sp++] = g; max_sp = std::max(max_sp, sp); }
            steps++;
            const ipt::WideNode& nd = wt.nodes[node];
            float A[3], B[3];
            for (int k = 0; k < 3; k++) { A[k] = nd.scale[k] * bi[k]; B[k] = std::fmaf(-8388608.f, A[k], std::fmaf(nd.origin[k], bi[k], -oi[k])); }
            const float tmax = (float)best * 1.0000004f;
            uint32_t hit8 = 0, pm = 0;
            for (int j = 0; j < 8; j++) {
                float n = 0.f, f = tmax;
                for (int k = 0; k < 3; k++) {
                    const float qn = 8388608.f + (float)nd.q[j][2 * k + sgn[k]], qf = 8388608.f + (float)nd.q[j][2 * k + 1 - sgn[k]];
                    n = std::fmax(n, std::fmaf(qn, A[k], B[k])); f = std::fmin(f, std::fmaf(qf, A[k], B[k]));
                }
                if (n <= f) { hit8 |= 1u << j; pm |= nd.pmask & (0xFu << (4 * j)); }
            }
            const uint32_t prim_base = (uint32_t)nd.q[0][6] | (uint32_t)nd.q[0][7] << 8 | (uint32_t)nd.q[1][6] << 16 | (uint32_t)nd.q[1][7] << 24;
            // the links say the same as the implicit addressing
            for (int j = 0; j < 8; j++) {
                const bool inner = (nd.base_imask >> (24 + j)) & 1u, leaf = (nd.pmask >> (4 * j)) & 0xFu;
                if (inner && nd.link[j] != (int32_t)((nd.base_imask & 0xFFFFFFu) + __builtin_popcount((nd.base_imask >> 24) & ((1u << j) - 1u)))) bad_links++;
                if (leaf) {
                    const uint32_t code = (uint32_t)~nd.link[j];
                    if ((code >> 4) != prim_base + (uint32_t)__builtin_popcount(nd.pmask & ((1u << (4 * j)) - 1u)) || (code & 15u) + 1u != (uint32_t)__builtin_popcount(nd.pmask & (0xFu << (4 * j)))) bad_links++;
                }
                if (!inner && !leaf && nd.link[j] != ipt::WIDE_EMPTY) bad_links++;
            }
            leaves += __builtin_popcount(hit8 & ~(nd.base_imask >> 24));
            while (pm) {
                const uint32_t bit = (uint32_t)__builtin_ctz(pm);
                pm &= pm - 1u;
                const uint32_t slot_new = prim_base + (uint32_t)__builtin_popcount(nd.pmask & ((1u << bit) - 1u));
                prims++;
                best = hit_prim(s, s->bvh_slot_prim[wt.perm[slot_new]], r, best);
            }
            // hit inner children, in the order the device's table gives: slot j at bit 7 - (j ^ oct)
            const uint32_t inner = hit8 & (nd.base_imask >> 24);
            uint32_t permuted = 0;
            for (uint32_t j = 0; j < 8; j++) permuted |= ((inner >> j) & 1u) << (7u - (j ^ oct));
            if (g.hits) { /* already pushed */ }
            g = Group{nd.base_imask, permuted};
        }
        hits += best < 1e20;
        if (check) {
            double lin = 1e20;
            for (uint32_t k = 0; k < s->n_spheres; k++) lin = hit_prim(s, k, r, lin);
            for (uint32_t k = 0; k < s->n_rects; k++) lin = hit_prim(s, 0x80000000u | k, r, lin);
            wrong += lin != best;
        }
    }
    std::printf("per ray: %.2f node steps, %.2f leaf children entered, %.2f primitive tests; deepest stack %u (tree depth %u); %.1f %% hit; %llu link mismatches\n",
                (double)steps / n_rays, (double)leaves / n_rays, (double)prims / n_rays, max_sp, wt.depth, 100.0 * hits / n_rays, bad_links);
    if (check) std::printf("check: %llu of %d rays differ from the linear scan\n", wrong, n_rays);
    ipt_host_free_scene(hs);
    return (wrong || bad_links) ? 1 : 0;
}
