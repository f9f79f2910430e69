// wide_stats — development tool (CPU only): the 8-wide quantised tree of csrc/ipt_wide.h walked the way k_extend_wide
// walks it (fp32 slab arithmetic on the quantised planes, nearest hit child first, the others pushed with their entry
// distance and culled when popped), against the linear scan of Renderer.cu:227-243.
//   g++ -O2 -std=c++17 -Iinclude tools/wide_stats.cpp -Limproved-path-tracer_b200 -lipt_b200 -Wl,-rpath,$PWD/improved-path-tracer_b200 -o /tmp/wide_stats
//   /tmp/wide_stats scene.json [leaf2=4] [leaf_max=8] [rays=20000] [check] [sorted]
// Prints node steps, leaf steps, pops and primitive tests per ray and the deepest stack seen; with `check` every ray is
// also tested against every primitive and the nearest distance must agree exactly (exit 1 otherwise).
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
    if (argc < 2) { std::fprintf(stderr, "usage: wide_stats scene.json [leaf2] [leaf_max] [rays] [check] [sorted]\n"); return 2; }
    const uint32_t leaf2 = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 4, leaf_max = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 8;
    const int n_rays = argc > 4 ? std::atoi(argv[4]) : 20000;
    bool check = false, sorted = false, octant = false, stat = false; int multi = 0;
    for (int i = 5; i < argc; i++) { check = check || std::string(argv[i]) == "check"; sorted = sorted || std::string(argv[i]) == "sorted"; octant = octant || std::string(argv[i]) == "octant"; stat = stat || std::string(argv[i]) == "static"; if (std::string(argv[i]).rfind("multi", 0) == 0) multi = std::atoi(argv[i] + 5); }
    char msg[256];
    ipt_host_scene* hs = ipt_host_load_scene(argv[1], msg, sizeof msg);
    if (!hs) { std::fprintf(stderr, "%s\n", msg); return 1; }
    if (ipt_host_build_bvh(hs, leaf2, 0) < 0) { std::fprintf(stderr, "BVH build failed\n"); return 1; }
    const ipt_scene* s = ipt_host_scene_view(hs);
    ipt::WideTree wt;
    if (const char* e = ipt::wide_collapse(s->bvh_nodes, s->n_bvh_nodes, s->n_bvh_slots, leaf_max, wt)) { std::fprintf(stderr, "%s\n", e); return 1; }
    std::printf("2-wide: %u nodes (leaf %u)  ->  8-wide: %zu nodes, %.2f children per node, depth %u, stack need <= %u\n", s->n_bvh_nodes, leaf2,
                wt.nodes.size(), wt.sum_children / wt.nodes.size(), wt.depth, wt.stack_need);
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> N(0.0, 1.0);
    const double lo[3] = {30, -480, 30}, hi[3] = {1250, 680, 690};
    unsigned long long steps = 0, leaves = 0, prims = 0, pops = 0, culled = 0, wrong = 0, hits = 0, pushes = 0;
    uint32_t max_sp = 0;
    struct Entry { int32_t link; float tn; };
    std::vector<Entry> stack(1024);
    std::vector<Entry> ms[8];
    for (auto& v : ms) v.reserve(256);
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
        double best = 1e20;
        uint32_t sp = 0;
        int32_t cur = 0;
        for (;;) {
            if (cur >= 0) {
                steps++;
                const ipt::WideNode& nd = wt.nodes[cur];
                float A[3], B[3];
                for (int k = 0; k < 3; k++) { A[k] = nd.scale[k] * bi[k]; B[k] = std::fmaf(-8388608.f, A[k], std::fmaf(nd.origin[k], bi[k], -oi[k])); }
                float tn[8]; bool h[8]; int nh = 0, first = -1;
                for (int j = 0; j < 8; j++) {
                    float n = 0.f, f = (float)best;
                    for (int k = 0; k < 3; k++) {
                        const float qn = 8388608.f + (float)nd.q[j][2 * k + sgn[k]], qf = 8388608.f + (float)nd.q[j][2 * k + 1 - sgn[k]];
                        n = std::fmax(n, std::fmaf(qn, A[k], B[k])); f = std::fmin(f, std::fmaf(qf, A[k], B[k]));
                    }
                    tn[j] = n; h[j] = n <= f * 1.0000004f;
                    if (h[j]) { nh++; if (first < 0 || n < tn[first]) first = j; }
                }
                if (!nh) cur = ipt::WIDE_EMPTY;
                else {
                    int order[8], m = 0;
                    for (int j = 0; j < 8; j++) if (h[j] && j != first) order[m++] = j;
                    if (sorted) std::sort(order, order + m, [&](int a, int b) { return tn[a] > tn[b]; });   // farthest first: nearest on top
                    if (octant) {   // by the child centre along the ray's sign vector (what a per-octant slot order can approximate)
                        float key[8];
                        for (int j = 0; j < 8; j++) { key[j] = 0; for (int k = 0; k < 3; k++) key[j] += (sgn[k] ? -1.f : 1.f) * nd.scale[k] * (float)(nd.q[j][2 * k] + nd.q[j][2 * k + 1]); }
                        std::sort(order, order + m, [&](int a, int b) { return key[a] > key[b]; });
                    }
                    if (stat) {     // what the kernel does: increasing (slot ^ octant) from the top of the stack down
                        const int oct = sgn[0] | sgn[1] << 1 | sgn[2] << 2;
                        std::sort(order, order + m, [&](int a, int b) { return (a ^ oct) > (b ^ oct); });
                    }
                    if (multi) { for (int k = 0; k < m; k++) { ms[order[k] / (8 / multi)].push_back(Entry{nd.link[order[k]], tn[order[k]]}); uint32_t tot = 0, mx = 0; for (int q = 0; q < multi; q++) { tot += ms[q].size(); mx = std::max<uint32_t>(mx, ms[q].size()); } max_sp = std::max(max_sp, mx); } }
                    else
                    for (int k = 0; k < m; k++) stack[sp++] = Entry{nd.link[order[k]], tn[order[k]]};
                    pushes += m;
                    max_sp = std::max(max_sp, sp);
                    cur = nd.link[first];
                }
            } else {
                leaves++;
                const uint32_t code = (uint32_t)~cur, f = code >> 4, cnt = (code & 15u) + 1u;
                for (uint32_t k = 0; k < cnt; k++) { prims++; best = hit_prim(s, s->bvh_slot_prim[f + k], r, best); }
                cur = ipt::WIDE_EMPTY;
            }
            while (multi && cur == ipt::WIDE_EMPTY) {
                int bq = -1;
                for (int q = 0; q < multi; q++) if (!ms[q].empty() && (bq < 0 || ms[q].back().tn < ms[bq].back().tn)) bq = q;
                if (bq < 0) break;
                pops++;
                const Entry e = ms[bq].back(); ms[bq].pop_back();
                if ((double)e.tn <= best * 1.0000004) cur = e.link; else culled++;
            }
            while (!multi && cur == ipt::WIDE_EMPTY && sp) {
                pops++;
                const Entry e = stack[--sp];
                if ((double)e.tn <= best * 1.0000004) cur = e.link; else culled++;
            }
            if (cur == ipt::WIDE_EMPTY) break;
        }
        hits += best < 1e20;
        if (check) {
            double lin = 1e20;
            for (uint32_t k = 0; k < s->n_spheres; k++) lin = hit_prim(s, k, r, lin);
            for (uint32_t k = 0; k < s->n_rects; k++) lin = hit_prim(s, 0x80000000u | k, r, lin);
            wrong += lin != best;
        }
    }
    std::printf("per ray: %.2f node steps, %.2f leaf steps, %.2f primitive tests, %.2f pushes, %.2f pops (%.2f culled); deepest stack %u; %.1f %% hit\n",
                (double)steps / n_rays, (double)leaves / n_rays, (double)prims / n_rays, (double)pushes / n_rays, (double)pops / n_rays,
                (double)culled / n_rays, max_sp, 100.0 * hits / n_rays);
    if (check) std::printf("check: %llu of %d rays differ from the linear scan\n", wrong, n_rays);
    ipt_host_free_scene(hs);
    return wrong ? 1 : 0;
}
