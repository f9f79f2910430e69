// bvh.cpp — host BVH builder (binned SAH, 2-wide, 64-byte nodes emitted breadth-first).
// New work: the reference has no acceleration structure, its nearest hit is a linear scan over all objects
// (Renderer.cu:227-243).  The BVH must return the same nearest hit as that scan, so boxes are conservative:
// padded for the +-5e-5 edge tolerance of Plane.cu:87-100, for the MARGIN = 1e-4 near-surface roots of
// Sphere.cu:36-37 and for fp32 rounding of the bounds themselves.
//
// Non-unit directions.  Rays leaving a refractive sphere are not normalised (AObject.hpp:59) and Sphere::intersect is
// used unchanged on them, so its "hit" is not the geometric one.  Culling by geometric boxes is still exact:
//  * lengths never exceed 1: |T|^2 = 1 - eta^2 (1 - l^2) on entry and <= that on exit for an incoming length l <= 1;
//    reflection keeps the length, the diffuse rule resets it to 1;
//  * for |d| = l <= 1, with x = -(op.d^) > 0 and c = op.op - r^2 > 0 the reference's root is f(l x) and the geometric
//    entry (in units of d) is f(x)/l with f(y) = y - sqrt(y^2 - c) = c / (y + sqrt(y^2 - c)); then
//    l f(l x) = c / (x + sqrt(x^2 - c/l^2)) >= c / (x + sqrt(x^2 - c)) = f(x): the reported t is never in front of the
//    sphere's (hence the box's) entry point, and delta_ref >= 0 implies the line does meet the sphere;
//  * c <= 0 means the origin is inside the sphere, hence inside (or within the padding of) its box.
// So a box that the reported hit needs is never culled by `entry <= best t`.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <thread>

#include "host_scene.hpp"

namespace {

// Builder boxes are fp32 (the build is memory-bound: 40-byte instead of 80-byte primitives); fp64 bounds are rounded
// outwards when they enter a box, so boxes stay conservative.
struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const double* p)
    {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], std::nextafterf((float)p[k], -INFINITY));
            hi[k] = std::max(hi[k], std::nextafterf((float)p[k], INFINITY));
        }
    }
    void grow(const float* p) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    double area() const
    {
        const double x = (double)hi[0] - lo[0], y = (double)hi[1] - lo[1], z = (double)hi[2] - lo[2];
        return (x < 0 || y < 0 || z < 0) ? 0.0 : 2 * (x * y + y * z + z * x);
    }
};

struct Prim { Box box; float c[3]; uint32_t ref; };

struct BNode { Box box; int left = -1, right = -1; uint32_t first = 0, count = 0; };

struct Builder {
    std::vector<Prim>& prims;
    std::vector<BNode> nodes;
    uint32_t leaf_size;
    int par_depth = 0;   // levels below this node that may still fork a thread

    // Splits [first, first+count) in place; returns the split position, or first when the range becomes a leaf.
    // `level`: depth of the node being split (root = 1).  The traversal stacks hold 64 entries and ipt_ctx_set_scene refuses
    // trees deeper than 60 levels; SAH on a skewed distribution (a geometric progression of sizes) can peel one primitive
    // per level, so from level 40 on a range is halved by index (every level halves the count: at most 32 more levels).
    static constexpr int SAH_MAX_LEVEL = 40;
    uint32_t split(uint32_t first, uint32_t count, Box& box, int level)
    {
        Box cb;
        for (uint32_t i = first; i < first + count; i++) { box.grow(prims[i].box); cb.grow(prims[i].c); }
        if (count <= leaf_size) return first;
        if (level >= SAH_MAX_LEVEL) return first + count / 2;
        // binned SAH over the centroid bounds, 16 bins per axis
        constexpr int NB = 16;
        constexpr uint32_t SMALL = 16;
        int bestAxis = -1, bestSplit = -1;
        double bestCost = DBL_MAX;
        if (count <= SMALL) {
            // small ranges (the bulk of the nodes, and the ones whose boxes a ray grazes most often): exact SAH - sort by
            // centroid along each axis and evaluate every split position (48 bins per node would cost more than this)
            static const bool median_only = std::getenv("IPT_BVH_MEDIAN_SMALL") != nullptr;   // A/B knob: the old rule
            if (median_only) {
                int ax = 0;
                for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[ax] - cb.lo[ax]) ax = k;
                const uint32_t mid = first + count / 2;
                std::nth_element(prims.begin() + first, prims.begin() + mid, prims.begin() + first + count,
                                 [ax](const Prim& a, const Prim& b) { return a.c[ax] < b.c[ax]; });
                return mid;
            }
            uint32_t order[3][SMALL];
            double rightArea[SMALL];
            int bestAx = -1; uint32_t bestK = 0;
            for (int ax = 0; ax < 3; ax++) {
                uint32_t* o = order[ax];
                for (uint32_t i = 0; i < count; i++) o[i] = first + i;
                std::sort(o, o + count, [&](uint32_t a, uint32_t b) { return prims[a].c[ax] < prims[b].c[ax] || (prims[a].c[ax] == prims[b].c[ax] && a < b); });
                Box acc;
                for (uint32_t i = count - 1; i > 0; i--) { acc.grow(prims[o[i]].box); rightArea[i] = acc.area(); }
                acc = Box();
                for (uint32_t k = 1; k < count; k++) {          // left = o[0..k), right = o[k..count)
                    acc.grow(prims[o[k - 1]].box);
                    const double cost = acc.area() * k + rightArea[k] * (count - k);
                    if (cost < bestCost) { bestCost = cost; bestAx = ax; bestK = k; }
                }
            }
            if (bestAx < 0) return first + count / 2;   // no finite cost (boxes of +-FLT_MAX extent): split by index
            Prim tmp[SMALL];
            for (uint32_t i = 0; i < count; i++) tmp[i] = prims[order[bestAx][i]];
            for (uint32_t i = 0; i < count; i++) prims[first + i] = tmp[i];
            return first + bestK;
        }
        // one pass over the primitives fills the bins of all three axes (chunks of a large range on several threads)
        struct Bins { Box bb[3][NB]; uint32_t bc[3][NB]; Bins() { std::memset(bc, 0, sizeof(bc)); } };
        double kk[3];
        for (int ax = 0; ax < 3; ax++) { const double ext = (double)cb.hi[ax] - cb.lo[ax]; kk[ax] = ext > 0 ? NB * (1 - 1e-9) / ext : 0.0; }
        auto fill = [&](Bins& B, uint32_t a, uint32_t e) {
            for (uint32_t i = a; i < e; i++)
                for (int ax = 0; ax < 3; ax++) {
                    const int b = std::min(NB - 1, std::max(0, (int)(((double)prims[i].c[ax] - cb.lo[ax]) * kk[ax])));
                    B.bb[ax][b].grow(prims[i].box); B.bc[ax][b]++;
                }
        };
        Bins bins;
        const int nthr = (par_depth > 0 && count > 262144) ? (1 << par_depth) : 1;
        if (nthr == 1) fill(bins, first, first + count);
        else {
            std::vector<Bins> part(nthr);
            std::vector<std::thread> th;
            for (int t = 0; t < nthr; t++)
                th.emplace_back([&, t] { fill(part[t], first + (uint32_t)((uint64_t)count * t / nthr), first + (uint32_t)((uint64_t)count * (t + 1) / nthr)); });
            for (auto& t : th) t.join();
            for (const Bins& P : part)
                for (int ax = 0; ax < 3; ax++)
                    for (int b = 0; b < NB; b++) { if (P.bc[ax][b]) bins.bb[ax][b].grow(P.bb[ax][b]); bins.bc[ax][b] += P.bc[ax][b]; }
        }
        for (int ax = 0; ax < 3; ax++) {
            if (!(kk[ax] > 0)) continue;
            const Box* bb = bins.bb[ax]; const uint32_t* bc = bins.bc[ax];
            double rightArea[NB]; uint32_t rightCnt[NB];
            Box acc; uint32_t cnt = 0;
            for (int b = NB - 1; b > 0; b--) { acc.grow(bb[b]); cnt += bc[b]; rightArea[b] = acc.area(); rightCnt[b] = cnt; }
            acc = Box(); cnt = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bb[b]); cnt += bc[b];
                if (cnt == 0 || rightCnt[b + 1] == 0) continue;
                const double cost = acc.area() * cnt + rightArea[b + 1] * rightCnt[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestSplit = b; }
            }
        }
        uint32_t mid;
        if (bestAxis < 0) {
            mid = first + count / 2;   // all centroids coincide: split by index
        } else {
            const double k = kk[bestAxis], lo = cb.lo[bestAxis];
            auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const Prim& p) {
                return std::min(NB - 1, std::max(0, (int)(((double)p.c[bestAxis] - lo) * k))) <= bestSplit;
            });
            mid = (uint32_t)(it - prims.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        return mid;
    }

    int build(uint32_t first, uint32_t count, int level = 1)
    {
        const int id = (int)nodes.size();
        nodes.emplace_back();
        Box box;
        const uint32_t mid = split(first, count, box, level);
        nodes[id].box = box;
        if (mid == first) { nodes[id].first = first; nodes[id].count = count; return id; }
        if (par_depth > 0 && count > 65536) {
            // large subtree: the two halves are independent (disjoint primitive ranges) -> build them on two threads
            // into private node arrays, then splice them in with an index offset
            Builder lb{prims, {}, leaf_size, par_depth - 1}, rb{prims, {}, leaf_size, par_depth - 1};
            std::thread t([&] { lb.build(first, mid - first, level + 1); });
            rb.build(mid, first + count - mid, level + 1);
            t.join();
            auto splice = [&](const std::vector<BNode>& sub) {
                const int off = (int)nodes.size();
                for (BNode n : sub) { if (n.left >= 0) { n.left += off; n.right += off; } nodes.push_back(n); }
                return off;
            };
            const int l = splice(lb.nodes), r = splice(rb.nodes);
            nodes[id].left = l; nodes[id].right = r;
            return id;
        }
        const int l = build(first, mid - first, level + 1);
        const int r = build(mid, first + count - mid, level + 1);
        nodes[id].left = l; nodes[id].right = r;
        return id;
    }
};

// Objects arrive as the caller's bytes (ipt_render_objects): coordinates may be infinite or NaN.  The reference's scan just
// never hits such an object or hits it everywhere; here the builder must stay well defined: centroids become finite
// (NaN -> 0), box bounds are clamped to +-FLT_MAX (a NaN coordinate never enters a box: std::min/max keep the other operand).
inline float finite_or(float v, float other) { return std::isnan(v) ? other : std::min(FLT_MAX, std::max(-FLT_MAX, v)); }
void sanitize(Prim& p)
{
    for (int k = 0; k < 3; k++) {
        p.c[k] = finite_or(p.c[k], 0.0f);
        p.box.lo[k] = std::max(p.box.lo[k], -FLT_MAX);
        p.box.hi[k] = std::min(p.box.hi[k], FLT_MAX);
    }
}

inline float down(double v, double pad) { return std::nextafterf((float)(v - pad), -INFINITY); }
inline float up(double v, double pad) { return std::nextafterf((float)(v + pad), INFINITY); }
inline double dabs(float v) { return std::fabs((double)v); }

void put_box(const Box& b, float* lo, float* hi)
{
    for (int k = 0; k < 3; k++) {
        const double pad = 2e-3 + 2e-6 * std::max(dabs(b.lo[k]), dabs(b.hi[k]));
        lo[k] = down(b.lo[k], pad); hi[k] = up(b.hi[k], pad);
    }
}

}  // namespace

static int build_bvh_impl(ipt_host_scene* s, uint32_t leaf_size, uint32_t brute_max)
{
    if (!s) return -1;
    s->bvh_nodes.clear(); s->bvh_slot_prim.clear();
    s->grid_cell_start.clear(); s->grid_refs.clear(); s->grid_big.clear();
    const uint32_t ns = (uint32_t)s->sphere_object.size(), nr = (uint32_t)s->rect_object.size(), n = ns + nr;
    if (n <= brute_max) { s->refresh_view(); return 0; }
    const auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (std::getenv("IPT_VERBOSE")) std::fprintf(stderr, "[bvh] %s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count()); };
    leaf_size = std::min(16u, std::max(1u, leaf_size));
    std::vector<Prim> prims(n);
    // boxes of a million primitives: 0.15 s on one thread, so spread over the host threads (independent slots)
    auto parallel_for = [](uint32_t count, auto body) {
        unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        if (const char* e = std::getenv("IPT_HOST_THREADS")) nt = (unsigned)std::max(1, std::min(256, std::atoi(e)));
        if (count < 65536 || nt == 1) { body(0u, count); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) th.emplace_back([=] { body((uint32_t)((uint64_t)count * t / nt), (uint32_t)((uint64_t)count * (t + 1) / nt)); });
        for (auto& x : th) x.join();
    };
    parallel_for(ns, [&](uint32_t i0, uint32_t i1) {
    for (uint32_t i = i0; i < i1; i++) {
        const double* c = &s->sphere_cxyzr[4 * (size_t)i];
        const double r = std::fabs(c[3]);
        Prim& p = prims[i];
        const double lo[3] = {c[0] - r, c[1] - r, c[2] - r}, hi[3] = {c[0] + r, c[1] + r, c[2] + r};
        p.box.grow(lo); p.box.grow(hi);
        for (int k = 0; k < 3; k++) p.c[k] = (float)c[k];
        p.ref = i;
        sanitize(p);
    }
    });
    parallel_for(nr, [&](uint32_t j0, uint32_t j1) {
    for (uint32_t j = j0; j < j1; j++) {
        const double *c = &s->rect_center[3 * (size_t)j], *N = &s->rect_north[3 * (size_t)j], *E = &s->rect_east[3 * (size_t)j];
        Prim& p = prims[ns + j];
        for (int sn = -1; sn <= 1; sn += 2)
            for (int se = -1; se <= 1; se += 2) {
                double q[3];
                for (int k = 0; k < 3; k++) q[k] = c[k] + sn * N[k] + se * E[k];
                p.box.grow(q);
            }
        for (int k = 0; k < 3; k++) p.c[k] = (float)c[k];
        p.ref = 0x80000000u | j;
        sanitize(p);
    }
    });
    // fork threads on the top levels of large scenes (2^par_depth subtrees in flight)
    int par_depth = 0;
    for (unsigned hw = std::max(1u, std::thread::hardware_concurrency()); (1u << par_depth) < hw && par_depth < 5;) par_depth++;
    if (const char* e = std::getenv("IPT_BVH_PAR_DEPTH")) par_depth = std::atoi(e);
    lap("prims");
    Builder b{prims, {}, leaf_size, par_depth};
    b.nodes.reserve(2 * (size_t)n / leaf_size + 16);
    const int root = b.build(0, n);
    lap("tree");
    s->bvh_slot_prim.resize(n);
    for (uint32_t i = 0; i < n; i++) s->bvh_slot_prim[i] = prims[i].ref;

    auto leaf_code = [](const BNode& nd) { return ~(int32_t)nd.first; };
    if (b.nodes[root].left < 0) {
        // the whole scene is one leaf: a root whose second child is an empty (inverted) box
        ipt_bvh_node o;
        std::memset(&o, 0, sizeof(o));
        put_box(b.nodes[root].box, o.lo0, o.hi0);
        for (int k = 0; k < 3; k++) { o.lo1[k] = FLT_MAX; o.hi1[k] = -FLT_MAX; }   // inverted: never hit
        o.child[0] = leaf_code(b.nodes[root]); o.count[0] = b.nodes[root].count;
        o.child[1] = ~0; o.count[1] = 1;
        s->bvh_nodes.push_back(o);
        s->refresh_view();
        return 1;
    }
    // breadth-first numbering of the inner nodes: the top levels get the lowest indices (they are staged in shared memory)
    std::vector<int> order, index(b.nodes.size(), -1);
    std::queue<int> q;
    q.push(root);
    while (!q.empty()) {
        const int id = q.front(); q.pop();
        index[id] = (int)order.size();
        order.push_back(id);
        const BNode& nd = b.nodes[id];
        if (b.nodes[nd.left].left >= 0) q.push(nd.left);
        if (b.nodes[nd.right].left >= 0) q.push(nd.right);
    }
    s->bvh_nodes.resize(order.size());
    for (size_t i = 0; i < order.size(); i++) {
        const BNode& nd = b.nodes[order[i]];
        ipt_bvh_node& o = s->bvh_nodes[i];
        std::memset(&o, 0, sizeof(o));
        const BNode &l = b.nodes[nd.left], &r = b.nodes[nd.right];
        put_box(l.box, o.lo0, o.hi0);
        put_box(r.box, o.lo1, o.hi1);
        if (l.left >= 0) o.child[0] = index[nd.left]; else { o.child[0] = leaf_code(l); o.count[0] = l.count; }
        if (r.left >= 0) o.child[1] = index[nd.right]; else { o.child[1] = leaf_code(r); o.count[1] = r.count; }
    }
    lap("emit");
    s->build_grid();
    lap("grid");
    s->refresh_view();
    return (int)s->bvh_nodes.size();
}

extern "C" int ipt_host_build_bvh(ipt_host_scene* s, uint32_t leaf_size, uint32_t brute_max)
{
    try { return build_bvh_impl(s, leaf_size, brute_max); }
    catch (...) {                                   // out of memory: nothing is thrown across the C ABI; the scene keeps no half-built tree
        if (s) { s->bvh_nodes.clear(); s->bvh_slot_prim.clear(); s->grid_cell_start.clear(); s->grid_refs.clear(); s->grid_big.clear(); s->refresh_view(); }
        return -1;
    }
}
