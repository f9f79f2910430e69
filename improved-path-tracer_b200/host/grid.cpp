// grid.cpp — uniform grid over the slots of the BVH (ipt_scene::grid_*), for scenes whose primitives are small against
// their spacing and evenly spread (BASELINE config 5: a million primitives of radius 1-4 on a jittered lattice).  New work:
// the reference has no acceleration structure (Renderer.cu:227-243 scans every object); what has to be preserved is the
// scan's answer.  A ray walking a hierarchy over such a scene crosses ~49 inner nodes to find ~6 leaves
// (profiles/README.md, round 2); walking the cells its segment passes through needs no hierarchy at all.
//
// Which scenes qualify (anything else keeps the BVH pipeline): at most 64 "big" primitives (more cells than BIG_CELLS
// under their box: walls, large lights - they go into a list every ray tests), at most 12 references per primitive and
// 16 per occupied cell on average.  Cell size: IPT_GRID_DENSITY (default 0.35) small primitives per cell, cubic cells.
// IPT_NO_GRID=1 switches the grid off (A/B runs, and the tests that want the BVH kernels).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "host_scene.hpp"

namespace {

struct Box { double lo[3], hi[3]; bool valid; };

template <typename F> void parallel_for(size_t n, F f)
{
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if (const char* e = std::getenv("IPT_HOST_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    nt = (unsigned)std::min<size_t>(nt, std::max<size_t>(1, n / 4096));
    if (nt <= 1) { f(0, n); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back([=] { f(n * t / nt, n * (t + 1) / nt); });
    for (auto& x : th) x.join();
}

}  // namespace

void ipt_host_scene::build_grid()
{
    grid_cell_start.clear(); grid_refs.clear(); grid_big.clear();
    grid_res[0] = grid_res[1] = grid_res[2] = 0;
    const size_t n = bvh_slot_prim.size();
    if (n == 0 || bvh_nodes.empty() || std::getenv("IPT_NO_GRID")) return;
    const uint32_t ns = (uint32_t)sphere_object.size();
    // bounding box per slot, padded: the traversal's arithmetic is fp32 (cell boundaries and hit points are off by ~1e-4 at
    // |x| ~ 1e3, and a rectangle accepts hits 5e-5 outside its edges, Plane.cu:87-100), the padding is 40 times that
    std::vector<Box> box(n);
    parallel_for(n, [&](size_t a, size_t b) {
        for (size_t s = a; s < b; s++) {
            const uint32_t prim = bvh_slot_prim[s];
            Box& bx = box[s];
            bx.valid = true;
            if (!(prim & 0x80000000u)) {
                const double* c = &sphere_cxyzr[4 * (size_t)prim];
                const double r = std::fabs(c[3]);
                for (int k = 0; k < 3; k++) { bx.lo[k] = c[k] - r; bx.hi[k] = c[k] + r; }
                (void)ns;
            } else {
                const size_t j = prim & 0x7FFFFFFFu;
                const double *c = &rect_center[3 * j], *N = &rect_north[3 * j], *E = &rect_east[3 * j];
                for (int k = 0; k < 3; k++) {
                    const double e = std::fabs(N[k]) + std::fabs(E[k]);
                    bx.lo[k] = c[k] - e; bx.hi[k] = c[k] + e;
                }
            }
            for (int k = 0; k < 3; k++) {
                if (!(bx.lo[k] == bx.lo[k]) || !(bx.hi[k] == bx.hi[k]) || !std::isfinite(bx.lo[k]) || !std::isfinite(bx.hi[k])) bx.valid = false;
                const double pad = 4e-3 + 1e-6 * std::max(std::fabs(bx.lo[k]), std::fabs(bx.hi[k]));
                bx.lo[k] -= pad; bx.hi[k] += pad;
            }
        }
    });
    // a first guess at "big": longest side above 16 times the median longest side
    std::vector<float> side(n);
    for (size_t s = 0; s < n; s++) side[s] = box[s].valid ? (float)std::max({box[s].hi[0] - box[s].lo[0], box[s].hi[1] - box[s].lo[1], box[s].hi[2] - box[s].lo[2]}) : 0.f;
    std::vector<float> tmp(side);
    std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
    const float big_side = 16.f * std::max(tmp[n / 2], 1e-6f);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    size_t n_small = 0;
    std::vector<uint8_t> is_big(n, 0);
    for (size_t s = 0; s < n; s++) {
        if (!box[s].valid || side[s] > big_side) { is_big[s] = 1; continue; }   // an unboundable primitive is tested for every ray
        n_small++;
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], box[s].lo[k]); hi[k] = std::max(hi[k], box[s].hi[k]); }
    }
    if (n_small < 64 || n - n_small > 64) return;
    double ext[3], vol = 1;
    for (int k = 0; k < 3; k++) { ext[k] = std::max(hi[k] - lo[k], 1e-3); vol *= ext[k]; }
    const double density = std::getenv("IPT_GRID_DENSITY") ? std::max(0.01, std::atof(std::getenv("IPT_GRID_DENSITY"))) : 0.35;
    const double cell = std::cbrt(vol * density / (double)n_small);
    uint64_t n_cells = 1;
    for (int k = 0; k < 3; k++) {
        grid_res[k] = (uint32_t)std::min(1024.0, std::max(1.0, std::floor(ext[k] / cell + 0.5)));
        grid_cell[k] = (float)(ext[k] / grid_res[k] * (1.0 + 1e-6));
        grid_lo[k] = std::nextafterf((float)lo[k], -INFINITY);
        n_cells *= grid_res[k];
    }
    if (n_cells > (1ull << 26)) { grid_res[0] = 0; return; }
    auto range = [&](const Box& b, int k, uint32_t& a, uint32_t& z) {
        const double inv = 1.0 / (double)grid_cell[k];
        const double fa = std::floor((b.lo[k] - (double)grid_lo[k]) * inv), fz = std::floor((b.hi[k] - (double)grid_lo[k]) * inv);
        a = (uint32_t)std::min<double>(grid_res[k] - 1, std::max(0.0, fa));
        z = (uint32_t)std::min<double>(grid_res[k] - 1, std::max(0.0, fz));
    };
    // a sphere is filed under the cells it actually reaches (centre-to-cell distance <= padded radius), not under every cell of its
    // bounding box: the corner cells of the box are the ones a small sphere misses
    std::vector<float> sph_r(n, -1.f);
    std::vector<double> sph_c(3 * n, 0.0);
    const bool box_filing = std::getenv("IPT_GRID_BOX_FILING") != nullptr;   // A/B knob: every cell of the bounding box
    for (size_t s = 0; s < n && !box_filing; s++) {
        const uint32_t prim = bvh_slot_prim[s];
        if (prim & 0x80000000u || !box[s].valid) continue;
        const double* c = &sphere_cxyzr[4 * (size_t)prim];
        sph_r[s] = (float)(std::fabs(c[3]) + 4e-3 + 1e-6 * std::max({std::fabs(c[0]), std::fabs(c[1]), std::fabs(c[2])}) + 1e-6 * std::fabs(c[3]));
        for (int k = 0; k < 3; k++) sph_c[3 * s + k] = c[k];
    }
    auto reaches = [&](size_t s, uint32_t x, uint32_t y, uint32_t z) {
        if (sph_r[s] < 0.f) return true;
        const uint32_t idx[3] = {x, y, z};
        double d2 = 0;
        for (int k = 0; k < 3; k++) {
            const double a = (double)grid_lo[k] + (double)idx[k] * (double)grid_cell[k], b = a + (double)grid_cell[k], c = sph_c[3 * s + k];
            const double d = c < a ? a - c : (c > b ? c - b : 0.0);
            d2 += d * d;
        }
        return d2 <= (double)sph_r[s] * (double)sph_r[s];
    };
    std::vector<std::atomic<uint32_t>> count(n_cells + 1);
    parallel_for(n_cells + 1, [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) count[i].store(0, std::memory_order_relaxed); });
    const uint32_t rx = grid_res[0], ry = grid_res[1];
    parallel_for(n, [&](size_t a, size_t b) {
        for (size_t s = a; s < b; s++) {
            if (is_big[s]) continue;
            uint32_t x0, x1, y0, y1, z0, z1;
            range(box[s], 0, x0, x1); range(box[s], 1, y0, y1); range(box[s], 2, z0, z1);
            for (uint32_t z = z0; z <= z1; z++) for (uint32_t y = y0; y <= y1; y++) for (uint32_t x = x0; x <= x1; x++)
                if (reaches(s, x, y, z)) count[x + (size_t)rx * (y + (size_t)ry * z)].fetch_add(1, std::memory_order_relaxed);
        }
    });
    grid_cell_start.resize(n_cells + 1);
    uint64_t total = 0, occupied = 0;
    for (size_t c = 0; c < n_cells; c++) {
        const uint32_t k = count[c].load(std::memory_order_relaxed);
        grid_cell_start[c] = (uint32_t)total;
        count[c].store((uint32_t)total, std::memory_order_relaxed);   // becomes the fill cursor
        total += k; occupied += k != 0;
    }
    grid_cell_start[n_cells] = (uint32_t)total;
    const bool ok = total < (1ull << 31) && total <= 12ull * n_small && (occupied == 0 || total <= 16ull * occupied);
    if (std::getenv("IPT_VERBOSE"))
        std::fprintf(stderr, "[grid] %u x %u x %u cells of %.2f x %.2f x %.2f, %zu small + %zu big primitives, %.2f references per primitive, %.2f per occupied cell (%.0f %% occupied)%s\n",
                     grid_res[0], grid_res[1], grid_res[2], grid_cell[0], grid_cell[1], grid_cell[2], n_small, n - n_small, (double)total / n_small,
                     occupied ? (double)total / occupied : 0.0, 100.0 * occupied / n_cells, ok ? "" : " - rejected, the BVH pipeline stays");
    if (!ok) { grid_cell_start.clear(); grid_res[0] = grid_res[1] = grid_res[2] = 0; return; }
    grid_refs.resize(total);
    parallel_for(n, [&](size_t a, size_t b) {
        for (size_t s = a; s < b; s++) {
            if (is_big[s]) continue;
            uint32_t x0, x1, y0, y1, z0, z1;
            range(box[s], 0, x0, x1); range(box[s], 1, y0, y1); range(box[s], 2, z0, z1);
            for (uint32_t z = z0; z <= z1; z++) for (uint32_t y = y0; y <= y1; y++) for (uint32_t x = x0; x <= x1; x++)
                if (reaches(s, x, y, z)) grid_refs[count[x + (size_t)rx * (y + (size_t)ry * z)].fetch_add(1, std::memory_order_relaxed)] = (uint32_t)s;
        }
    });
    // references of a cell in slot order: the order threads filled them in must not show anywhere (it does not change a
    // frame - ties are resolved by object index - but it would change the arrays from run to run)
    parallel_for(n_cells, [&](size_t a, size_t b) {
        for (size_t c = a; c < b; c++) std::sort(grid_refs.begin() + grid_cell_start[c], grid_refs.begin() + grid_cell_start[c + 1]);
    });
    for (size_t s = 0; s < n; s++) if (is_big[s]) grid_big.push_back((uint32_t)s);
}
