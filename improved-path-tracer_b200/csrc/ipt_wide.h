// ipt_wide.h — the 8-wide, quantised form of the bounding volume hierarchy that the fp32 traversal kernel walks
// (k_extend_cw, ipt_kernels.cuh; opt-in: IPT_BVH8=1), and the host routine that derives it from the 2-wide tree of the C ABI
// (ipt_bvh_node, include/ipt_abi.h).  New work: the reference has no acceleration structure (Renderer.cu:227-243 is a
// linear scan); what has to be preserved is the scan's answer, so every box here is a superset of the 2-wide tree's
// (already padded) box it stands for.
//
// One node = one 128-byte cache line:
//   bytes  0..23   origin.xyz, scale.xyz (fp32)                     plane = origin + q * scale, q in 0..255 (origin = box - 1 step)
//   bytes 24..27   child_base (24 bits) | inner mask << 24          the inner children of a node are consecutive nodes, in slot
//                                                                   order: child in slot s = child_base + popc(imask & ((1 << s) - 1))
//   bytes 28..31   primitive mask: bits [4s, 4s + count) set for a leaf child in slot s (a leaf holds 1..4 primitives)
//   bytes 32..95   8 children x 8 bytes {lo.x, hi.x, lo.y, hi.y, lo.z, hi.z, -, -}   (lo rounded down, hi rounded up;
//                  an unused child has lo = 255 > hi = 0 and is never entered); the spare bytes of children 0 and 1 hold
//                  prim_base (low and high half): the primitives of the node's leaf children are consecutive slots, in
//                  bit order of the primitive mask: bit b = slot prim_base + popc(pmask & ((1 << b) - 1))
//   bytes 96..127  8 links (the same information per child, for tools and tests): >= 0 index of an inner node; < 0 leaf,
//                  ~link = first_slot * 16 + (count - 1)
// The traversal (k_extend_cw) walks one ray per lane and reads 96 bytes per node with three 256-bit loads.
// Why this shape: profiles/r02_ncu_extend_v7.txt - the 2-wide per-lane traversal was bound by L1 wavefronts (86 %,
// every lane fetching its own 64-byte node 53 times per ray, hit rate 3 %) at 17 of 32 lanes active.
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace ipt {

struct alignas(128) WideNode {
    float origin[3];
    float scale[3];
    uint32_t base_imask;     // child_base | imask << 24
    uint32_t pmask;          // bits [4s, 4s + count) for a leaf child in slot s
    uint8_t q[8][8];
    int32_t link[8];
};
static_assert(sizeof(WideNode) == 128, "one node per 128-byte line");
static constexpr int32_t WIDE_EMPTY = 0x7FFFFFFF;
// Slot s of a node is stored at position s of q[] / link[].
}  // namespace ipt

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../include/ipt_abi.h"

namespace ipt {

struct WideTree {
    std::vector<WideNode> nodes;
    std::vector<uint32_t> perm;   // the slot order the tree asks for: new slot i holds what was slot perm[i] of the 2-wide tree's
                                  // order (whole leaves move, so every leaf of the 2-wide tree stays a consecutive range)
    std::vector<uint32_t> inv;    // old slot -> new slot
    uint32_t depth = 0;        // levels of inner nodes on the longest root-to-leaf path
    double sum_children = 0;
};

// Collapses the 2-wide tree into 8-wide nodes.  Which 2-wide nodes become the children of a wide node is chosen by the
// surface-area cost of the result (the dynamic programme of Ylitie, Karras, Laine 2017, section 3.1): C(n, j) is the cheapest way
// to stand for the subtree of n with at most j children of one wide node - one child (a leaf, or an inner node whose own
// children are the best split of n into at most 8) or the best distribution of j over n's two children; a visited node
// costs c_node, a tested primitive c_prim, each weighted by the surface area of its box.  (`greedy`: the first version -
// open the child of largest area until there are eight - which left 3.6 children per node on the 1M-primitive scene.)
// A 2-wide subtree of at most `leaf_max` (<= 4) primitives in consecutive slots may become one leaf.  Nodes are emitted
// breadth-first.  Requires child index > parent index for inner children (what host/bvh.cpp emits; it also rules out cycles).
inline const char* wide_collapse(const ipt_bvh_node* n2, uint32_t n_nodes, uint32_t n_slots, uint32_t leaf_max, WideTree& out,
                                 bool greedy = false, double c_node = 4.0, double c_prim = 1.0)
{
    out = WideTree();
    if (!n2 || n_nodes == 0) return "no 2-wide tree";
    leaf_max = std::min(4u, std::max(1u, leaf_max));     // a leaf child owns 4 bits of the node's primitive mask
    auto valid_box = [](const float* lo, const float* hi) { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; };
    // slot range and contiguity of every 2-wide subtree, bottom-up
    std::vector<uint32_t> first(n_nodes), count(n_nodes);
    std::vector<uint8_t> contig(n_nodes);
    for (uint32_t i = n_nodes; i-- > 0;) {
        uint32_t f[2], c[2]; bool g[2], present[2];
        for (int k = 0; k < 2; k++) {
            const int32_t ch = n2[i].child[k];
            present[k] = valid_box(k ? n2[i].lo1 : n2[i].lo0, k ? n2[i].hi1 : n2[i].hi0);
            if (ch >= 0) {
                if ((uint32_t)ch <= i || (uint32_t)ch >= n_nodes) return "2-wide tree: inner children must follow their parent";
                f[k] = first[ch]; c[k] = count[ch]; g[k] = contig[ch] != 0;
            } else {
                f[k] = (uint32_t)~ch; c[k] = n2[i].count[k]; g[k] = true;
                if (c[k] < 1 || c[k] > 16 || (uint64_t)f[k] + c[k] > n_slots) return "2-wide tree: bad leaf";
                if (c[k] > 4) return "unsupported: a leaf of more than 4 primitives";
            }
        }
        if (!present[0] && !present[1]) return "2-wide tree: node without children";
        if (!present[1]) { first[i] = f[0]; count[i] = c[0]; contig[i] = g[0]; continue; }
        if (!present[0]) { first[i] = f[1]; count[i] = c[1]; contig[i] = g[1]; continue; }
        count[i] = c[0] + c[1];
        first[i] = std::min(f[0], f[1]);
        contig[i] = g[0] && g[1] && (f[0] + c[0] == f[1] || f[1] + c[1] == f[0]);
    }
    struct Item { int32_t n2; bool leaf; uint32_t first, count; float lo[3], hi[3]; };   // n2 < 0: a leaf of the 2-wide tree
    auto child_items = [&](uint32_t i, std::vector<Item>& dst) {
        for (int k = 0; k < 2; k++) {
            Item it;
            std::memcpy(it.lo, k ? n2[i].lo1 : n2[i].lo0, 12); std::memcpy(it.hi, k ? n2[i].hi1 : n2[i].hi0, 12);
            if (!valid_box(it.lo, it.hi)) continue;
            const int32_t ch = n2[i].child[k];
            if (ch < 0) { it.n2 = -1; it.leaf = true; it.first = (uint32_t)~ch; it.count = n2[i].count[k]; }
            else { it.n2 = ch; it.leaf = contig[ch] && count[ch] <= leaf_max; it.first = first[ch]; it.count = count[ch]; }
            dst.push_back(it);
        }
    };
    auto area = [](const Item& it) {
        const double x = (double)it.hi[0] - it.lo[0], y = (double)it.hi[1] - it.lo[1], z = (double)it.hi[2] - it.lo[2];
        return x * y + y * z + z * x;
    };
    // ---- the dynamic programme, bottom-up (children have higher indices than their parent)
    struct Cost { float c[9]; uint8_t split[9]; uint8_t eff[9]; uint8_t one_is_leaf; uint8_t alias; };   // index j = 1..8 (8: Cdist only)
    std::vector<Cost> dp;
    std::vector<float> areaN;
    auto box_area = [](const float* lo, const float* hi) {
        const double x = (double)hi[0] - lo[0], y = (double)hi[1] - lo[1], z = (double)hi[2] - lo[2];
        return (float)(x * y + y * z + z * x);
    };
    if (!greedy) {
        dp.resize(n_nodes);
        areaN.assign(n_nodes, 0.f);
        {   // areas top-down: a node's box is stored in its parent
            float lo[3], hi[3];
            for (int k = 0; k < 3; k++) { lo[k] = 3e38f; hi[k] = -3e38f; }
            if (valid_box(n2[0].lo0, n2[0].hi0)) for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], n2[0].lo0[k]); hi[k] = std::max(hi[k], n2[0].hi0[k]); }
            if (valid_box(n2[0].lo1, n2[0].hi1)) for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], n2[0].lo1[k]); hi[k] = std::max(hi[k], n2[0].hi1[k]); }
            areaN[0] = box_area(lo, hi);
            for (uint32_t i = 0; i < n_nodes; i++)
                for (int k = 0; k < 2; k++)
                    if (n2[i].child[k] >= 0 && valid_box(k ? n2[i].lo1 : n2[i].lo0, k ? n2[i].hi1 : n2[i].hi0))
                        areaN[n2[i].child[k]] = box_area(k ? n2[i].lo1 : n2[i].lo0, k ? n2[i].hi1 : n2[i].hi0);
        }
        const float INF = 3e38f;
        // cost of child k of node i as at most j children of a wide node
        auto child_cost = [&](uint32_t i, int k, int j) -> float {
            const int32_t ch = n2[i].child[k];
            if (ch >= 0) return dp[ch].c[std::min(j, 7)];
            return box_area(k ? n2[i].lo1 : n2[i].lo0, k ? n2[i].hi1 : n2[i].hi0) * (float)n2[i].count[k] * (float)c_prim;
        };
        for (uint32_t i = n_nodes; i-- > 0;) {
            Cost& d = dp[i];
            std::memset(&d, 0, sizeof(d));
            const bool p0 = valid_box(n2[i].lo0, n2[i].hi0), p1 = valid_box(n2[i].lo1, n2[i].hi1);
            if (!(p0 && p1)) {       // one child only: the node stands for that child
                d.alias = p0 ? 1 : 2;
                for (int j = 1; j <= 8; j++) d.c[j] = child_cost(i, p0 ? 0 : 1, j);
                continue;
            }
            float dist[9];
            for (int j = 2; j <= 8; j++) {
                dist[j] = INF;
                for (int a = 1; a < j; a++) {
                    const float v = child_cost(i, 0, a) + child_cost(i, 1, j - a);
                    if (v < dist[j]) { dist[j] = v; d.split[j] = (uint8_t)a; }
                }
            }
            const float as_leaf = (contig[i] && count[i] <= leaf_max) ? areaN[i] * (float)count[i] * (float)c_prim : INF;
            const float as_node = areaN[i] * (float)c_node + dist[8];
            d.one_is_leaf = as_leaf <= as_node;
            d.c[1] = std::min(as_leaf, as_node); d.eff[1] = 1;
            for (int j = 2; j <= 7; j++) {
                if (dist[j] < d.c[j - 1]) { d.c[j] = dist[j]; d.eff[j] = (uint8_t)j; }
                else { d.c[j] = d.c[j - 1]; d.eff[j] = d.eff[j - 1]; }
            }
            d.c[8] = dist[8]; d.eff[8] = 8;
        }
    }
    // the children of the wide node that stands for 2-wide node i: the best split of i into at most 8
    std::vector<Item> acc;
    struct Frame { uint32_t node; int side; int j; };      // side -1: the node itself with budget j; else child `side` of `node`
    auto emit_optimal = [&](uint32_t root_n2, std::vector<Item>& dst) {
        std::vector<Frame> st;
        st.push_back({root_n2, -1, 8});
        bool top = true;
        while (!st.empty()) {
            Frame f = st.back(); st.pop_back();
            uint32_t i = f.node;
            if (f.side >= 0) {
                const int32_t ch = n2[i].child[f.side];
                if (ch < 0) {    // a leaf of the 2-wide tree
                    Item it;
                    std::memcpy(it.lo, f.side ? n2[i].lo1 : n2[i].lo0, 12); std::memcpy(it.hi, f.side ? n2[i].hi1 : n2[i].hi0, 12);
                    it.n2 = -1; it.leaf = true; it.first = (uint32_t)~ch; it.count = n2[i].count[f.side];
                    dst.push_back(it);
                    continue;
                }
                // an inner child with budget j: as one child of the wide node, or distributed further
                const Cost& d = dp[ch];
                if (d.alias) { st.push_back({(uint32_t)ch, d.alias - 1, f.j}); continue; }
                const int j = d.eff[std::min(f.j, 7)];
                if (j == 1) {
                    Item it;
                    std::memcpy(it.lo, f.side ? n2[i].lo1 : n2[i].lo0, 12); std::memcpy(it.hi, f.side ? n2[i].hi1 : n2[i].hi0, 12);
                    it.n2 = ch; it.leaf = d.one_is_leaf != 0; it.first = first[ch]; it.count = count[ch];
                    dst.push_back(it);
                    continue;
                }
                st.push_back({(uint32_t)ch, 1, j - d.split[j]});
                st.push_back({(uint32_t)ch, 0, d.split[j]});
                continue;
            }
            // the node that becomes a wide node: its own best split into at most 8
            const Cost& d = dp[i];
            if (d.alias) { st.push_back({i, d.alias - 1, top ? 8 : f.j}); top = false; continue; }
            top = false;
            st.push_back({i, 1, 8 - d.split[8]});
            st.push_back({i, 0, d.split[8]});
        }
    };
    struct Pending { uint32_t n2, wide, depth, stack; };
    std::vector<Pending> queue;
    queue.push_back({0u, 0u, 1u, 0u});
    out.nodes.resize(1);
    std::vector<Item> items, tmp;
    for (size_t head = 0; head < queue.size(); head++) {
        const Pending p = queue[head];
        items.clear();
        if (!greedy) emit_optimal(p.n2, items);
        else child_items(p.n2, items);
        // greedy: first the subtrees that must stay inner nodes, then - while slots are left - the small subtrees that would
        // otherwise become one leaf each
        for (int pass = 0; pass < 2 && greedy; pass++)
            while (items.size() < 8) {
                int best = -1; double ba = -1;
                for (size_t k = 0; k < items.size(); k++)
                    if (items[k].n2 >= 0 && items[k].leaf == (pass == 1) && area(items[k]) > ba) { ba = area(items[k]); best = (int)k; }
                if (best < 0) break;
                tmp.clear();
                child_items((uint32_t)items[best].n2, tmp);
                items.erase(items.begin() + best);
                items.insert(items.end(), tmp.begin(), tmp.end());
            }
        WideNode nd;
        std::memset(&nd, 0, sizeof(nd));
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (const Item& it : items) for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], (double)it.lo[k]); hi[k] = std::max(hi[k], (double)it.hi[k]); }
        // The grid of a node starts one step below its box and spends 253 steps on it: every child plane can then be moved
        // outwards by a whole step without leaving 0..255.  That step covers the traversal's decoding error (it evaluates
        // t = (2^23 + q) A + (B - 2^23 A) in fp32, off by up to half a step) on top of the 2-wide tree's own padding.
        double sc[3];
        for (int k = 0; k < 3; k++) {
            float s = (float)std::max((hi[k] - lo[k]) / 253.0 * (1.0 + 1e-6), 1e-30);
            for (;;) {
                nd.origin[k] = std::nextafterf((float)(lo[k] - (double)s), -INFINITY);
                if ((double)nd.origin[k] + (double)s <= lo[k] && (double)nd.origin[k] + 254.0 * (double)s >= hi[k]) break;
                s = std::nextafterf(s, INFINITY) * 1.000001f;
            }
            nd.scale[k] = s; sc[k] = (double)s;
        }
        const uint32_t stack_here = p.stack + (uint32_t)items.size() - 1;
        out.depth = std::max(out.depth, p.depth);
        out.sum_children += (double)items.size();
        // Slot assignment (Ylitie, Karras, Laine 2017): slot s stands for the corner direction (s&1 ? +x : -x, s&2 ? +y : -y,
        // s&4 ? +z : -z); children go greedily to the slots their centre (relative to the node's) points at most.  A ray
        // with direction signs `oct` (bit k set = negative along axis k) then meets the slots roughly front to back in
        // increasing s ^ oct, which is the order the traversal pushes the children it does not enter at once.
        int slot_of[8], child_at[8];
        for (int j = 0; j < 8; j++) { slot_of[j] = -1; child_at[j] = -1; }
        {
            double score[8][8];
            for (size_t c = 0; c < items.size(); c++)
                for (int sl = 0; sl < 8; sl++) {
                    double v = 0;
                    for (int k = 0; k < 3; k++) v += (0.5 * ((double)items[c].lo[k] + items[c].hi[k]) - 0.5 * (lo[k] + hi[k])) * ((sl >> k) & 1 ? 1.0 : -1.0);
                    score[c][sl] = v;
                }
            for (size_t n = 0; n < items.size(); n++) {
                int bc = -1, bs = -1; double bv = -1e300;
                for (size_t c = 0; c < items.size(); c++) {
                    if (slot_of[c] >= 0) continue;
                    for (int sl = 0; sl < 8; sl++) if (child_at[sl] < 0 && score[c][sl] > bv) { bv = score[c][sl]; bc = (int)c; bs = sl; }
                }
                slot_of[bc] = bs; child_at[bs] = bc;
            }
        }
        for (int sl = 0; sl < 8; sl++) {
            const int j = sl;
            uint8_t* q = nd.q[j];
            if (child_at[sl] < 0) { q[0] = q[2] = q[4] = 255; q[1] = q[3] = q[5] = 0; nd.link[j] = WIDE_EMPTY; continue; }
            const Item& it = items[child_at[sl]];
            for (int k = 0; k < 3; k++) {
                // rounded outwards, with a margin of 1e-3 of a step for the rounding of (x - origin) / scale itself
                const double a = std::floor(((double)it.lo[k] - (double)nd.origin[k]) / sc[k] - 1e-3) - 1.0;
                const double b = std::ceil(((double)it.hi[k] - (double)nd.origin[k]) / sc[k] + 1e-3) + 1.0;
                q[2 * k] = (uint8_t)std::min(255.0, std::max(0.0, a));
                q[2 * k + 1] = (uint8_t)std::min(255.0, std::max(0.0, b));
            }
            if (it.leaf) {
                if (it.first >= (1u << 27)) return "too many primitive slots for the leaf encoding";
                nd.link[j] = ~(int32_t)((it.first << 4) | (it.count - 1));
                nd.pmask |= ((1u << it.count) - 1u) << (4 * j);
            } else {
                const uint32_t w = (uint32_t)out.nodes.size();
                if (w >= (1u << 24)) return "too many nodes for the child_base encoding";
                if (!(nd.base_imask >> 24)) nd.base_imask = w;          // first inner child; the others follow it in slot order
                nd.base_imask |= 1u << (24 + j);
                out.nodes.emplace_back();
                nd.link[j] = (int32_t)w;
                queue.push_back({(uint32_t)it.n2, w, p.depth + 1, stack_here});
            }
        }
        out.nodes[p.wide] = nd;
    }
    // Slot order: node by node, the primitives of a node's leaf children in slot order (= bit order of pmask).
    out.perm.reserve(n_slots);
    out.inv.assign(n_slots, 0xFFFFFFFFu);
    for (WideNode& nd : out.nodes) {
        const uint32_t prim_base = (uint32_t)out.perm.size();
        nd.q[0][6] = (uint8_t)prim_base; nd.q[0][7] = (uint8_t)(prim_base >> 8); nd.q[1][6] = (uint8_t)(prim_base >> 16); nd.q[1][7] = (uint8_t)(prim_base >> 24);
        for (int j = 0; j < 8; j++) {
            if (nd.link[j] >= 0) continue;
            const uint32_t code = (uint32_t)~nd.link[j], first = code >> 4, cnt = (code & 15u) + 1u, moved = (uint32_t)out.perm.size();
            for (uint32_t k = 0; k < cnt; k++) {
                if (out.inv[first + k] != 0xFFFFFFFFu) return "2-wide tree: a primitive slot is referenced twice";
                out.inv[first + k] = (uint32_t)out.perm.size();
                out.perm.push_back(first + k);
            }
            nd.link[j] = ~(int32_t)((moved << 4) | (cnt - 1));
        }
    }
    if (out.perm.size() != n_slots) return "2-wide tree: not every primitive slot is referenced";
    return nullptr;
}

}  // namespace ipt
