// ipt_wide.h — the 8-wide, quantised form of the bounding volume hierarchy that the fp32 traversal kernel walks
// (k_extend_wide, ipt_kernels.cuh), and the host routine that derives it from the 2-wide tree of the C ABI
// (ipt_bvh_node, include/ipt_abi.h).  New work: the reference has no acceleration structure (Renderer.cu:227-243 is a
// linear scan); what has to be preserved is the scan's answer, so every box here is a superset of the 2-wide tree's
// (already padded) box it stands for.
//
// One node = one 128-byte cache line:
//   bytes  0..31   origin.xyz, scale.xyz (fp32), 2 spare words      plane = origin + q * scale, q in 0..255 (origin = box - 1 step)
//   bytes 32..95   8 children x 8 bytes {lo.x, hi.x, lo.y, hi.y, lo.z, hi.z, 0, 0}   (lo rounded down, hi rounded up;
//                  an unused child has lo = 255 > hi = 0 and is never entered)
//   bytes 96..127  8 links: >= 0 index of an inner node; < 0 leaf, ~link = first_slot * 16 + (count - 1)
// A ray is walked by LPR cooperating lanes (1, 2 or 4), each decoding 8 / LPR children: the header is one broadcast
// 256-bit load, a lane's child boxes and links one vector load each - the lanes of a ray touch one line.
// Why this shape: profiles/r02_ncu_extend_v7.txt - the 2-wide per-lane traversal was bound by L1 wavefronts (86 %,
// every lane fetching its own 64-byte node, hit rate 3 %) at 15 of 32 lanes active.
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
    uint32_t reserved[2];
    uint8_t q[8][8];
    int32_t link[8];
};
static_assert(sizeof(WideNode) == 128, "one node per 128-byte line");
static constexpr int32_t WIDE_EMPTY = 0x7FFFFFFF;
// Slot s of a node is stored at position s of q[] / link[]; with LPR lanes per ray, lane l decodes slots
// [l * 8 / LPR, (l + 1) * 8 / LPR).
}  // namespace ipt

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "../../include/ipt_abi.h"

namespace ipt {

struct WideTree {
    std::vector<WideNode> nodes;
    uint32_t depth = 0;        // levels of inner nodes on the longest root-to-leaf path
    uint32_t stack_need = 0;   // most entries a traversal can have pending: sum over a path of (children - 1)
    double sum_children = 0;
};

// Collapses the 2-wide tree into 8-wide nodes: a node's child list starts as its two children and the inner child
// with the largest surface area is replaced by its own two children until there are eight (or only leaves are left).
// A 2-wide subtree of at most `leaf_max` primitives in consecutive slots becomes one leaf (the cooperating lanes test
// its primitives side by side, so a fuller leaf costs no more than a small one).  Nodes are emitted breadth-first.
// Requires child index > parent index for inner children (what host/bvh.cpp emits; it also rules out cycles).
inline const char* wide_collapse(const ipt_bvh_node* n2, uint32_t n_nodes, uint32_t n_slots, uint32_t leaf_max, WideTree& out)
{
    out = WideTree();
    if (!n2 || n_nodes == 0) return "no 2-wide tree";
    leaf_max = std::min(16u, std::max(1u, leaf_max));
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
    struct Pending { uint32_t n2, wide, depth, stack; };
    std::vector<Pending> queue;
    queue.push_back({0u, 0u, 1u, 0u});
    out.nodes.resize(1);
    std::vector<Item> items, tmp;
    for (size_t head = 0; head < queue.size(); head++) {
        const Pending p = queue[head];
        items.clear();
        child_items(p.n2, items);
        // first the subtrees that must stay inner nodes, then - while slots are left - the small subtrees that would
        // otherwise become one leaf each: more, tighter leaves cost the cooperating lanes nothing
        for (int pass = 0; pass < 2; pass++)
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
        out.stack_need = std::max(out.stack_need, stack_here);
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
            } else {
                const uint32_t w = (uint32_t)out.nodes.size();
                out.nodes.emplace_back();
                nd.link[j] = (int32_t)w;
                queue.push_back({(uint32_t)it.n2, w, p.depth + 1, stack_here});
            }
        }
        out.nodes[p.wide] = nd;
    }
    return nullptr;
}

}  // namespace ipt
