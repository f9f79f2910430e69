// ipt_kernels.cuh — the wavefront kernels of the B200 radiance path.
//
// One *batch* is a contiguous range of camera samples (default 64 Mi).  A batch is advanced by kernel launches
// ("passes"): pass p reads the ray queue written by pass p-1, finds the nearest hit, accumulates emission, scatters, and
// appends the continuation rays to the other queue, compacted with warp ballot + popc - one bounce per pass in the
// generic and BVH pipelines, several in the typed-list kernel.  Queue lengths live on the device and every launch is
// a persistent grid (a fixed number of CTAs per SM), so launch geometry never depends on the queue length and the
// host never synchronises inside a render.  Three pipelines share the device functions of ipt_device.cuh:
//   k_bounce_fast<FIRST, SHAPE>          fp32, scenes without a BVH (the shipped scenes, the 4K config): typed primitive
//                                        lists in shared memory (straight-line scan for box rooms), static slice
//                                        schedule with cp.async prefetch, up to 8 bounces per ray in registers between
//                                        two queue round trips (schedule kept on the device, fast_schedule),
//                                        warp-private output blocks (one atomic per 128 outputs);
//   k_raygen -> k_extend_bvh -> k_bounce<float, MODE_SHADE>
//                                        fp32, BVH scenes: traversal with lane-level refill split from shading;
//   k_bounce<R, MODE_BRUTE|MODE_BVH, FIRST, DEFER>
//                                        the generic fused step: fp64 parity mode, and maxDepth >= 130 (DEFER).
//
// Reference being replaced: the single kernel cudaMain<<<22,22>>> (Renderer.cu:254-265) in which each of 484 threads
// walks ~1900 pixels x spp x bounces serially with recursion (firstLayer/secondLayer/deepLayers, :149-225).
#pragma once
#include "ipt_device.cuh"
#include "ipt_wide.h"

namespace ipt {

enum { MODE_BRUTE = 0, MODE_BVH = 1, MODE_SHADE = 2 };   // MODE_SHADE: hits were found by k_extend_bvh, k_bounce only shades

static constexpr int BLOCK_THREADS = 256;
static constexpr int GRAB = 128;          // generic kernels: rays a warp claims per atomic on the work counter (4 iterations of 32)
static constexpr int BVH_TOP_NODES = 512; // top of the BVH staged in shared memory (32 KB)

// counters[] layout (uint32): [CNT + p] rays queued for pass p, [WORK + p] work-claim counter of pass p
// (k_bounce_fast claims no work: there [WORK + p] is the depth pass p starts at and [WORK_EXTEND + p] its bounce count)
static constexpr int MAX_PASSES = 256;
static constexpr int CNT = 0, WORK = MAX_PASSES, WORK_EXTEND = 2 * MAX_PASSES;   // WORK_EXTEND: claim counter of k_extend_bvh
static constexpr int CLAIM = 3 * MAX_PASSES;   // k_bounce_fast: next unclaimed ray of pass p (slices of 32 are claimed by warps)
static constexpr int N_COUNTERS = 4 * MAX_PASSES;

// Uniform grid (ipt_scene::grid_*, host/grid.cpp) as k_extend_grid sees it
static constexpr uint32_t GRID_MAX_BIG = 64;
struct GridHeader {
    uint32_t res[3];
    uint32_t n_big;
    float lo[3], hi[3], cs[3], inv_cs[3];
    uint32_t big[GRID_MAX_BIG];     // slots of the primitives every ray tests
};

template <typename R> struct KParams {
    SceneView<R> sc;
    // camera, RenderController.cu:39 + Renderer.cu:112-147
    V3<R> camO, camD, camX, camZ;
    R fov;                      // (R)0.0009f  (Renderer.cu:27 is a float constant)
    uint32_t W, H, spp, maxDepth;
    uint32_t depth;             // depth of every ray in this pass
    uint32_t strat_n;           // floor(sqrt(spp)) for IPT_FLAG_STRATIFIED
    uint32_t flags;
    PhiloxKeys keys;
    // tile schedule of this rank: local tile lt -> global tile id tile_ids[lt]
    const uint32_t* tile_ids;
    uint32_t n_tiles_local, tiles_x, tile_w, tile_h;
    uint32_t mt_x, mt_per_tile; // micro-tiles (8x4 pixels) per tile row / per tile
    // the micro-tiles of this rank that can see the scene (k_active_microtiles), packed (y << 16 | x) in units of 8x4
    // pixels, in tile-major order; the sample ids of a render enumerate only these
    const uint32_t* mt_list;
    uint32_t n_mt;
    // batch
    uint32_t base_mt, mt_count, base_sample, n_first;   // pass 0: the batch = micro-tiles [base_mt, base_mt + mt_count) x samples
                                                         // from base_sample on; n_first sample ids (32 per micro-tile and sample)
    Queue qin, qout;
    uint32_t* counters;
    unsigned long long* traced; // total nearest-hit queries (stats)
    unsigned long long* frame;  // 3 x 64-bit accumulators per pixel (fixed point, or fp64 bits with IPT_FLAG_FLOAT_ACCUM)
    double fixed_scale;
    uint32_t refill_min;        // k_extend_bvh: idle lanes per warp that trigger a refill
    uint32_t descend_min;       // k_extend_bvh: the descent phase ends when fewer lanes than this are still descending
    uint2* hits;                // split pipeline: {t bits, slot} per ray of the current pass (k_extend_bvh -> k_bounce<MODE_SHADE>)
    const uint4* fast_blob;     // fp32 brute-force layout (FastScene), null otherwise
    uint32_t fast_words;
    uint32_t fast_k;            // k_bounce_fast, passes from depth 2 on: bounces a ray makes in registers per pass;
                                // 0 = chosen on the device from the measured survival rate (fast_schedule)
    uint32_t pass, n_passes;    // k_bounce_fast: index of this launch in its batch (0 and 1 are depth 0 and 1), launches per batch
    uint32_t* fast_hint;        // k_bounce_fast: bounces per pass the previous batch settled on (device word, 0 = none yet)
    FastHeader fast_hd;
    const WideNode* wide;       // k_extend_cw: the 8-wide quantised tree (ipt_wide.h), null = 2-wide traversal (k_extend_bvh)
    uint2* wide_spill;          // k_extend_cw: stack entries beyond CW_STACK, wide_spill_cap per resident lane
    uint32_t wide_spill_cap;
    uint32_t leaf_min;          // k_extend_cw: lanes with a primitive to test that make a primitive step worth a warp instruction
    const uint2* grid_cells;    // k_extend_grid: {first reference, count} per cell, null = walk a tree
    const uint32_t* grid_refs;  // k_extend_grid: slots, cell by cell
    GridHeader grid;
    uint32_t static_slices;     // k_bounce_fast: 1 = slices assigned to warps statically instead of claimed (A/B runs)
};

// Deterministic accumulation: a contribution is rounded once to a multiple of 1/fixed_scale and added with an
// integer atomic, so the frame is bit-identical for any scheduling, batch size, tile size or GPU count.
template <typename R>
__device__ __forceinline__ void accumulate(const KParams<R>& p, uint32_t pixel, V3<R> v)
{
    unsigned long long* f = p.frame + 3ull * pixel;
    if (p.flags & 0x4u) {   // IPT_FLAG_FLOAT_ACCUM
        atomicAdd(reinterpret_cast<double*>(f), (double)v.x);
        atomicAdd(reinterpret_cast<double*>(f) + 1, (double)v.y);
        atomicAdd(reinterpret_cast<double*>(f) + 2, (double)v.z);
    } else {
        atomicAdd(f, (unsigned long long)__double2ll_rn((double)v.x * p.fixed_scale));
        atomicAdd(f + 1, (unsigned long long)__double2ll_rn((double)v.y * p.fixed_scale));
        atomicAdd(f + 2, (unsigned long long)__double2ll_rn((double)v.z * p.fixed_scale));
    }
}

// Sample id -> (pixel, sample).  32 consecutive ids are one 8x4 pixel block at one sample index (a warp's camera rays are
// neighbours); consecutive groups walk the MICRO-TILES of the batch, then the samples.  (Round 1 walked the samples of a
// block first: the ~3500 warps in flight then worked on 3-4 blocks, and the emission they found - three 64-bit atomics per
// hit - went to ~330 addresses; pass 0 of the batches that only see the light ran at 19 % issue utilisation,
// profiles/r02_ncu_pass0_v2.txt.)
template <typename R>
__device__ __forceinline__ bool decode_sample(const KParams<R>& p, uint32_t sid, uint32_t& px, uint32_t& pz, uint32_t& sample)
{
    const uint32_t g = sid >> 5, l = sid & 31u;
    const uint32_t q = g / p.mt_count;
    sample = p.base_sample + q;
    const uint32_t mt = p.base_mt + (g - q * p.mt_count);
    if (mt >= p.n_mt || sample >= p.spp) return false;
    const uint32_t xy = p.mt_list[mt];
    px = (xy & 0xFFFFu) * 8u + (l & 7u);
    pz = (xy >> 16) * 4u + (l >> 3);
    return px < p.W && pz < p.H;
}

// Renderer.cu:112-147: per-pixel gaze (not jittered), +-1 pixel box jitter on the origin only.
template <typename R>
__device__ __forceinline__ void camera_ray(const KParams<R>& p, uint32_t px, uint32_t pz, uint32_t sample, Ray<R>& r)
{
    const R corr = (p.W % 2 == 0) ? (R)0.5 : (R)0;   // :118-119 — the z axis also tests the WIDTH's parity
    const R stepX = (px < p.W / 2) ? (R)(p.W / 2 - px) - corr : ((R)p.W / (R)2 - (R)px - (R)1) + ((corr == (R)0) ? (R)1 : corr);
    const R stepZ = (pz < p.H / 2) ? (R)(p.H / 2 - pz) - corr : ((R)p.H / (R)2 - (R)pz - (R)1) + ((corr == (R)0) ? (R)1 : corr);
    r.d = normalize(p.camD + p.camX * stepX * p.fov + p.camZ * stepZ * p.fov);                       // :127
    const uint32_t pixel = pz * p.W + px;
    const uint4 rnd = philox4x32(pixel, sample, NODE_CAMERA, CTR_TAG, p.keys);
    R jx = s24<R>(rnd.x), jz = s24<R>(rnd.y);                                                         // :133-134
    if (p.flags & 0x10u) {   // IPT_FLAG_STRATIFIED (extension): jitter of sample i drawn inside stratum i of an n x n grid
        const uint32_t n = p.strat_n;
        if (sample < n * n) {
            const R inv = (R)1 / (R)n;
            jx = ((R)(sample % n) + (jx + (R)1) * (R)0.5) * inv * (R)2 - (R)1;
            jz = ((R)(sample / n) + (jz + (R)1) * (R)0.5) * inv * (R)2 - (R)1;
        }
    }
    const V3<R> tent = p.camX * jx + p.camZ * jz;                                                     // :135
    const V3<R> origin = p.camO + p.camX * stepX + p.camZ * stepZ + tent;                             // :138
    r.o = origin + p.camD * (R)IPT_VIEWPORT_DISTANCE;                                                 // :139
    r.thr = mk<R>(1, 1, 1);
    r.pixel = pixel;
    r.meta = make_meta(0, 0, false, false, sample);
    r.self = NO_OBJECT;
}

// Exact pruning of camera rays (the counterpart of SURVEY.md App. A.8 for pass 0): a pixel whose UN-jittered camera ray
// misses the scene's bounding box grown by `grow` (>= the +-1 pixel box jitter of Renderer.cu:133-138 plus rounding
// slack) is exactly 0 for every sample — all its rays miss every object (Renderer.cu:153).  One thread per micro-tile
// (8x4 pixels) of this rank, in tile-major order: flag = 1 iff any of its pixels can see the grown box.  fp64.
struct ActiveParams {
    double camO[3], camD[3], camX[3], camZ[3], fov, lo[3], hi[3];
    uint32_t W, H;
    const uint32_t* tile_ids;
    uint32_t n_tiles_local, tiles_x, tile_w, tile_h, mt_x, mt_per_tile;
    uint32_t* packed;   // out: (y << 16 | x) of the micro-tile
    uint8_t* flags;     // out
};
__global__ void __launch_bounds__(256) k_active_microtiles(const __grid_constant__ ActiveParams p)
{
    const uint32_t total = p.n_tiles_local * p.mt_per_tile;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t lt = i / p.mt_per_tile, m = i % p.mt_per_tile;
        const uint32_t tile = p.tile_ids[lt];
        const uint32_t mx = (tile % p.tiles_x) * p.mt_x + m % p.mt_x, my = (tile / p.tiles_x) * (p.tile_h / 4) + m / p.mt_x;
        p.packed[i] = (my << 16) | mx;
        bool any = false;
        for (uint32_t l = 0; l < 32 && !any; l++) {
            const uint32_t px = mx * 8 + (l & 7), pz = my * 4 + (l >> 3);
            if (px >= p.W || pz >= p.H) continue;
            const double corr = (p.W % 2 == 0) ? 0.5 : 0.0;
            const double stepX = (px < p.W / 2) ? (double)(p.W / 2 - px) - corr : ((double)p.W / 2 - px - 1.0) + ((corr == 0.0) ? 1.0 : corr);
            const double stepZ = (pz < p.H / 2) ? (double)(p.H / 2 - pz) - corr : ((double)p.H / 2 - pz - 1.0) + ((corr == 0.0) ? 1.0 : corr);
            double o[3], d[3];   // the slab test does not need a unit direction
            for (int k = 0; k < 3; k++) {
                d[k] = p.camD[k] + p.camX[k] * stepX * p.fov + p.camZ[k] * stepZ * p.fov;
                o[k] = p.camO[k] + p.camX[k] * stepX + p.camZ[k] * stepZ + p.camD[k] * IPT_VIEWPORT_DISTANCE;
            }
            double tn = 0.0, tf = 1e300;
            bool miss = false;
            for (int k = 0; k < 3; k++) {
                if (d[k] == 0.0) { if (o[k] < p.lo[k] || o[k] > p.hi[k]) miss = true; continue; }
                const double a = (p.lo[k] - o[k]) / d[k], b = (p.hi[k] - o[k]) / d[k];
                tn = fmax(tn, fmin(a, b)); tf = fmin(tf, fmax(a, b));
            }
            any = !miss && tn <= tf;
        }
        p.flags[i] = any ? 1 : 0;
    }
}

// Copies `n16` 16-byte words from global to shared memory with the whole CTA (128-bit, coalesced).
__device__ __forceinline__ void stage(uint4* dst, const uint4* src, uint32_t n16)
{
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
}

// Nearest hit for the generic kernels: brute-force slots, or the BVH (fp32: typed leaf records when present).
template <typename R, int MODE>
__device__ __forceinline__ Hit<R> nearest_any(const SceneView<R>& sc, const float4* top, uint32_t n_top, const V3<R> o, const V3<R> d,
                                              uint32_t self, bool onSurf)
{
    if (MODE == MODE_BRUTE) return nearest_brute<R>(sc, o, d, self, onSurf);
    return nearest_bvh<R>(sc, top, n_top, o, d, self, onSurf);
}
template <>
__device__ __forceinline__ Hit<float> nearest_any<float, MODE_BVH>(const SceneView<float>& sc, const float4* top, uint32_t n_top,
                                                                   const V3<float> o, const V3<float> d, uint32_t self, bool onSurf)
{
    if (sc.bslot) return nearest_bvh_f32(sc, top, n_top, o, d, self, onSurf);
    return nearest_bvh<float>(sc, top, n_top, o, d, self, onSurf);
}

// One bounce of one batch.  FIRST: rays are generated from sample ids instead of read from qin.
//
// DEFER (maxDepth >= 130 only): the reference folds a deep path's radiance with an int8_t index starting at
// depth_end - 2 (Renderer.cu:216), so a deepLayers() call that ends at depth_end >= 130 — by reaching maxDepth, or by
// missing the scene at depth >= 130 — returns 0 as a whole, emission of its earlier hits included.  To reproduce that,
// the emission a path collects from depth 2 on travels with the ray (extra queue planes) and reaches the frame only
// when the path misses the scene at depth <= 129; a path with deferred radiance is traced to its end even when its
// throughput has dropped to zero.
template <typename R, int MODE, bool FIRST, bool DEFER>
__global__ void __launch_bounds__(BLOCK_THREADS) k_bounce(const __grid_constant__ KParams<R> p)
{
    extern __shared__ uint4 smem[];
    SceneView<R> sc = p.sc;
    const float4* top = nullptr;
    uint32_t n_top = 0;
    if (MODE == MODE_SHADE) {
        // nothing to stage: geometry and materials of the hit slot are read from global memory
    } else if (MODE == MODE_BRUTE) {
        // whole scene -> shared memory: geometry slots, slot->object ids, materials
        constexpr uint32_t W16 = sizeof(R4<R>) / 16;
        const uint32_t ng = sc.n_slots * 4 * W16, nm = sc.n_objects * 2 * W16, ni = (sc.n_slots + 3) / 4;
        stage(smem, reinterpret_cast<const uint4*>(p.sc.geom), ng);
        stage(smem + ng, reinterpret_cast<const uint4*>(p.sc.mat), nm);
        stage(smem + ng + nm, reinterpret_cast<const uint4*>(p.sc.slot_obj), ni);
        sc.geom = reinterpret_cast<const R4<R>*>(smem);
        sc.mat = reinterpret_cast<const R4<R>*>(smem + ng);
        sc.slot_obj = reinterpret_cast<const uint32_t*>(smem + ng + nm);
    } else {
        n_top = sc.n_nodes < (uint32_t)BVH_TOP_NODES ? sc.n_nodes : (uint32_t)BVH_TOP_NODES;
        stage(smem, reinterpret_cast<const uint4*>(p.sc.nodes), n_top * 4);
        top = reinterpret_cast<const float4*>(smem);
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t n_in = FIRST ? p.n_first : p.counters[CNT + p.depth];
    const uint32_t depth = p.depth;
    uint32_t* work = p.counters + WORK + depth;
    uint32_t* out_count = p.counters + CNT + depth + 1;
    unsigned long long my_traced = 0;       // warp-uniform: casts of the rays this warp processed
    uint32_t my_shadow = 0;                 // per lane: visibility casts of the next-event extension

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(work, (uint32_t)GRAB);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_in) break;
#pragma unroll 1
        for (uint32_t k = 0; k < GRAB; k += 32) {
            const uint32_t i = base + k + lane;
            bool live = i < n_in;
            Ray<R> r;
            if (FIRST) {
                uint32_t px = 0, pz = 0, sample = 0;
                live = live && decode_sample(p, i, px, pz, sample);
                if (live) camera_ray(p, px, pz, sample, r);
            } else if (live) {
                q_load(p.qin, i, r);
            }
            bool has0 = false, has1 = false;
            Ray<R> o0, o1;
            V3<R> acc = mk<R>(0, 0, 0);
            if (DEFER && !FIRST && live) q_load_acc(p.qin, i, acc);
            if (live) {
                const bool onSurf = (r.meta & META_ONSURF) != 0;
                const bool deep = DEFER && depth >= 2 && !(r.meta & META_PROBE);
                Hit<R> h;
                if (MODE == MODE_SHADE) {
                    const uint2 hv = p.hits[i];
                    h.t = (R)__uint_as_float(hv.x); h.slot = hv.y;
                    h.obj = hv.y == NO_OBJECT ? NO_OBJECT : __ldg(sc.slot_obj + hv.y);
                } else {
                    h = nearest_any<R, MODE>(sc, top, n_top, r.o, r.d, r.self, onSurf);
                }
                if (h.slot != NO_OBJECT) {
                    const bool isRect = (h.obj & RECT_BIT) != 0;
                    const uint32_t obj = h.obj & ~RECT_BIT;
                    const R4<R> m0 = sc.mat[2 * obj], m1 = sc.mat[2 * obj + 1];
                    bool count_emission = m1.w != (R)0;   // E of every hit counts (Renderer.cu:170,193,211)
                    if (count_emission && (r.meta & META_NEE) && !isRect) {
                        // next-event extension: the hit this ray left sampled the emissive spheres it lies outside of
                        // explicitly; their emission must not be counted a second time by the ray that found them by chance
                        const R4<R> g = sc.geom[4 * (size_t)h.slot];
                        const V3<R> oc = r.o - xyz(g);
                        const R rr1 = g.w * ((R)1 + (R)1e-4);
                        count_emission = !(dot(oc, oc) > rr1 * rr1);
                    }
                    if (count_emission) {
                        if (deep) acc = acc + mul(r.thr, xyz(m1));
                        else accumulate(p, r.pixel, mul(r.thr, xyz(m1)));
                    }
                    const bool probe = (r.meta & META_PROBE) != 0;
                    // Continuation exists iff another hit would still be evaluated: depth+1 < maxDepth (:161,:184,:201).
                    // Paths whose throughput is exactly 0 and the second branch of a depth-0 split after its first hit
                    // contribute exactly 0 and are not traced.
                    if (!probe && depth + 1 < p.maxDepth) {
                        V3<R> nthr = mul(r.thr, xyz(m0));
                        const bool pending = deep && (acc.x != (R)0 || acc.y != (R)0 || acc.z != (R)0);
                        if (nthr.x != (R)0 || nthr.y != (R)0 || nthr.z != (R)0 || pending) {
                            const uint32_t lane_id = (r.meta >> 8) & 3u, sample = (r.meta >> 12) & 0xFFFFu;
                            const V3<R> P = r.o + r.d * h.t;                                       // :156,:179,:207
                            const uint4 rnd = philox4x32(r.pixel, sample, (lane_id << 8) | depth, CTR_TAG, p.keys);
                            const Spawn<R> sp = scatter<R>(isRect, sc.geom[4 * (size_t)h.slot], (int)m0.w, P, r.d, depth, rnd);
                            bool alive = sp.has0;
                            const bool onS = isRect || fabs(dot(r.d, r.d) - (R)1) < (R)1e-3;
                            bool nee = false;
                            if constexpr (MODE != MODE_SHADE)
                            if ((p.flags & 0x20u) && sc.n_lights && (int)m0.w == 0 && alive) {
                                // IPT_FLAG_NEXT_EVENT (extension, off by default): at a diffuse hit one emissive sphere is
                                // chosen uniformly and a direction uniformly inside the cone it subtends; the reference's
                                // "diffuse" lobe (AObject.hpp:35-45: cube-normalised, flipped into the hemisphere of N,
                                // weight = colour) has density 1 / (12 max|w_i|^3) there, which weights the sample.
                                nee = true;
                                const uint4 rl = philox4x32(r.pixel, sample, (lane_id << 8) | depth, CTR_TAG + 2u, p.keys);
                                const uint32_t li = min((uint32_t)(u23<R>(rl.x) * (R)sc.n_lights), sc.n_lights - 1u);
                                const double* L = sc.lights + 8 * (size_t)li;
                                const V3<R> lc = mk<R>((R)__ldg(L), (R)__ldg(L + 1), (R)__ldg(L + 2));
                                const R lr = (R)__ldg(L + 3);
                                const V3<R> w = lc - P;
                                const R dist2 = dot(w, w), rr1 = lr * ((R)1 + (R)1e-4);
                                if (dist2 > rr1 * rr1) {
                                    const V3<R> g0n = xyz(sc.geom[4 * (size_t)h.slot]);
                                    V3<R> N;                                                      // A.3 conventions, as in scatter()
                                    if (isRect) N = dot(r.d, g0n) < (R)0 ? g0n : -g0n;
                                    else { const V3<R> raw = normalize(P - g0n); N = dot(r.d, raw) < (R)0 ? -raw : raw; }
                                    const R sin2 = lr * lr / dist2, cosmax = sqrt(fmax((R)0, (R)1 - sin2));
                                    const R omc = sin2 / ((R)1 + cosmax);                          // 1 - cosmax without cancellation
                                    const R omt = u23<R>(rl.y) * omc, ct = (R)1 - omt, st = sqrt(omt * ((R)1 + ct));
                                    const R phi = (R)6.283185307179586 * u23<R>(rl.z);
                                    const V3<R> a = w * ((R)1 / sqrt(dist2));
                                    const V3<R> up = fabs(a.x) < (R)0.57 ? mk<R>(1, 0, 0) : mk<R>(0, 1, 0);
                                    const V3<R> t1 = normalize(cross(a, up)), t2 = cross(a, t1);
                                    const V3<R> wl = a * ct + (t1 * cos(phi) + t2 * sin(phi)) * st;
                                    if (dot(wl, N) > (R)0) {
                                        const R m = fmax(fabs(wl.x), fmax(fabs(wl.y), fabs(wl.z)));
                                        const R lobe = (R)1 / ((R)12 * m * m * m);
                                        const R inv_q = (R)6.283185307179586 * omc;               // 1 / cone density
                                        my_shadow++;
                                        const Hit<R> hs = nearest_any<R, MODE>(sc, top, n_top, P, wl, h.obj, onS);
                                        if (hs.slot != NO_OBJECT && hs.obj == (uint32_t)__ldg(L + 7)) {
                                            const V3<R> Le = mk<R>((R)__ldg(L + 4), (R)__ldg(L + 5), (R)__ldg(L + 6));
                                            const V3<R> c = mul(nthr, Le) * (lobe * inv_q * (R)sc.n_lights);
                                            if (deep) acc = acc + c;
                                            else accumulate(p, r.pixel, c);
                                        }
                                    }
                                }
                            }
                            if ((p.flags & 0x8u) && depth >= 3 && alive) {   // IPT_FLAG_RUSSIAN_ROULETTE (extension, off by default)
                                const R q = fmin((R)1, fmax((R)0.05, fmax(nthr.x, fmax(nthr.y, nthr.z))));
                                const uint4 rr = philox4x32(r.pixel, sample, (lane_id << 8) | depth, CTR_TAG + 1u, p.keys);
                                if (u23<R>(rr.x) >= q) alive = false;
                                else nthr = nthr * ((R)1 / q);
                            }
                            if (alive) {
                                has0 = true;
                                o0.o = P; o0.d = sp.d0; o0.thr = nthr * sp.w0; o0.pixel = r.pixel; o0.self = h.obj;
                                o0.meta = make_meta(depth + 1, lane_id, false, onS, sample) | (nee ? META_NEE : 0u);
                                if (sp.teleport) { o0.o = mk<R>(0, 0, 0); o0.self = NO_OBJECT; o0.meta = make_meta(depth + 1, lane_id, false, false, sample); }
                            }
                            if (sp.has1) {
                                // depth 0: the second ray is an emission probe of the next surface (its deeper recursion
                                // folds to 0 in the reference, Renderer.cu:173,216); depth 1: a full second path (lane 1)
                                has1 = true;
                                o1.o = P; o1.d = sp.d1; o1.thr = nthr * sp.w1; o1.pixel = r.pixel; o1.self = h.obj;
                                o1.meta = make_meta(depth + 1, depth == 0 ? 2u : 1u, depth == 0, onS, sample);
                            }
                        }
                    }
                } else if (deep && depth <= 129) {
                    // the path leaves the scene at depth <= 129: its fold index fits an int8_t, its radiance counts
                    if (acc.x != (R)0 || acc.y != (R)0 || acc.z != (R)0) accumulate(p, r.pixel, acc);
                }
            }
            // ---- compaction: one atomic per warp
            const uint32_t m_live = __ballot_sync(0xffffffffu, live);
            const uint32_t m0b = __ballot_sync(0xffffffffu, has0), m1b = __ballot_sync(0xffffffffu, has1);
            if (MODE != MODE_SHADE) my_traced += __popc(m_live);   // the split pipeline counts casts in k_extend_bvh
            const uint32_t c0 = __popc(m0b), tot = c0 + __popc(m1b);
            if (tot) {
                uint32_t ob = 0;
                if (lane == 0) ob = atomicAdd(out_count, tot);
                ob = __shfl_sync(0xffffffffu, ob, 0);
                if (has0) q_store(p.qout, ob + __popc(m0b & lt_mask), o0);
                if (has1) q_store(p.qout, ob + c0 + __popc(m1b & lt_mask), o1);
                if (DEFER) {
                    // the main continuation inherits the deferred radiance; a ray spawned by a split starts with none
                    if (has0) q_store_acc(p.qout, ob + __popc(m0b & lt_mask), acc);   // non-zero only on deep paths
                    if (has1) q_store_acc(p.qout, ob + c0 + __popc(m1b & lt_mask), mk<R>(0, 0, 0));
                }
            }
        }
    }
    my_traced += __reduce_add_sync(0xffffffffu, my_shadow);
    if (lane == 0 && my_traced) atomicAdd(p.traced, my_traced);
}

// Work counters of the traversal kernels (ipt_stats.node_steps ...): warp sums, one atomic per counter and warp.
// stats[] = {casts, queue records, node steps, box tests, leaf steps, sphere tests, rectangle tests}
__device__ __forceinline__ void add_work(unsigned long long* stats, uint32_t nodes, uint32_t boxes, uint32_t leaves, uint32_t sph, uint32_t rect)
{
    nodes = __reduce_add_sync(0xffffffffu, nodes); boxes = __reduce_add_sync(0xffffffffu, boxes); leaves = __reduce_add_sync(0xffffffffu, leaves);
    sph = __reduce_add_sync(0xffffffffu, sph); rect = __reduce_add_sync(0xffffffffu, rect);
    if ((threadIdx.x & 31u) == 0) {
        atomicAdd(stats + 2, (unsigned long long)nodes); atomicAdd(stats + 3, (unsigned long long)boxes); atomicAdd(stats + 4, (unsigned long long)leaves);
        atomicAdd(stats + 5, (unsigned long long)sph); atomicAdd(stats + 6, (unsigned long long)rect);
    }
}

// ---------------------------------------------------------------------------------------------- split pipeline (BVH, fp32)
// Stage 1 of the split wavefront step used for BVH scenes: camera rays -> queue (compacted, one atomic per warp).
template <typename R>
__global__ void __launch_bounds__(BLOCK_THREADS) k_raygen(const __grid_constant__ KParams<R> p)
{
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    uint32_t* out_count = p.counters + CNT;   // rays queued for pass 0
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < p.n_first; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        uint32_t px = 0, pz = 0, sample = 0;
        const bool live = i < p.n_first && decode_sample(p, i, px, pz, sample);
        Ray<R> r;
        if (live) camera_ray(p, px, pz, sample, r);
        const uint32_t m = __ballot_sync(0xffffffffu, live);
        if (m) {
            uint32_t ob = 0;
            if (lane == 0) ob = atomicAdd(out_count, (uint32_t)__popc(m));
            ob = __shfl_sync(0xffffffffu, ob, 0);
            if (live) q_store(p.qout, ob + __popc(m & lt_mask), r);
        }
    }
}

// Stage 2: nearest hit through the BVH with persistent threads and LANE-level refill.  Traversal lengths differ by an
// order of magnitude between rays of one warp (the fused kernel ran at 5.5 of 32 lanes active on the 1M-primitive
// scene, profiles/r01_ncu_bvh_v1.txt); here a lane that has finished its ray writes the hit and, as soon as
// `refill_min` lanes of the warp are idle, the idle lanes claim new rays with one warp-aggregated atomic.

__global__ void __launch_bounds__(BLOCK_THREADS, 5) k_extend_bvh(const __grid_constant__ KParams<float> p)
{
    extern __shared__ uint4 smem[];
    const SceneView<float> sc = p.sc;
    const uint32_t n_top = sc.n_nodes < (uint32_t)BVH_TOP_NODES ? sc.n_nodes : (uint32_t)BVH_TOP_NODES;
    stage(smem, reinterpret_cast<const uint4*>(sc.nodes), n_top * 4);
    const float4* top = reinterpret_cast<const float4*>(smem);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const uint32_t n_in = p.counters[CNT + p.depth];
    uint32_t* work = p.counters + WORK_EXTEND + p.depth;
    unsigned long long my_traced = 0;
    uint32_t w_nodes = 0, w_leaves = 0, w_prims = 0, w_sph = 0;   // work counters of ipt_stats (per lane)

    int stack[64];
    int sp = 0, node = 0;
    bool has = false, exhausted = false;
    uint32_t idx = 0, self = NO_OBJECT;
    bool onSurf = false;
    V3<float> o = mk<float>(0, 0, 0), d = o, inv = o, bi = o, oi = o;
    Hit<float> best;
    best.t = (float)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;
    const float slack = 1.0000004f, tiny = 1e-18f;

    for (;;) {
        const uint32_t idle = __ballot_sync(0xffffffffu, !has);
        if (!exhausted && (idle == 0xffffffffu || (uint32_t)__popc(idle) >= p.refill_min)) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(work, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            exhausted = base + (uint32_t)__popc(idle) >= n_in;
            if (!has) {
                idx = base + __popc(idle & lt_mask);
                if (idx < n_in) {
                    // streaming loads / stores for the ray and hit records: read once, they should not displace BVH nodes in L1
                    const uint4 a = __ldcs(p.qin.base + idx), b = __ldcs(p.qin.base + p.qin.capacity + idx), c = __ldcs(p.qin.base + 2u * p.qin.capacity + idx);
                    o = mk<float>(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
                    d = mk<float>(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
                    onSurf = (c.z & META_ONSURF) != 0; self = c.w;
                    inv = mk<float>(1.f / d.x, 1.f / d.y, 1.f / d.z);
                    bi = mk<float>(1.f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)), 1.f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
                                   1.f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
                    oi = mk<float>(o.x * bi.x, o.y * bi.y, o.z * bi.z);
                    best.t = (float)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;
                    sp = 0; node = 0; has = true;
                    my_traced++;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, has) == 0) break;
        // Both phases are warp-synchronous loops driven by __any_sync: the warp reconverges before the leaf phase
        // (left to the compiler, lanes that found their leaf early ran the primitive tests one or two at a time).
        bool done = false;
        // ---- phase 1: lanes with a ray descend inner nodes.  The phase ends when fewer than `descend_min` lanes are
        // still descending and at least one lane holds a leaf to test (the stragglers keep their state and resume in
        // the next round) — waiting for the slowest lane left this loop at 8.9 of 32 lanes active.
        for (;;) {
            const uint32_t desc = __ballot_sync(0xffffffffu, has && !done && node >= 0);
            if (desc == 0) break;
            if ((uint32_t)__popc(desc) < p.descend_min && __any_sync(0xffffffffu, has && !done && node < 0)) break;
            if (has && !done && node >= 0) {
                float4 a, b, c, e;
                if ((uint32_t)node < n_top) { const float4* q = top + 4 * node; a = q[0]; b = q[1]; c = q[2]; e = q[3]; }
                else { const float4* q = sc.nodes + 4 * (size_t)node; ldg256(q, a, b); ldg256(q + 2, c, e); }
                float t0x = fmaf(a.x, bi.x, -oi.x), t1x = fmaf(a.w, bi.x, -oi.x);
                float t0y = fmaf(a.y, bi.y, -oi.y), t1y = fmaf(b.x, bi.y, -oi.y);
                float t0z = fmaf(a.z, bi.z, -oi.z), t1z = fmaf(b.y, bi.z, -oi.z);
                const float n0 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.f));
                const float f0 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best.t)) * slack;
                t0x = fmaf(b.z, bi.x, -oi.x); t1x = fmaf(c.y, bi.x, -oi.x);
                t0y = fmaf(b.w, bi.y, -oi.y); t1y = fmaf(c.z, bi.y, -oi.y);
                t0z = fmaf(c.x, bi.z, -oi.z); t1z = fmaf(c.w, bi.z, -oi.z);
                const float n1 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.f));
                const float f1 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best.t)) * slack;
                const bool h0 = n0 <= f0, h1 = n1 <= f1;
                const int c0 = __float_as_int(e.x), c1 = __float_as_int(e.y);
                w_nodes++;
                if (h0 && h1) {
                    const bool swap = n1 < n0;
                    stack[sp++] = swap ? c0 : c1;
                    node = swap ? c1 : c0;
                } else if (h0 || h1) {
                    node = h0 ? c0 : c1;
                } else if (sp == 0) {
                    done = true;
                } else {
                    node = stack[--sp];
                }
            }
        }
        // ---- phase 2: the lanes holding a leaf test its primitives, one primitive per lane per iteration
        const bool at_leaf = has && !done && node < 0;
        uint32_t s_cur = 0, s_end = 0;
        if (at_leaf) {
            const uint32_t code = (uint32_t)(~node);
            s_cur = code >> 4; s_end = s_cur + (code & 15u) + 1u;
            w_leaves++; w_prims += s_end - s_cur;
        }
        while (__any_sync(0xffffffffu, s_cur < s_end)) {
            if (s_cur < s_end) { test_bslot(sc, s_cur, o, d, inv, self, onSurf, best, w_sph); s_cur++; }
        }
        if (at_leaf) {
            if (sp == 0) done = true;
            else node = stack[--sp];
        }
        if (has && done) {
            __stcs(p.hits + idx, make_uint2(__float_as_uint(best.t), best.slot));
            has = false;
        }
    }
    if (my_traced) atomicAdd(p.traced, my_traced);
    add_work(p.traced, w_nodes, 2u * w_nodes, w_leaves, w_sph, w_prims - w_sph);
}

// ---------------------------------------------------------------------------------------------- grid traversal
// Stage 2 of the split pipeline over the uniform grid (ipt_scene::grid_*, host/grid.cpp): the nearest hit of
// Renderer.cu:227-243 for scenes of many small, evenly spread primitives, found by walking the cells the ray passes through
// (3D-DDA) instead of a hierarchy - on BASELINE config 5 about 13 cell steps and 17 primitive tests per ray against 49 inner
// nodes (98 box tests) and 18 primitive tests through the 2-wide tree.  One ray per lane, persistent warps with lane-level
// refill as in k_extend_bvh; primitives are tested by the same typed records and arithmetic (test_bslot) with the reference's
// tie rule, so frames are bit-identical with the tree traversals.
//  * a new ray first tests the "big" primitives (walls, large lights: up to 64 slots kept in the kernel parameters), then is
//    clipped against the grid's box;
//  * per cell: one 8-byte load {first reference, count}; the references are slots, tested one per lane and iteration;
//  * the loads are taken off the critical path.  The capture of the first version (profiles/r02_ncu_grid_final.txt and its source
//    page) had a third of all stall samples on three dependent loads inside one step: the cell entry (8.6 %), the reference
//    (11.2 %) and the record it points at (14.0 %).  Now the walk runs one cell ahead of the cell being opened - a cell step reads
//    the entry requested by the step before it and requests the entry of the cell after (speculatively: a ray that ends in this
//    cell has asked for one entry too many) - and the reference a primitive step needs was requested when the cell was opened or
//    by the primitive step before it; what is left exposed is the record load (profiles/r02_ncu_grid_ahead.txt);
//  * a ray ends when the exit distance of the cell it has just finished is not below its nearest hit (a primitive that
//    overlaps several cells may report a hit beyond the current cell: it stays a candidate and the walk goes on), or when
//    it leaves the grid;
//  * the next boundary along an axis is evaluated from the integer cell coordinate, t = i * (cell / d) + (lo - o) / d, not
//    accumulated: no drift over hundreds of steps; what is left (~1e-4 at |x| ~ 1e3) is covered 40 times by the padding of
//    the boxes the host filed the primitives under;
//  * cell steps and primitive tests are warp-synchronous blocks, each run when enough lanes want it (GRID_CELL_MIN,
//    GRID_PRIM_MIN) or nobody wants the other.
// (The first version ran bursts of up to 4 steps of one kind before the warp looked at its finished and idle lanes again: 2.6 ->
// 3.2 Gbounces/s then; since steps run only when enough lanes want them, a burst's second trip mostly found too few and left again:
// 1 / 2 / 4 steps gave 3.88 / 3.63 / 3.64 Gbounces/s, and the bursts went.)
#ifndef IPT_GRID_CTAS
#define IPT_GRID_CTAS 4
#endif
// Lanes that make a refill / a cell step / a primitive step worth running (measured on config 5, profiles/README.md: 12 / 12 / 16).
// Compile-time constants (-DIPT_GRID_REFILL_MIN=... for A/B builds): as kernel parameters each test in the loop was a constant-bank
// load and a register compare instead of a compare with an immediate.
#ifndef IPT_GRID_REFILL_MIN
#define IPT_GRID_REFILL_MIN 12
#define IPT_GRID_CELL_MIN 12
#define IPT_GRID_PRIM_MIN 16
#endif
static constexpr uint32_t GRID_REFILL_MIN = IPT_GRID_REFILL_MIN, GRID_CELL_MIN = IPT_GRID_CELL_MIN, GRID_PRIM_MIN = IPT_GRID_PRIM_MIN;
__global__ void __launch_bounds__(BLOCK_THREADS, IPT_GRID_CTAS) k_extend_grid(const __grid_constant__ KParams<float> p)
{
    const SceneView<float> sc = p.sc;
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const uint32_t n_in = p.counters[CNT + p.depth];
    uint32_t* work = p.counters + WORK_EXTEND + p.depth;
    const float tiny = 1e-18f;
    const uint32_t rx = p.grid.res[0], rxy = p.grid.res[0] * p.grid.res[1];

    bool exhausted = false;
    uint32_t my_traced = 0;
    uint32_t w_cells = 0, w_prims = 0, w_sph = 0, w_occupied = 0;        // work counters of ipt_stats
    bool has = false, marching = false;                                   // marching: a requested cell entry is waiting to be opened
    bool ahead_in = false;                                                // the cell after it lies inside the grid
    uint32_t idx = 0, self = NO_OBJECT;
    bool onSurf = false;
    V3<float> o = mk<float>(0, 0, 0), d = o, inv = o;
    V3<float> fi = o, tmx = o, A = o, B = o;                              // the walk (one cell ahead): cell coordinate, next boundary per axis, t = fi * A + B
    uint32_t cell = 0;
    uint2 ce_nx = make_uint2(0u, 0u);                                     // entry of the cell to open next (in flight)
    float tex_nx = 0.f;                                                   // where the ray leaves that cell
    uint32_t slot_nx = 0;                                                 // the reference at pr_cur (in flight since the cell was opened / the primitive step before)
    uint32_t pr_cur = 0, pr_end = 0;                                      // references of the current cell still to test
    float t_exit = 0.f;                                                   // where the ray leaves the current cell
    Hit<float> best;
    best.t = (float)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;

    // The walk's cell becomes the one to open next - its entry is requested now, read by the next cell step - and the walk moves on.
    auto advance = [&]() {
        ce_nx = __ldg(p.grid_cells + cell);
        tex_nx = fminf(tmx.x, fminf(tmx.y, tmx.z));
        const bool sx = tmx.x <= tmx.y && tmx.x <= tmx.z, sy = !sx && tmx.y <= tmx.z;
        const float a_k = sx ? A.x : (sy ? A.y : A.z);                      // its sign is the direction of travel along the axis
        const float stepf = a_k > 0.f ? 1.f : -1.f;
        const uint32_t stride = sx ? 1u : (sy ? rx : rxy);
        const float f = (sx ? fi.x : (sy ? fi.y : fi.z)) + stepf;
        const float lim = (float)(sx ? p.grid.res[0] : (sy ? p.grid.res[1] : p.grid.res[2]));
        ahead_in = f >= 0.f && f < lim;
        const float nt = fmaf(f, a_k, sx ? B.x : (sy ? B.y : B.z));
        fi.x = sx ? f : fi.x; fi.y = sy ? f : fi.y; fi.z = (!sx && !sy) ? f : fi.z;
        tmx.x = sx ? nt : tmx.x; tmx.y = sy ? nt : tmx.y; tmx.z = (!sx && !sy) ? nt : tmx.z;
        cell = a_k > 0.f ? cell + stride : cell - stride;
    };

    for (;;) {
        if (has && !marching && pr_cur >= pr_end) {                       // finished (or never entered the grid)
            __stcs(p.hits + idx, make_uint2(__float_as_uint(best.t), best.slot));
            has = false;
        }
        // ---- refill
        const uint32_t idle = __ballot_sync(0xffffffffu, !has);
        if (!exhausted && (idle == 0xffffffffu || (uint32_t)__popc(idle) >= GRID_REFILL_MIN)) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(work, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            exhausted = base + (uint32_t)__popc(idle) >= n_in;
            if (!has) {
                idx = base + __popc(idle & lt_mask);
                if (idx < n_in) {
                    const uint4 a = __ldcs(p.qin.base + idx), b = __ldcs(p.qin.base + p.qin.capacity + idx), c = __ldcs(p.qin.base + 2u * p.qin.capacity + idx);
                    o = mk<float>(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
                    d = mk<float>(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
                    onSurf = (c.z & META_ONSURF) != 0; self = c.w;
                    inv = mk<float>(1.f / d.x, 1.f / d.y, 1.f / d.z);       // primitive tests: +-inf for zero components (Plane.cu:55)
                    best.t = (float)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;
                    has = true; marching = false; pr_cur = pr_end = 0;
                    my_traced++;
                    // the grid's box: components below 1e-18 become +-1e-18, so no inf - inf arises (as the tree traversals do)
                    const V3<float> bi = mk<float>(rcp_fast(fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)), rcp_fast(fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
                                                   rcp_fast(fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
                    const float ax = (p.grid.lo[0] - o.x) * bi.x, bx = (p.grid.hi[0] - o.x) * bi.x;
                    const float ay = (p.grid.lo[1] - o.y) * bi.y, by = (p.grid.hi[1] - o.y) * bi.y;
                    const float az = (p.grid.lo[2] - o.z) * bi.z, bz = (p.grid.hi[2] - o.z) * bi.z;
                    const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
                    const float t1g = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
                    // the entry cell (clamped: an entry point on the far faces rounds to res); where the ray enters the grid does not
                    // depend on the big primitives, so the first cell entry travels while they are tested
                    bool enters = t0 <= t1g;
                    if (enters) {
                        const float ex = fmaf(d.x, t0, o.x), ey = fmaf(d.y, t0, o.y), ez = fmaf(d.z, t0, o.z);
                        fi.x = fminf(fmaxf(floorf((ex - p.grid.lo[0]) * p.grid.inv_cs[0]), 0.f), (float)(p.grid.res[0] - 1u));
                        fi.y = fminf(fmaxf(floorf((ey - p.grid.lo[1]) * p.grid.inv_cs[1]), 0.f), (float)(p.grid.res[1] - 1u));
                        fi.z = fminf(fmaxf(floorf((ez - p.grid.lo[2]) * p.grid.inv_cs[2]), 0.f), (float)(p.grid.res[2] - 1u));
                        cell = (uint32_t)fi.x + rx * (uint32_t)fi.y + rxy * (uint32_t)fi.z;
                        // boundary an axis crosses next: the upper face of the cell when the ray goes up, else the lower one
                        A = mk<float>(p.grid.cs[0] * bi.x, p.grid.cs[1] * bi.y, p.grid.cs[2] * bi.z);
                        // (by the sign of the CLAMPED direction: a component of +0 walks up, like +1e-18)
                        B = mk<float>(fmaf(p.grid.lo[0] - o.x, bi.x, bi.x > 0.f ? A.x : 0.f), fmaf(p.grid.lo[1] - o.y, bi.y, bi.y > 0.f ? A.y : 0.f),
                                      fmaf(p.grid.lo[2] - o.z, bi.z, bi.z > 0.f ? A.z : 0.f));
                        tmx = mk<float>(fmaf(fi.x, A.x, B.x), fmaf(fi.y, A.y, B.y), fmaf(fi.z, A.z, B.z));
                        advance();
                    }
                    for (uint32_t k = 0; k < p.grid.n_big; k++) { test_bslot(sc, p.grid.big[k], o, d, inv, self, onSurf, best, w_sph); w_prims++; }
                    marching = enters && t0 <= best.t;                       // a hit on a big primitive before the grid's box: nothing to walk
                }
            }
        }
        // ---- what the lanes want: a lane with a ray wants a cell step, a primitive step, or has finished (written out at the top
        // of the next trip).  One vote each; who runs is decided from the two masks.
        const bool want_cell = has && marching && pr_cur >= pr_end;
        bool want_prim = has && pr_cur < pr_end;
        const uint32_t m_cell = __ballot_sync(0xffffffffu, want_cell);
        uint32_t m_prim = __ballot_sync(0xffffffffu, want_prim);
        if ((m_cell | m_prim) == 0) {
            if (exhausted && __ballot_sync(0xffffffffu, has) == 0) break;
            continue;
        }
        // ---- cell step: open the cell whose entry was requested a step ago, request the next one - when enough lanes want it, or
        // nobody holds primitives
        const bool run_cell = m_cell != 0 && ((uint32_t)__popc(m_cell) >= GRID_CELL_MIN || m_prim == 0);
        if (run_cell) {
            if (want_cell) {
                const uint2 ce = ce_nx;
                t_exit = tex_nx;
                pr_cur = ce.x; pr_end = ce.x + ce.y;
                w_cells++; w_occupied += ce.y ? 1u : 0u;
                if (ce.y) slot_nx = __ldg(p.grid_refs + ce.x);
                // the walk ends with this cell when the next one lies outside, or (empty cell) when the nearest hit is before its exit
                marching = ahead_in && !(ce.y == 0u && best.t <= t_exit);
                if (ahead_in) advance();
            }
            want_prim = has && pr_cur < pr_end;                           // the lanes that opened an occupied cell hold primitives now
            m_prim = __ballot_sync(0xffffffffu, want_prim);
        }
        // ---- primitive step: one reference per lane.  Few lanes hold primitives: they wait only if enough lanes march for a cell
        // step to run next (progress either way)
        bool run_prim = m_prim != 0;
        if (run_prim && (uint32_t)__popc(m_prim) < GRID_PRIM_MIN) {
            const uint32_t m_cell_now = run_cell ? __ballot_sync(0xffffffffu, has && marching && pr_cur >= pr_end) : m_cell;
            run_prim = (uint32_t)__popc(m_cell_now) < GRID_CELL_MIN;
        }
        if (run_prim && want_prim) {
            const uint32_t slot = slot_nx;
            float4 a, b;
            ldg256(sc.bslot + 2 * (size_t)slot, a, b);
            pr_cur++;
            if (pr_cur < pr_end) slot_nx = __ldg(p.grid_refs + pr_cur);
            test_brec(sc, a, b, slot, o, d, inv, self, onSurf, best, w_sph);
            w_prims++;
            // the cell is done: stop if the nearest hit lies before its exit
            if (pr_cur >= pr_end && best.t <= t_exit) marching = false;
        }
    }
    my_traced = __reduce_add_sync(0xffffffffu, my_traced);
    if (lane == 0 && my_traced) atomicAdd(p.traced, (unsigned long long)my_traced);
    add_work(p.traced, w_cells, 0u, w_occupied, w_sph, w_prims - w_sph);
}

// ---------------------------------------------------------------------------------------------- 8-wide traversal
// Stage 2 of the split pipeline over the 8-wide quantised tree (ipt_wide.h): one ray per lane, persistent warps with
// lane-level refill as in k_extend_bvh.  What the capture of the 2-wide kernel asked for
// (profiles/r02_ncu_extend_v7.txt: L1 wavefronts 86 %, hit rate 3 %, 17 of 32 lanes, 455 warp instructions per ray, 53
// node visits of 64 bytes per ray) and what a first 8-wide kernel taught (profiles/r02_ncu_wide_v1.txt: 2-4 lanes per
// ray, a stack entry with its entry distance per hit child, a pop loop that culls by distance - 2.2x the instructions of
// the 2-wide kernel, 1.1 against 1.65 Gbounces/s):
//  * a node step reads 96 bytes of ONE line with three 256-bit loads and tests eight quantised child boxes; byte ->
//    plane distance in two instructions: PRMT drops the byte into the mantissa of 2^23 and
//    t = (2^23 + q) * A + (B - 2^23 A) with A = scale / d, B = (origin - o) / d per node and axis (the half step this
//    can be off by is inside the builder's one-step padding); near / far byte chosen by PRMT selectors that depend only
//    on the ray's direction signs;
//  * no per-child stack entries and no distance sort (Ylitie, Karras, Laine 2017): the hit inner children of a node are
//    ONE 64-bit entry {child_base | imask, hit bits}; the bits are permuted by the ray's octant (a 2 KB table in shared
//    memory) so that "highest bit first" is roughly front to back; a child index is child_base + popc(imask below slot);
//  * the hit leaf children become a bit mask over the node's (consecutive) primitive slots, tested before the walk goes
//    on: one primitive per lane and iteration, by the same typed records and arithmetic as the 2-wide path (test_bslot)
//    and the reference's tie rule (Renderer.cu:235) - frames are bit-identical with the 2-wide traversal;
//  * stacks live in shared memory ([entry][lane], conflict-free), one entry per level: 12 entries, deeper ones spill;
//  * node steps and primitive tests are warp-synchronous blocks, each run when enough lanes want it (descend_min,
//    leaf_min) or nobody wants the other.
static constexpr uint32_t CW_STACK = 12;                 // stack entries per lane in shared memory
static constexpr uint32_t CW_SMEM = 2048 + BLOCK_THREADS * CW_STACK * 8;

__device__ __forceinline__ float max3f(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min3f(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
// 2^23 + (byte `sel & 3` of w) as a float: the byte goes into the low mantissa bits of 0x4B000000
__device__ __forceinline__ float u8m(uint32_t w, uint32_t sel) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u), "r"(sel)); return __uint_as_float(r); }
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ uint2 lds64(uint32_t addr) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory"); return v; }

#ifndef IPT_CW_CTAS
#define IPT_CW_CTAS 4
#endif
__global__ void __launch_bounds__(BLOCK_THREADS, IPT_CW_CTAS) k_extend_cw(const __grid_constant__ KParams<float> p)
{
    extern __shared__ uint4 smem[];
    const SceneView<float> sc = p.sc;
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    // octant table: lut[oct * 256 + m] = the hit bits m (bit s = slot s) with slot s moved to bit 7 - (s ^ oct)
    uint8_t* lut = reinterpret_cast<uint8_t*>(smem);
    for (uint32_t i = threadIdx.x; i < 2048u; i += blockDim.x) {
        const uint32_t oct = i >> 8, m = i & 255u;
        uint32_t r = 0;
        for (uint32_t sl = 0; sl < 8; sl++) r |= ((m >> sl) & 1u) << (7u - (sl ^ oct));
        lut[i] = (uint8_t)r;
    }
    __syncthreads();
    const uint32_t lut_base = (uint32_t)__cvta_generic_to_shared(smem);
    // this lane's stack: entry e at sbase + e * 256 (32 lanes x 8 bytes per entry row of the warp)
    const uint32_t sbase = lut_base + 2048u + ((threadIdx.x >> 5) * CW_STACK * 32u + lane) * 8u;
    uint2* spill = p.wide_spill + ((size_t)blockIdx.x * BLOCK_THREADS + threadIdx.x) * p.wide_spill_cap;
    const uint32_t n_in = p.counters[CNT + p.depth];
    uint32_t* work = p.counters + WORK_EXTEND + p.depth;
    const float slack = 1.0000004f, tiny = 1e-18f;

    bool exhausted = false;
    uint32_t my_traced = 0;
    uint32_t w_nodes = 0, w_leaves = 0, w_prims = 0, w_sph = 0;            // work counters of ipt_stats
    bool has = false;
    uint32_t idx = 0, self = NO_OBJECT, sp = 0;
    bool onSurf = false;
    V3<float> o = mk<float>(0, 0, 0), d = o, bi = o, oi = o;
    uint32_t selNx = 0, selFx = 0, selNy = 0, selFy = 0, selNz = 0, selFz = 0, lut_oct = 0, octx = 0;
    uint32_t g_base = 0, g_hits = 0;                                         // node group: child_base | imask << 24, permuted hit bits
    uint32_t p_base = 0, p_valid = 0, p_hits = 0;                            // primitive group: first slot, the node's pmask, bits to test
    Hit<float> best;
    best.t = (float)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;

    for (;;) {
        // ---- a ray with nothing left to visit is finished
        if (has && p_hits == 0 && g_hits == 0 && sp == 0) {
            __stcs(p.hits + idx, make_uint2(__float_as_uint(best.t), best.slot));
            has = false;
        }
        // ---- refill: idle lanes claim the next rays of the queue with one warp-aggregated atomic
        const uint32_t idle = __ballot_sync(0xffffffffu, !has);
        if (!exhausted && (idle == 0xffffffffu || (uint32_t)__popc(idle) >= p.refill_min)) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(work, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, 0);
            exhausted = base + (uint32_t)__popc(idle) >= n_in;
            if (!has) {
                idx = base + __popc(idle & lt_mask);
                if (idx < n_in) {
                    const uint4 a = __ldcs(p.qin.base + idx), b = __ldcs(p.qin.base + p.qin.capacity + idx), c = __ldcs(p.qin.base + 2u * p.qin.capacity + idx);
                    o = mk<float>(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
                    d = mk<float>(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
                    onSurf = (c.z & META_ONSURF) != 0; self = c.w;
                    // box tests: t = plane * (1/d) - o * (1/d); components below 1e-18 become +-1e-18 so that no inf - inf
                    // arises (the boxes are padded by far more than this moves a plane's t)
                    bi = mk<float>(rcp_fast(fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)), rcp_fast(fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
                                   rcp_fast(fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
                    oi = mk<float>(o.x * bi.x, o.y * bi.y, o.z * bi.z);
                    const uint32_t sx = __float_as_uint(d.x) >> 31, sy = __float_as_uint(d.y) >> 31, sz = __float_as_uint(d.z) >> 31;
                    selNx = 0x7650u | sx; selFx = 0x7651u ^ sx;              // word 0 of a child: lo.x hi.x lo.y hi.y
                    selNy = 0x7652u | sy; selFy = 0x7653u ^ sy;
                    selNz = 0x7650u | sz; selFz = 0x7651u ^ sz;              // word 1: lo.z hi.z
                    const uint32_t oct = sx | (sy << 1) | (sz << 2);
                    lut_oct = lut_base + oct * 256u; octx = oct ^ 7u;
                    best.t = (float)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;
                    // the root: "slot 0 of a node whose only inner child is node 0"
                    g_base = 1u << 24; g_hits = 1u << octx; p_hits = 0; sp = 0; has = true;
                    my_traced++;
                }
            }
        }
        const bool want_prim = has && p_hits != 0, want_node = has && p_hits == 0 && (g_hits != 0 || sp != 0);
        const uint32_t m_prim = __ballot_sync(0xffffffffu, want_prim), m_node = __ballot_sync(0xffffffffu, want_node);
        if ((m_prim | m_node) == 0) {
            if (exhausted && __ballot_sync(0xffffffffu, has) == 0) break;
            continue;
        }
        // ---- primitive step: one primitive per lane
        if (m_prim && ((uint32_t)__popc(m_prim) >= p.leaf_min || m_node == 0 || (uint32_t)__popc(m_node) < p.descend_min)) {
            if (want_prim) {
                const uint32_t bit = __ffs(p_hits) - 1u;
                const uint32_t slot = p_base + __popc(p_valid & ((1u << bit) - 1u));
                p_hits &= p_hits - 1u;
                const V3<float> inv = mk<float>(1.f / d.x, 1.f / d.y, 1.f / d.z);   // as the 2-wide path: +-inf for zero components (Plane.cu:55)
                test_bslot(sc, slot, o, d, inv, self, onSurf, best, w_sph);
                w_prims++;
            }
        }
        // ---- node step
        if (m_node && ((uint32_t)__popc(m_node) >= p.descend_min || m_prim == 0)) {
            if (want_node) {
                if (g_hits == 0) {   // the group is used up: the next one comes from the stack
                    sp--;
                    const uint2 e = sp < CW_STACK ? lds64(sbase + sp * 256u) : spill[sp - CW_STACK];
                    g_base = e.x; g_hits = e.y;
                }
                const uint32_t qb = 31u - (uint32_t)__clz(g_hits), slot = qb ^ octx;
                g_hits &= ~(1u << qb);
                const uint32_t node = (g_base & 0xFFFFFFu) + (uint32_t)__popc((g_base >> 24) & ((1u << slot) - 1u));
                if (g_hits) {        // the siblings still to visit wait on the stack
                    if (sp < CW_STACK) sts64(sbase + sp * 256u, g_base, g_hits);
                    else spill[sp - CW_STACK] = make_uint2(g_base, g_hits);
                    sp++;
                }
                const float4* nd = reinterpret_cast<const float4*>(p.wide + node);
                float4 h0, h1, c0, c1, c2, c3;
                ldg256(nd, h0, h1); ldg256(nd + 2, c0, c1); ldg256(nd + 4, c2, c3);
                const float Ax = h0.w * bi.x, Ay = h1.x * bi.y, Az = h1.y * bi.z;
                const float Bx = fmaf(-8388608.f, Ax, fmaf(h0.x, bi.x, -oi.x)), By = fmaf(-8388608.f, Ay, fmaf(h0.y, bi.y, -oi.y)),
                            Bz = fmaf(-8388608.f, Az, fmaf(h0.z, bi.z, -oi.z));
                const float tmax = best.t * slack;
                const uint32_t valid = __float_as_uint(h1.w);
                const uint32_t qw[16] = {__float_as_uint(c0.x), __float_as_uint(c0.y), __float_as_uint(c0.z), __float_as_uint(c0.w),
                                         __float_as_uint(c1.x), __float_as_uint(c1.y), __float_as_uint(c1.z), __float_as_uint(c1.w),
                                         __float_as_uint(c2.x), __float_as_uint(c2.y), __float_as_uint(c2.z), __float_as_uint(c2.w),
                                         __float_as_uint(c3.x), __float_as_uint(c3.y), __float_as_uint(c3.z), __float_as_uint(c3.w)};
                uint32_t hit8 = 0, prims = 0;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const uint32_t w0 = qw[2 * c], w1 = qw[2 * c + 1];
                    const float n = fmaxf(max3f(fmaf(u8m(w0, selNx), Ax, Bx), fmaf(u8m(w0, selNy), Ay, By), fmaf(u8m(w1, selNz), Az, Bz)), 0.f);
                    const float f = fminf(min3f(fmaf(u8m(w0, selFx), Ax, Bx), fmaf(u8m(w0, selFy), Ay, By), fmaf(u8m(w1, selFz), Az, Bz)), tmax);
                    const bool hit = n <= f;
                    hit8 |= hit ? (1u << c) : 0u;
                    prims |= hit ? (valid & (0xFu << (4 * c))) : 0u;
                }
                uint32_t pb;
                asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pb) : "r"(qw[1]), "r"(qw[3]));    // spare bytes of children 0 and 1
                p_base = pb; p_valid = valid; p_hits = prims;
                g_base = __float_as_uint(h1.z);
                const uint32_t inner = hit8 & (g_base >> 24);
                uint32_t gh;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(gh) : "r"(lut_oct + inner));
                g_hits = gh;
                w_nodes++; w_leaves += (uint32_t)__popc(hit8 & ~(g_base >> 24));
            }
        }
    }
    my_traced = __reduce_add_sync(0xffffffffu, my_traced);
    if (lane == 0 && my_traced) atomicAdd(p.traced, (unsigned long long)my_traced);
    add_work(p.traced, w_nodes, 8u * w_nodes, w_leaves, w_sph, w_prims - w_sph);
}

// fp32 contributions are exact in fixed point without going through fp64: v * 2^k is exact in fp32.
__device__ __forceinline__ void accumulate_fast(const KParams<float>& p, uint32_t pixel, V3<float> v)
{
    unsigned long long* f = p.frame + 3ull * pixel;
    if (p.flags & 0x4u) {
        atomicAdd(reinterpret_cast<double*>(f), (double)v.x);
        atomicAdd(reinterpret_cast<double*>(f) + 1, (double)v.y);
        atomicAdd(reinterpret_cast<double*>(f) + 2, (double)v.z);
    } else {
        const float sc = (float)p.fixed_scale;
        atomicAdd(f, (unsigned long long)__float2ll_rn(v.x * sc));
        atomicAdd(f + 1, (unsigned long long)__float2ll_rn(v.y * sc));
        atomicAdd(f + 2, (unsigned long long)__float2ll_rn(v.z * sc));
    }
}

// The product kernel for brute-force scenes: same wavefront step as k_bounce<float, MODE_BRUTE, FIRST>, on the typed
// fp32 scene lists (FastScene) with branch-free intersection and scatter.  What the ncu captures led to
// (profiles/README.md):
//  * slices of 32 rays are claimed by warps from a per-pass counter in chunks, one chunk ahead (see the loop below), so the next
//    slice is known early - its three 16-byte record planes are prefetched with cp.async into a per-thread shared-memory
//    slot while the current slice is computed (double buffered, no barrier: a thread only ever reads what it fetched itself);
//  * compaction is warp ballot + popc into a WARP-PRIVATE block of OUT_BLOCK queue slots, reserved with one atomic per
//    block.  One atomic per warp iteration serialises on the queue counter at ~0.9 G atomics/s on B200 (which capped
//    these scenes at ~29 Grays/s); per-CTA aggregation fixed that but cost two barriers per slice.  The unused tail of
//    a warp's last block is filled with META_DEAD records, which the next pass skips.
static constexpr uint32_t OUT_BLOCK = 128;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// One pass of the typed-list kernel advances every ray by up to nk bounces in registers before it compacts the
// survivors into the next queue.  A closed room loses well under 1 % of its rays per bounce, so compacting and
// round-tripping 96 B per ray through HBM after every bounce buys nothing there; with nk = 8 the queue traffic drops
// to an eighth and, with it, the board power that capped the one-bounce-per-pass kernel.
//
// Pass 0 generates the camera rays and makes the bounces at depth 0 and 1 itself, because those are the ones where a
// path can split (AObject.hpp:91-94, :122-125): the depth-0 split's second ray is a one-cast emission probe
//  and is cast on the spot, the depth-1 split's second ray is appended to the queue next to the
// main ray, both at depth 2.  Later passes only ever see rays that cannot split, all at the same depth (each ray
// carries its depth in meta bits 0-7).  The RNG is keyed by (pixel, sample, lane, depth) and the accumulation is
// fixed point, so the frame and the cast count do not depend on nk.
//
// The schedule lives on the device, so the host never waits for a queue length: pass `p.pass` finds the smallest depth
// its rays can have in counters[WORK + pass] and leaves that + nk for the next one.  nk comes from the survival rate
// the previous pass measured (queue length in / out): idle lanes cost about as much as they save once
// nk * (fraction lost per bounce) exceeds ~0.2 - a leaky scene gets 1-2, a closed room the cap of 8.  Pass 1 has no
// usable measurement of its own batch (camera misses and splits distort it) and starts from what the previous batch
// settled on (*fast_hint: written only by passes >= 2, read only by pass 1, so never inside one grid).  The host launches a fixed number of passes per batch; the remaining depth is spread over the launches
// that are left when that is more than the survival rate asks for (the last launch always finishes the paths), and
// surplus launches return at once.  Every thread of the grid computes the same values from the same finished
// counters.  Returns true when every path of the batch has already ended.
static constexpr uint32_t FAST_K_MAX = 8;
static constexpr uint32_t FAST_LATER_LAUNCHES = 8;       // launches per batch after pass 0 in adaptive mode
template <typename P>
__device__ __forceinline__ bool fast_schedule(const P& p, uint32_t& nk)
{
    const uint32_t dmin = p.pass == 0 ? 0u : p.counters[WORK + p.pass];
    const bool publish = blockIdx.x == 0 && threadIdx.x == 0;
    if (dmin >= p.maxDepth) {
        if (publish && p.pass + 1 < p.n_passes) p.counters[WORK + p.pass + 1] = dmin;
        return true;
    }
    if (p.pass == 0) nk = 2;                             // exactly depth 0 and 1: every ray it queues is at depth 2
    else if (p.fast_k) nk = p.fast_k;
    else if (p.pass == 1) { const uint32_t h = *p.fast_hint; nk = h ? h : 2u; }
    else {
        const uint32_t nk_prev = p.counters[WORK_EXTEND + p.pass - 1];
        const uint32_t n_prev = p.counters[CNT + p.pass - 1], n_cur = p.counters[CNT + p.pass];
        const unsigned long long lost = n_prev > n_cur ? n_prev - n_cur : 1u;
        const unsigned long long k = 22ull * n_prev * nk_prev / (100ull * lost);
        nk = (uint32_t)(k < 1ull ? 1ull : (k > FAST_K_MAX ? FAST_K_MAX : k));
        if (publish && n_prev > (1u << 20)) *p.fast_hint = nk;   // queues this long make the padding of block tails negligible
    }
    const uint32_t remaining = p.maxDepth - dmin, launches_left = p.n_passes - p.pass;
    if (p.pass) nk = max(nk, (remaining + launches_left - 1) / launches_left);
    if (p.pass && !p.fast_k && remaining <= nk + nk / 2) nk = remaining;   // no short trailing pass: a round trip for < nk/2 bounces
    nk = min(nk, remaining);
    if (publish) {
        if (p.pass + 1 < p.n_passes) p.counters[WORK + p.pass + 1] = dmin + nk;
        p.counters[WORK_EXTEND + p.pass] = nk;
    }
    return false;
}

// Threads per CTA and CTAs per SM of the passes after pass 0 (pass 0 uses BLOCK_THREADS x 3)
// (round 1, at 80 registers: 256 x 3 beat 224 x 4 and 128 x 7 by 2-3 %; since the kernel needs 71 the compiler fits it into the 73 of
// 128 x 7 without spills and the four extra warps per SM give +0.8 % on the default bench, +2 % on mirrors.json, scripts/r02_gpu56.sh)
#ifndef IPT_FAST_THREADS
#define IPT_FAST_THREADS 128
#define IPT_FAST_CTAS 7
#endif
#ifndef IPT_FIRST_THREADS
#define IPT_FIRST_THREADS BLOCK_THREADS
#endif
#ifndef IPT_FIRST_CTAS
#define IPT_FIRST_CTAS 3
#endif
// slices of 32 rays a warp claims per atomic: pass 0 (short slices) and the multi-bounce passes
#ifndef IPT_FIRST_CHUNK
#define IPT_FIRST_CHUNK 8u
#endif
#ifndef IPT_DEEP_CHUNK
#define IPT_DEEP_CHUNK 2u
#endif
template <bool FIRST> struct FastCfg { static constexpr int THREADS = FIRST ? IPT_FIRST_THREADS : IPT_FAST_THREADS, CTAS = FIRST ? IPT_FIRST_CTAS : IPT_FAST_CTAS; };

template <bool FIRST, int SHAPE = 0, bool RR = false>
__global__ void __launch_bounds__(FastCfg<FIRST>::THREADS, FastCfg<FIRST>::CTAS) k_bounce_fast(const __grid_constant__ KParams<float> p)
{
    constexpr int BLOCK_THREADS = FastCfg<FIRST>::THREADS;   // shadows the file-wide constant inside this kernel
    uint32_t nk = 1;                                    // bounces a ray makes in this pass
    if (fast_schedule(p, nk)) return;                   // the batch finished in fewer passes than were launched
    extern __shared__ uint4 smem[];
    uint4* stagebuf = smem + p.fast_words;              // [2 buffers][3 planes][BLOCK_THREADS]
    stage(smem, p.fast_blob, p.fast_words);
    __syncthreads();
    const FastScene sc = fast_view(smem, p.fast_hd);

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t n_in = FIRST ? p.n_first : p.counters[CNT + p.pass];
    uint32_t* out_count = p.counters + CNT + p.pass + 1;
    uint32_t my_traced = 0;                              // per lane and pass (summed over the warp at the end): far below 2^32
    uint32_t blk_base = 0, blk_used = OUT_BLOCK;         // no block reserved yet

    // Compaction into the warp's private output block: the rays of the lanes in `mask` fill the rest of the current
    // block and spill into a freshly reserved one (one atomic per OUT_BLOCK outputs, no holes inside blocks).
    auto emit = [&](bool has, uint32_t mask, const Ray<float>& ray) {
        const uint32_t tot = __popc(mask);
        const uint32_t room = OUT_BLOCK - blk_used;
        uint32_t nb = blk_base;
        if (tot > room) {
            if (lane == 0) nb = atomicAdd(out_count, OUT_BLOCK);
            nb = __shfl_sync(0xffffffffu, nb, 0);
        }
        const uint32_t j = __popc(mask & lt_mask);
        if (has) q_store(p.qout, j < room ? blk_base + blk_used + j : nb + (j - room), ray);
        if (tot > room) { blk_base = nb; blk_used = tot - room; }
        else blk_used += tot;
    };

    // Work is claimed from a per-pass counter in chunks of CHUNK slices of 32 rays, one chunk ahead: the chunk being computed,
    // the next one (claimed, its first records already on their way into shared memory when the current one ends) and the
    // claim after that (issued, not yet read).  Static assignment (slice = warp + k * warps) left the SMs idle at the end of a
    // pass - warps on SMs with slower memory finish late; ncu showed 30.5 % achieved against 37.5 % theoretical occupancy
    // (profiles/r01_ncu_spheres4k_final_deep_pass.txt) - and one claim per slice ran into the ~0.5 G/s a single address takes
    // in pass 0, whose slices are short (profiles/r02_ncu_pass0_v2.txt: 28 % of the stall samples on the claim).
    constexpr uint32_t CHUNK = FIRST ? IPT_FIRST_CHUNK : IPT_DEEP_CHUNK, CHUNK_RAYS = CHUNK * 32u;
    uint32_t* claim_ctr = p.counters + CLAIM + p.pass;
    // (p.static_slices, IPT_STATIC_SLICES=1: the static assignment, for A/B runs - no atomics)
    const uint32_t stride = gridDim.x * (BLOCK_THREADS / 32) * CHUNK_RAYS;
    uint32_t static_next = (blockIdx.x * (BLOCK_THREADS / 32) + (threadIdx.x >> 5)) * CHUNK_RAYS;
    auto claim_issue = [&]() {
        uint32_t v = 0;
        if (p.static_slices) { v = static_next; static_next = static_next + stride < static_next ? 0xFFFFFF00u : static_next + stride; }
        else if (lane == 0) v = atomicAdd(claim_ctr, CHUNK_RAYS);
        return v;
    };
    auto prefetch = [&](uint32_t base, int b) {
        const uint32_t j = base + lane;
        if (!FIRST && j < n_in)
            for (int pl = 0; pl < 3; pl++) cp_async16(stagebuf + (b * 3 + pl) * BLOCK_THREADS + threadIdx.x, p.qin.base + (size_t)pl * p.qin.capacity + j);
        cp_async_commit();
    };
    uint32_t chunk = __shfl_sync(0xffffffffu, claim_issue(), 0);
    uint32_t chunk_next = __shfl_sync(0xffffffffu, claim_issue(), 0);
    uint32_t pending = claim_issue();
    uint32_t off = 0;
    int buf = 0;
    prefetch(chunk, buf);

    for (;;) {
        const uint32_t cur = chunk + off;
        if (cur >= n_in) break;                              // claims only grow: nothing is left for this warp
        const uint32_t noff = off + 32u;
        const uint32_t nxt = noff < CHUNK_RAYS ? chunk + noff : chunk_next;
        const uint32_t i = cur + lane;
        bool live = i < n_in;
        Ray<float> r;
        if (FIRST) {
            uint32_t px = 0, pz = 0, sample = 0;
            live = live && decode_sample(p, i, px, pz, sample);
            if (live) camera_ray(p, px, pz, sample, r);
        } else {
            cp_async_wait_all();
            prefetch(nxt, buf ^ 1);
            if (live) {
                const uint4 a = stagebuf[(buf * 3 + 0) * BLOCK_THREADS + threadIdx.x], b = stagebuf[(buf * 3 + 1) * BLOCK_THREADS + threadIdx.x],
                            c = stagebuf[(buf * 3 + 2) * BLOCK_THREADS + threadIdx.x];
                r.o = mk<float>(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
                r.d = mk<float>(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
                r.thr = mk<float>(__uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(c.x));
                r.pixel = c.y; r.meta = c.z; r.self = c.w;
                live = !(r.meta & META_DEAD);
            }
            buf ^= 1;
        }
        off = noff;
        if (off == CHUNK_RAYS) { chunk = chunk_next; chunk_next = __shfl_sync(0xffffffffu, pending, 0); pending = claim_issue(); off = 0; }
        bool has0 = live;
        // The meta word taken apart for the bounces made in registers and put together again for the queue: the sample and lane
        // bits never change, the depth counts up, 'starts on a surface' follows the last hit (the masks and shifts on the packed
        // word were 5 of ~400 warp instructions per bounce).
        const uint32_t meta_keep = r.meta & 0x0FFFF300u, sample = (r.meta >> 12) & 0xFFFFu, ctr_lane = r.meta & 0x300u;
        uint32_t depth = r.meta & 0xFFu;
        bool on_surf = (r.meta & META_ONSURF) != 0;
        for (uint32_t k = 0; k < nk; k++) {
            // (no vote on "has every ray of this slice ended": a closed room loses under 1 % of its rays per bounce, a leaky scene
            // gets a small nk from fast_schedule; the vote, its branch and the popc were 5 of ~410 warp instructions per bounce)
            const bool in = has0;
            has0 = false;
            bool has1 = false;
            Ray<float> o1;
            if (in) {
                my_traced++;
                const uint32_t dk = FIRST ? k : depth;
                const FastHit h = nearest_fast<SHAPE>(sc, r.o, r.d, r.self, on_surf);
                if (h.code != NO_OBJECT) {
                    const uint32_t obj = fast_hit_object(sc, h.code) & ~RECT_BIT;
                    const float4 m0 = sc.mat[2 * obj], m1 = sc.mat[2 * obj + 1];
                    if (m1.w != 0.f) accumulate_fast(p, r.pixel, mul(r.thr, mk<float>(m1.x, m1.y, m1.z)));
                    V3<float> nthr = mul(r.thr, mk<float>(m0.x, m0.y, m0.z));
                    const bool go = dk + 1 < p.maxDepth && any_nonzero(nthr);   // no probes in these queues
                    if (go) {
                        const V3<float> P = fast_hit_point(sc, h.code, r.o, r.d, h.t);
                        const uint4 rnd = philox4x32(r.pixel, sample, ctr_lane | dk, CTR_TAG, p.keys);
                        const Spawn<float> sp = scatter_fast<(SHAPE > 0)>(sc, h.code, (int)m0.w, P, r.d, FIRST ? dk : 2u, rnd);
                        bool alive = sp.has0;
                        if (RR && dk >= 3 && alive) {                 // IPT_FLAG_RUSSIAN_ROULETTE (extension): its own instantiations - as a
                                                                      // flag tested in the loop it cost 3 instructions per bounce and 8 registers
                            const float q = fminf(1.f, fmaxf(0.05f, fmaxf(nthr.x, fmaxf(nthr.y, nthr.z))));
                            const uint4 rr = philox4x32(r.pixel, sample, ctr_lane | dk, CTR_TAG + 1u, p.keys);
                            if (u23<float>(rr.x) >= q) alive = false;
                            else nthr = nthr * (1.f / q);
                        }
                        const bool onS = (h.code >> 28) != 0 || fabsf(dot(r.d, r.d) - 1.f) < 1e-3f;
                        if (FIRST && k < 2) {
                            has1 = sp.has1;
                            o1.o = P; o1.d = sp.d1; o1.thr = nthr * sp.w1; o1.pixel = r.pixel; o1.self = h.code;
                            o1.meta = (meta_keep & 0x0FFFF000u) | (onS ? META_ONSURF : 0u) | (dk + 1) | (k == 0 ? (0x200u | META_PROBE) : 0x100u);
                        }
                        has0 = alive;
                        r.o = P; r.d = sp.d0; r.thr = nthr * sp.w0; r.self = h.code;
                        depth = dk + 1; on_surf = onS;
                    }
                }
            }
            if (FIRST && k < 2) {
                const uint32_t m1b = __ballot_sync(0xffffffffu, has1);
                if (m1b && k == 0) {
                    // the second ray of a depth-0 split only ever contributes the emission of the first thing it hits
                    // (SURVEY.md App. A.6): cast it here instead of sending it through a queue
                    if (has1) {
                        my_traced++;
                        const FastHit hp = nearest_fast<SHAPE>(sc, o1.o, o1.d, o1.self, (o1.meta & META_ONSURF) != 0);
                        if (hp.code != NO_OBJECT) {
                            const float4 e = sc.mat[2 * (fast_hit_object(sc, hp.code) & ~RECT_BIT) + 1];
                            if (e.w != 0.f) accumulate_fast(p, o1.pixel, mul(o1.thr, mk<float>(e.x, e.y, e.z)));
                        }
                    }
                } else if (m1b) emit(has1, m1b, o1);       // depth-1 split: a full path from depth 2 on
            }
        }
        const uint32_t m0b = __ballot_sync(0xffffffffu, has0);
        r.meta = meta_keep | (on_surf ? META_ONSURF : 0u) | depth;
        if (m0b) emit(has0, m0b, r);
    }
    // the unused tail of the last block: dead records (skipped by the next pass)
    for (uint32_t s = blk_used + lane; s < OUT_BLOCK; s += 32) p.qout.base[2u * p.qout.capacity + blk_base + s] = make_uint4(0u, 0u, META_DEAD, NO_OBJECT);
    my_traced = __reduce_add_sync(0xffffffffu, my_traced);
    if (lane == 0 && my_traced) atomicAdd(p.traced, (unsigned long long)my_traced);
}

// End of a batch: every queue record was written once and read once, and counters[CNT + d] is the length of the queue
// a pass at depth d read (0 for depths no pass started at).  stats[1] += 2 * sum of those (ipt_stats.queue_bytes).
__global__ void k_batch_stats(const uint32_t* __restrict__ counters, unsigned long long* stats)
{
    unsigned long long s = 0;
    for (int d = threadIdx.x; d < MAX_PASSES; d += 32) s += counters[CNT + d];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) stats[1] += 2ull * s;
}

// Diagnostic (IPT_PASS_TIMES): SM clock right now, from ~8 us of %clock64 against %globaltimer on one warp.
__global__ void k_clock_probe(float* out_mhz)
{
    if (threadIdx.x != 0) return;
    unsigned long long t0, t1, c0, c1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    c0 = clock64();
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < 8000ull);
    c1 = clock64();
    *out_mhz = (float)((double)(c1 - c0) / (double)(t1 - t0) * 1e3);
}

// Frame accumulators -> mean radiance per pixel (Renderer.cu:142-144), for the tiles this rank owns.  `dst32/dst64`
// may point into another GPU's frame (NVLink peer access): the gather to rank 0 is these stores.
struct ResolveParams {
    const unsigned long long* frame;
    float* dst32;
    double* dst64;
    const uint32_t* tile_ids;
    uint32_t n_tiles_local, tiles_x, tile_w, tile_h, W, H;
    double inv;          // 1 / (fixed_scale * spp)  or 1 / spp with float accumulation
    uint32_t float_accum;
};

__global__ void __launch_bounds__(256) k_resolve(const __grid_constant__ ResolveParams p)
{
    const uint32_t per_tile = p.tile_w * p.tile_h;
    const unsigned long long total = (unsigned long long)p.n_tiles_local * per_tile;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t lt = (uint32_t)(i / per_tile), in = (uint32_t)(i % per_tile);
        const uint32_t tile = p.tile_ids[lt];
        const uint32_t px = (tile % p.tiles_x) * p.tile_w + in % p.tile_w, pz = (tile / p.tiles_x) * p.tile_h + in / p.tile_w;
        if (px >= p.W || pz >= p.H) continue;
        const size_t pixel = (size_t)pz * p.W + px;
        for (int c = 0; c < 3; c++) {
            const unsigned long long a = p.frame[3 * pixel + c];
            const double v = p.float_accum ? __longlong_as_double((long long)a) * p.inv : (double)(long long)a * p.inv;
            if (p.dst64) p.dst64[3 * pixel + c] = v;
            if (p.dst32) p.dst32[3 * pixel + c] = (float)v;
        }
    }
}

// Image.cpp:19-22 on the device: byte = clamp(int(x * 255), 0, 255), computed in fp64 from the fp32 frame exactly as
// ipt_host_to_rgb does on the host (NaN -> 0, saturation instead of the reference's undefined int overflow).
__global__ void __launch_bounds__(256) k_to_rgb8(const float* __restrict__ frame, uint8_t* __restrict__ out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = (double)frame[i] * 255.0;
        int b = 0;
        if (v == v) b = v >= 2147483647.0 ? 255 : (v <= -2147483648.0 ? 0 : min(max((int)v, 0), 255));
        out[i] = (uint8_t)b;
    }
}

// Function-level access for parity tests: nearest hit of explicit rays through the same device functions.
template <typename R, int MODE>
__global__ void __launch_bounds__(BLOCK_THREADS) k_trace(SceneView<R> scv, const double* rays, uint32_t n, int32_t* out_obj, double* out_t)
{
    extern __shared__ uint4 smem[];
    SceneView<R> sc = scv;
    const float4* top = nullptr;
    uint32_t n_top = 0;
    if (MODE == MODE_BRUTE) {
        constexpr uint32_t W16 = sizeof(R4<R>) / 16;
        const uint32_t ng = sc.n_slots * 4 * W16, ni = (sc.n_slots + 3) / 4;
        stage(smem, reinterpret_cast<const uint4*>(scv.geom), ng);
        stage(smem + ng, reinterpret_cast<const uint4*>(scv.slot_obj), ni);
        sc.geom = reinterpret_cast<const R4<R>*>(smem);
        sc.slot_obj = reinterpret_cast<const uint32_t*>(smem + ng);
    } else {
        n_top = sc.n_nodes < (uint32_t)BVH_TOP_NODES ? sc.n_nodes : (uint32_t)BVH_TOP_NODES;
        stage(smem, reinterpret_cast<const uint4*>(scv.nodes), n_top * 4);
        top = reinterpret_cast<const float4*>(smem);
    }
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const V3<R> o = mk<R>((R)rays[6 * i], (R)rays[6 * i + 1], (R)rays[6 * i + 2]);
        const V3<R> d = mk<R>((R)rays[6 * i + 3], (R)rays[6 * i + 4], (R)rays[6 * i + 5]);
        const Hit<R> h = nearest_any<R, MODE>(sc, top, n_top, o, d, NO_OBJECT, false);
        out_obj[i] = h.slot == NO_OBJECT ? -1 : (int32_t)(h.obj & ~RECT_BIT);
        out_t[i] = (double)h.t;
    }
}

}  // namespace ipt
