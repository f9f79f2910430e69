// ipt_device.cuh — device-side building blocks of the B200 radiance path: vectors, the counter-based RNG,
// nearest-hit tests for spheres and finite rectangles, the three scatter rules, and the ray-record codec.
// Templated on the arithmetic type R: float is the product path, double is the parity mode (IPT_FLAG_FP64).
//
// What is computed follows the reference (file:line cited at each function, paths relative to the
// AdamStudies-PWR/Improved-Path-Tracer tree); how it is computed does not: no virtual objects, no per-thread
// scene copies, no recursion — rays are records in a queue and every function here is a pure function of a record.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipt {

// ---------------------------------------------------------------------------------------------- constants
#define IPT_MARGIN 1e-4          /* scene/cuda/objects/Constants.hpp:8  */
#define IPT_INF 1e20             /* Renderer.cu:29                       */
#define IPT_VIEWPORT_DISTANCE 140.0  /* Renderer.cu:28                   */
static constexpr uint32_t NODE_CAMERA = 0xFFFFu;
static constexpr uint32_t CTR_TAG = 0x49505442u;  // "IPTB"
static constexpr uint32_t NO_OBJECT = 0xFFFFFFFFu;
static constexpr uint32_t RECT_BIT = 0x80000000u;

// ray meta word: depth[0:8) lane[8:10) probe[10] onSurf[11] sample[12:28)
static constexpr uint32_t META_PROBE = 1u << 10;
static constexpr uint32_t META_ONSURF = 1u << 11;
static constexpr uint32_t META_NEE = 1u << 28;    // the hit this ray started from sampled the emissive spheres explicitly
static constexpr uint32_t META_DEAD = 1u << 31;   // padding slot at the end of a warp's output block (k_bounce_fast): not a ray
__host__ __device__ inline uint32_t make_meta(uint32_t depth, uint32_t lane, bool probe, bool onSurf, uint32_t sample)
{
    return depth | (lane << 8) | (probe ? META_PROBE : 0u) | (onSurf ? META_ONSURF : 0u) | (sample << 12);
}

// ---------------------------------------------------------------------------------------------- vectors
template <typename R> struct V3 { R x, y, z; };
template <typename R> struct alignas(sizeof(R) * 4) R4 { R x, y, z, w; };

template <typename R> __device__ __forceinline__ V3<R> mk(R x, R y, R z) { V3<R> v; v.x = x; v.y = y; v.z = z; return v; }
template <typename R> __device__ __forceinline__ V3<R> operator+(V3<R> a, V3<R> b) { return mk<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename R> __device__ __forceinline__ V3<R> operator-(V3<R> a, V3<R> b) { return mk<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename R> __device__ __forceinline__ V3<R> operator*(V3<R> a, R s) { return mk<R>(a.x * s, a.y * s, a.z * s); }
template <typename R> __device__ __forceinline__ V3<R> operator-(V3<R> a) { return mk<R>(-a.x, -a.y, -a.z); }
template <typename R> __device__ __forceinline__ R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename R> __device__ __forceinline__ V3<R> mul(V3<R> a, V3<R> b) { return mk<R>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename R> __device__ __forceinline__ V3<R> xyz(const R4<R>& q) { return mk<R>(q.x, q.y, q.z); }

// MUFU.RSQ alone: rsqrtf() wraps the same instruction in a denormal guard (FSETP + two predicated FMUL, twice per bounce in
// the typed-list kernel's capture) that no argument here can need - the squared lengths are >= 3 * 2^-48 (s24 draws) or ~r^2;
// results for normal arguments are bit-identical (the committed frame hashes did not move).
__device__ __forceinline__ float rsqrt_(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }  // Vec3.hpp:48-51: v * (1/sqrt(v.v))
__device__ __forceinline__ float div_(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float sqrt_(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ double div_(double a, double b) { return a / b; }
template <typename R> __device__ __forceinline__ V3<R> cross(V3<R> a, V3<R> b) { return mk<R>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <typename R> __device__ __forceinline__ V3<R> normalize(V3<R> a) { return a * rsqrt_(dot(a, a)); }

// ---------------------------------------------------------------------------------------------- RNG
// Philox4x32-7 (Salmon, Moraes, Dror, Shaw, SC'11: Philox4x32 with R rounds; 7 is the smallest R that is Crush-resistant
// in their Table 2, 10 their default with a safety margin; Random123's known answers exist for both and pin the oracle's
// copy, tests/test_oracle_pin.py).  Round 1 used 10 rounds: 81 of ~540 warp instructions per 32 rays and bounce
// (profiles/r01_ncu_spheres4k_final_deep_pass.txt), 7 rounds save 24 of them.  One block = the four uniforms one scatter event can use,
// addressed by (pixel, sample, lane<<8|depth): no generator state travels with a ray, and the image does not
// depend on how rays are scheduled, batched, tiled or split over GPUs.  The reference seeds one XORWOW stream per
// CUDA thread (Renderer.cu:95-97) and walks it through ~1900 pixels, which no parallel schedule can reproduce;
// parity with it is therefore statistical, and exact against the oracle run on this same counter stream.
// The round keys k + r*W are the same for every thread of a render: the host expands them once into the kernel
// parameters (constant bank), so a round is two IMAD.WIDE and two LOP3 with a constant operand.
#ifndef IPT_PHILOX_ROUNDS
#define IPT_PHILOX_ROUNDS 7
#endif
static constexpr int PHILOX_ROUNDS = IPT_PHILOX_ROUNDS;
struct PhiloxKeys { uint32_t k[2 * PHILOX_ROUNDS]; };
__host__ __device__ inline PhiloxKeys philox_expand(uint32_t k0, uint32_t k1)
{
    PhiloxKeys ks;
    for (int r = 0; r < PHILOX_ROUNDS; r++) { ks.k[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u; ks.k[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u; }
    return ks;
}
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& ks)
{
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; r++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ ks.k[2 * r]; c1 = (uint32_t)p1; c2 = (uint32_t)(p0 >> 32) ^ c3 ^ ks.k[2 * r + 1]; c3 = (uint32_t)p0;
    }
    return make_uint4(c0, c1, c2, c3);
}
// Random reals, exactly representable in fp32 AND fp64 so both precisions and the CPU oracle see the same numbers:
//   s24(x) in (-1,1): the odd multiple (2k+1-2^24)/2^24 of 2^-24, k = x>>8  — the reference's one_one() = 2u-1
//                     (CudaUtils.hpp:14-17) at 24-bit resolution; never 0, so AObject.hpp:38's retry loop never runs
//   u23(x) in (0,1) : (2k+1)/2^24, k = x>>9 — the uniform of the stochastic lobe pick (AObject.hpp:94,127)
template <typename R> __device__ __forceinline__ R s24(uint32_t x) { return ((R)((int)(x >> 8) - 8388608) + (R)0.5) * (R)(1.0 / 8388608.0); }
template <typename R> __device__ __forceinline__ R u23(uint32_t x) { return ((R)(x >> 9) + (R)0.5) * (R)(1.0 / 8388608.0); }
// fp32: (k + 0.5) * 2^-23 = fma(k, 2^-23, 2^-24) - every intermediate and the result are exactly representable (|k| <= 2^23), so
// the one rounding of the FMA returns the same number as the add and the multiply above, in one instruction instead of two
template <> __device__ __forceinline__ float s24<float>(uint32_t x) { return fmaf((float)((int)(x >> 8) - 8388608), 1.f / 8388608.f, 1.f / 16777216.f); }
template <> __device__ __forceinline__ float u23<float>(uint32_t x) { return fmaf((float)(x >> 9), 1.f / 8388608.f, 1.f / 16777216.f); }

// ---------------------------------------------------------------------------------------------- ray record
template <typename R> struct Ray {
    V3<R> o, d, thr;
    uint32_t pixel, meta, self;
};

// Queue = structure of arrays of 16-byte planes: a warp reads/writes 512 contiguous bytes per plane (128-bit,
// fully coalesced).  fp32: 3 planes (48 B/ray); fp64: 6 planes (96 B/ray).
template <typename R> struct QPlanes;
template <> struct QPlanes<float> { static constexpr int N = 3, ACC = 1; };
template <> struct QPlanes<double> { static constexpr int N = 6, ACC = 2; };

struct Queue {
    uint4* base;        // plane p of ray i at base[p * capacity + i]
    uint32_t capacity;
};

__device__ __forceinline__ void q_store(const Queue& q, uint32_t i, const Ray<float>& r)
{
    q.base[i] = make_uint4(__float_as_uint(r.o.x), __float_as_uint(r.o.y), __float_as_uint(r.o.z), __float_as_uint(r.d.x));
    q.base[q.capacity + i] = make_uint4(__float_as_uint(r.d.y), __float_as_uint(r.d.z), __float_as_uint(r.thr.x), __float_as_uint(r.thr.y));
    q.base[2u * q.capacity + i] = make_uint4(__float_as_uint(r.thr.z), r.pixel, r.meta, r.self);
}
__device__ __forceinline__ void q_load(const Queue& q, uint32_t i, Ray<float>& r)
{
    const uint4 a = q.base[i], b = q.base[q.capacity + i], c = q.base[2u * q.capacity + i];
    r.o = mk<float>(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
    r.d = mk<float>(__uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y));
    r.thr = mk<float>(__uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(c.x));
    r.pixel = c.y; r.meta = c.z; r.self = c.w;
}
__device__ __forceinline__ uint4 pack2(double a, double b)
{
    const unsigned long long x = (unsigned long long)__double_as_longlong(a), y = (unsigned long long)__double_as_longlong(b);
    return make_uint4((uint32_t)x, (uint32_t)(x >> 32), (uint32_t)y, (uint32_t)(y >> 32));
}
__device__ __forceinline__ void unpack2(uint4 v, double& a, double& b)
{
    a = __longlong_as_double((long long)(((unsigned long long)v.y << 32) | v.x));
    b = __longlong_as_double((long long)(((unsigned long long)v.w << 32) | v.z));
}
__device__ __forceinline__ void q_store(const Queue& q, uint32_t i, const Ray<double>& r)
{
    const size_t c = q.capacity;
    q.base[i] = pack2(r.o.x, r.o.y);
    q.base[c + i] = pack2(r.o.z, r.d.x);
    q.base[2 * c + i] = pack2(r.d.y, r.d.z);
    q.base[3 * c + i] = pack2(r.thr.x, r.thr.y);
    uint4 e = pack2(r.thr.z, 0.0);
    e.z = r.pixel; e.w = r.meta;
    q.base[4 * c + i] = e;
    q.base[5 * c + i] = make_uint4(r.self, 0u, 0u, 0u);
}
__device__ __forceinline__ void q_load(const Queue& q, uint32_t i, Ray<double>& r)
{
    const size_t c = q.capacity;
    unpack2(q.base[i], r.o.x, r.o.y);
    unpack2(q.base[c + i], r.o.z, r.d.x);
    unpack2(q.base[2 * c + i], r.d.y, r.d.z);
    unpack2(q.base[3 * c + i], r.thr.x, r.thr.y);
    const uint4 e = q.base[4 * c + i];
    double dummy;
    unpack2(e, r.thr.z, dummy);
    r.pixel = e.z; r.meta = e.w;
    r.self = q.base[5 * c + i].x;
}

// Deferred radiance of a deep path (only when maxDepth >= 130, see k_bounce): extra planes after the regular ones.
__device__ __forceinline__ void q_store_acc(const Queue& q, uint32_t i, const V3<float>& a)
{
    q.base[3u * q.capacity + i] = make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), 0u);
}
__device__ __forceinline__ void q_load_acc(const Queue& q, uint32_t i, V3<float>& a)
{
    const uint4 v = q.base[3u * q.capacity + i];
    a = mk<float>(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z));
}
__device__ __forceinline__ void q_store_acc(const Queue& q, uint32_t i, const V3<double>& a)
{
    const size_t c = q.capacity;
    q.base[6 * c + i] = pack2(a.x, a.y);
    q.base[7 * c + i] = pack2(a.z, 0.0);
}
__device__ __forceinline__ void q_load_acc(const Queue& q, uint32_t i, V3<double>& a)
{
    const size_t c = q.capacity;
    double dummy;
    unpack2(q.base[6 * c + i], a.x, a.y);
    unpack2(q.base[7 * c + i], a.z, dummy);
}

// ---------------------------------------------------------------------------------------------- scene view
// Geometry "slots": 4 x R4 per primitive.
//   sphere    : g[0] = {cx, cy, cz, r}
//   rectangle : g[0] = {n.xyz, D}  g[1] = {u.xyz, u.c}  g[2] = {v.xyz, v.c}  g[3] = {u_lo, u_hi, v_lo, v_hi}
// slot_obj[s] = object (JSON) index, bit 31 set for rectangles.
// mat[2k] = {color.rgb, reflection}, mat[2k+1] = {emission.rgb, any(emission != 0)} per OBJECT k.
template <typename R> struct SceneView {
    const R4<R>* geom;
    const uint32_t* slot_obj;
    const R4<R>* mat;
    uint32_t n_slots, n_spheres;   // brute-force layout: spheres occupy slots [0, n_spheres), rectangles the rest
    uint32_t n_objects;
    const float4* nodes;           // BVH: 4 x float4 per node (see unpack in traverse), may be null
    uint32_t n_nodes;
    const float4* bslot;           // fp32 BVH leaf records, 2 x float4 per slot (see nearest_bvh_f32), may be null
    const double* lights;          // IPT_FLAG_NEXT_EVENT: 8 doubles per emissive sphere {c.xyz, r, E.rgb, object index}
    uint32_t n_lights;
};

template <typename R> struct Hit {
    R t;
    uint32_t slot;      // NO_OBJECT = miss
    uint32_t obj;       // object index | RECT_BIT
};

// Self-hit policy.  The reference rejects self intersections only through `t > 1e-4` (Sphere.cu:36-37, Plane.cu:58),
// which fp32 cannot resolve at |x| ~ 1e3.  In fp32 a ray therefore carries the object it starts on and
//   * never re-tests the rectangle it starts on (its only root is t ~ 0),
//   * on the sphere it starts on, when the start point is known to lie ON the surface (the ray that produced the hit
//     had unit length, so P = o + d t is a surface point), uses the exact second root t = -2 (op.d) of
//     t^2 + 2 b t + (op.op - r^2) = 0 with op.op - r^2 = 0 — the same root the reference's formula returns.
// In fp64 (parity mode) the literal tests are used.  Validated in SURVEY.md App. D (D10).
template <typename R> struct SelfRule { static constexpr bool enabled = false; };
template <> struct SelfRule<float> { static constexpr bool enabled = true; };

// Sphere.cu:25-39.  Formula kept as is for non-unit directions (refracted rays are not normalised, AObject.hpp:59).
template <typename R>
__device__ __forceinline__ void test_sphere(const R4<R> s, uint32_t slot, uint32_t obj, const V3<R> o, const V3<R> d,
                                            uint32_t self, bool onSurf, Hit<R>& best)
{
    const V3<R> op = o - xyz(s);
    const R b = dot(op, d);
    R t;
    if (SelfRule<R>::enabled && obj == self && onSurf) {
        t = (R)-2 * b;
        if (!(t > (R)IPT_MARGIN)) return;
    } else {
        const R delta = b * b - dot(op, op) + s.w * s.w;
        if (delta < (R)0) return;
        const R sq = sqrt(delta);
        t = -b - sq;
        if (!(t > (R)IPT_MARGIN)) {
            t = -b + sq;
            if (!(t > (R)IPT_MARGIN)) return;
        }
    }
    // Renderer.cu:235: `temp && temp < distance`, objects scanned in index order -> lowest index wins ties
    if (t < best.t || (t == best.t && obj < best.obj)) { best.t = t; best.slot = slot; best.obj = obj; }
}

// Plane.cu:47-68 with the bounds test of :87-100 solved for the hit position (see ipt_abi.h / DESIGN.md §4).
template <typename R>
__device__ __forceinline__ void test_rect(const R4<R>* g, uint32_t slot, uint32_t obj, const V3<R> o, const V3<R> d,
                                          uint32_t self, Hit<R>& best)
{
    if (SelfRule<R>::enabled && obj == self) return;
    const R4<R> pl = g[0];
    const R den = pl.x * d.x + pl.y * d.y + pl.z * d.z;
    if (den == (R)0) return;                                   // Plane.cu:55
    const R t = div_(pl.w - (pl.x * o.x + pl.y * o.y + pl.z * o.z), den);
    if (!(t > (R)IPT_MARGIN)) return;                          // Plane.cu:58 (NaN fails too)
    if (!(t < best.t || (t == best.t && obj < best.obj))) return;
    const V3<R> P = o + d * t;
    const R4<R> gu = g[1], gv = g[2], bd = g[3];
    const R su = fabs(gu.x * P.x + gu.y * P.y + gu.z * P.z - gu.w);
    const R sv = fabs(gv.x * P.x + gv.y * P.y + gv.z * P.z - gv.w);
    if (su >= bd.x && su <= bd.y && sv >= bd.z && sv <= bd.w) { best.t = t; best.slot = slot; best.obj = obj; }
}

// Renderer.cu:227-243 over a brute-force slot list (every lane of a warp walks the same slots: broadcast reads).
template <typename R>
__device__ __forceinline__ Hit<R> nearest_brute(const SceneView<R>& sc, const V3<R> o, const V3<R> d, uint32_t self, bool onSurf)
{
    Hit<R> best;
    best.t = (R)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;
    const uint32_t ns = sc.n_spheres, n = sc.n_slots;
#pragma unroll 2
    for (uint32_t s = 0; s < ns; s++) test_sphere<R>(sc.geom[4 * s], s, sc.slot_obj[s], o, d, self, onSurf, best);
#pragma unroll 2
    for (uint32_t s = ns; s < n; s++) test_rect<R>(sc.geom + 4 * s, s, sc.slot_obj[s], o, d, self, best);
    return best;
}

// BVH2 traversal: nodes are four float4 (128-bit loads); boxes are fp32 and padded by the builder, the slab test is
// conservative (fp64 rays are tested against the fp32 boxes in fp64).  `top` points at a shared-memory copy of the
// first n_top nodes (the builder emits nodes breadth-first, so these are the top levels every ray visits).
template <typename R>
__device__ __forceinline__ Hit<R> nearest_bvh(const SceneView<R>& sc, const float4* top, uint32_t n_top, const V3<R> o,
                                              const V3<R> d, uint32_t self, bool onSurf)
{
    Hit<R> best;
    best.t = (R)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;
    const R ix = (R)1 / d.x, iy = (R)1 / d.y, iz = (R)1 / d.z;   // +-inf for zero components: slab test handles it
    int stack[64];
    int sp = 0;
    int node = 0;
    const R slack = (R)1.0000004;   // 3 ulp(fp32) on the far distance: never cull a box the exact test would keep
    for (;;) {
        if (node >= 0) {
            const float4* p = ((uint32_t)node < n_top) ? (top + 4 * node) : (sc.nodes + 4 * (size_t)node);
            float4 a, b, c, e;
            if ((uint32_t)node < n_top) { a = p[0]; b = p[1]; c = p[2]; e = p[3]; }
            else { a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); e = __ldg(p + 3); }
            // a = lo0.xyz hi0.x | b = hi0.yz lo1.xy | c = lo1.z hi1.xyz | e = child0 child1 (int bits)
            R t0x = ((R)a.x - o.x) * ix, t1x = ((R)a.w - o.x) * ix;
            R t0y = ((R)a.y - o.y) * iy, t1y = ((R)b.x - o.y) * iy;
            R t0z = ((R)a.z - o.z) * iz, t1z = ((R)b.y - o.z) * iz;
            R n0 = fmax(fmax(fmin(t0x, t1x), fmin(t0y, t1y)), fmax(fmin(t0z, t1z), (R)0));
            R f0 = fmin(fmin(fmax(t0x, t1x), fmax(t0y, t1y)), fmin(fmax(t0z, t1z), best.t)) * slack;
            t0x = ((R)b.z - o.x) * ix; t1x = ((R)c.y - o.x) * ix;
            t0y = ((R)b.w - o.y) * iy; t1y = ((R)c.z - o.y) * iy;
            t0z = ((R)c.x - o.z) * iz; t1z = ((R)c.w - o.z) * iz;
            R n1 = fmax(fmax(fmin(t0x, t1x), fmin(t0y, t1y)), fmax(fmin(t0z, t1z), (R)0));
            R f1 = fmin(fmin(fmax(t0x, t1x), fmax(t0y, t1y)), fmin(fmax(t0z, t1z), best.t)) * slack;
            const bool h0 = n0 <= f0, h1 = n1 <= f1;
            const int c0 = __float_as_int(e.x), c1 = __float_as_int(e.y);
            if (h0 && h1) {
                const bool swap = n1 < n0;
                stack[sp++] = swap ? c0 : c1;
                node = swap ? c1 : c0;
            } else if (h0) node = c0;
            else if (h1) node = c1;
            else {
                if (sp == 0) break;
                node = stack[--sp];
            }
        } else {
            // leaf: ~node = first_slot * 16 + (count - 1)
            const uint32_t code = (uint32_t)(~node);
            const uint32_t first = code >> 4, cnt = (code & 15u) + 1u;
            for (uint32_t s = first; s < first + cnt; s++) {
                const uint32_t obj = __ldg(sc.slot_obj + s);
                if (obj & RECT_BIT) test_rect<R>(sc.geom + 4 * (size_t)s, s, obj, o, d, self, best);
                else test_sphere<R>(sc.geom[4 * (size_t)s], s, obj, o, d, self, onSurf, best);
            }
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    return best;
}

// 256-bit read-only global load (sm_100: LDG.E.256): a 64-byte BVH node is two of these instead of four 128-bit
// loads, a 32-byte leaf record one instead of two — half the L1 wavefronts of the divergent traversal loads.
// `p` must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b)
{
    // evict_last: nodes and leaf records are what L1 should keep; the ray records are read with streaming loads
    asm volatile("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// fp32 BVH traversal (BASELINE config 5).  Differences from the generic nearest_bvh<R>:
//  * while-while: every lane first descends inner nodes until it holds a leaf, then the warp tests leaf primitives
//    together — inner-node and leaf code no longer serialise against each other inside one loop iteration;
//  * slab tests in FMA form, t = lo * (1/d) - o * (1/d);
//  * compact typed leaf records (32 B instead of the 64-B generic slot), two independent 128-bit loads per primitive:
//      sphere                   a = {c.xyz, r}           b = {kind 0, obj, -, -}
//      axis-aligned rectangle   a = {p_K, c_I, c_J, h_I} b = {kind 1+K, obj|RECT, h_J, -}     (as in FastScene)
//      general rectangle        a = {-, -, -, -}         b = {kind 4, obj|RECT, -, -}          -> generic 64-B slot
//  Ties in t across primitives are resolved by object index explicitly (traversal order is arbitrary).
// The test itself, on a record that is already in registers (`slot` is what a hit reports).
__device__ __forceinline__ void test_brec(const SceneView<float>& sc, const float4 a, const float4 b, uint32_t slot, const V3<float>& o, const V3<float>& d,
                                          const V3<float>& inv, uint32_t self, bool onSurf, Hit<float>& best, uint32_t& n_sphere_tests)
{
    const uint32_t kind = __float_as_uint(b.x), obj = __float_as_uint(b.y);
    n_sphere_tests += kind == 0 ? 1u : 0u;          // work counters of ipt_stats (the flops model of bench.py)
    float t;
    bool hit;
    if (kind == 0) {
        const bool selfS = onSurf && obj == self;
        const V3<float> op = mk<float>(o.x - a.x, o.y - a.y, o.z - a.z);
        const float bb = dot(op, d);
        const float delta = fmaf(bb, bb, fmaf(a.w, a.w, -dot(op, op)));
        const float sq = sqrt_(fmaxf(delta, 0.f));
        const float t1 = -bb - sq, t2 = sq - bb;
        t = t1 > (float)IPT_MARGIN ? t1 : t2;
        t = selfS ? -2.f * bb : t;
        hit = (delta >= 0.f || selfS) && t > (float)IPT_MARGIN;
    } else if (kind <= 3) {
        const bool kx = kind == 1, kz = kind == 3;
        const float ok = kx ? o.x : (kz ? o.z : o.y), ik = kx ? inv.x : (kz ? inv.z : inv.y);
        const float oi = kx ? o.y : o.x, di = kx ? d.y : d.x, oj = kz ? o.y : o.z, dj = kz ? d.y : d.z;
        t = (a.x - ok) * ik;
        const float ei = fabsf(fmaf(di, t, oi) - a.y), ej = fabsf(fmaf(dj, t, oj) - a.z);
        hit = t > (float)IPT_MARGIN && ei <= a.w && ej <= b.z && obj != self;
    } else {
        test_rect<float>(sc.geom + 4 * (size_t)slot, slot, obj, o, d, self, best);
        return;
    }
    hit = hit && (t < best.t || (t == best.t && obj < best.obj));
    best.t = hit ? t : best.t;
    best.slot = hit ? slot : best.slot;
    best.obj = hit ? obj : best.obj;
}
__device__ __forceinline__ void test_bslot(const SceneView<float>& sc, uint32_t slot, const V3<float>& o, const V3<float>& d,
                                           const V3<float>& inv, uint32_t self, bool onSurf, Hit<float>& best, uint32_t& n_sphere_tests)
{
    float4 a, b;
    ldg256(sc.bslot + 2 * (size_t)slot, a, b);
    test_brec(sc, a, b, slot, o, d, inv, self, onSurf, best, n_sphere_tests);
}

__device__ __forceinline__ Hit<float> nearest_bvh_f32(const SceneView<float>& sc, const float4* top, uint32_t n_top, const V3<float> o,
                                                      const V3<float> d, uint32_t self, bool onSurf)
{
    Hit<float> best;
    best.t = (float)IPT_INF; best.slot = NO_OBJECT; best.obj = NO_OBJECT;
    const V3<float> inv = mk<float>(1.f / d.x, 1.f / d.y, 1.f / d.z);        // primitive tests: +-inf for zero components (Plane.cu:55)
    // box tests use t = lo * (1/d) - o * (1/d); an infinite 1/d would turn that into inf - inf, so components below
    // 1e-18 are replaced by +-1e-18 there (the boxes are padded by 2e-3, far more than the error this introduces)
    const float tiny = 1e-18f;
    const V3<float> bi = mk<float>(1.f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)), 1.f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
                                   1.f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
    const V3<float> oi = mk<float>(o.x * bi.x, o.y * bi.y, o.z * bi.z);
    int stack[64];
    int sp = 0;
    int node = 0;
    uint32_t n_sph = 0;                                  // the fused kernel does not report work counters
    const float slack = 1.0000004f;
    for (;;) {
        while (node >= 0) {
            float4 a, b, c, e;
            if ((uint32_t)node < n_top) { const float4* p = top + 4 * node; a = p[0]; b = p[1]; c = p[2]; e = p[3]; }
            else { const float4* p = sc.nodes + 4 * (size_t)node; ldg256(p, a, b); ldg256(p + 2, c, e); }
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
            if (h0 && h1) {
                const bool swap = n1 < n0;
                stack[sp++] = swap ? c0 : c1;
                node = swap ? c1 : c0;
            } else if (h0 || h1) {
                node = h0 ? c0 : c1;
            } else {
                if (sp == 0) return best;
                node = stack[--sp];
            }
        }
        // leaf: ~node = first_slot * 16 + (count - 1)
        const uint32_t code = (uint32_t)(~node);
        const uint32_t first = code >> 4, cnt = (code & 15u) + 1u;
        for (uint32_t s = first; s < first + cnt; s++) test_bslot(sc, s, o, d, inv, self, onSurf, best, n_sph);
        if (sp == 0) return best;
        node = stack[--sp];
    }
}

// ---------------------------------------------------------------------------------------------- scatter
// AObject.hpp:30-33: in - (n * (in.n)) * 2
template <typename R> __device__ __forceinline__ V3<R> reflect_dir(V3<R> in, V3<R> n) { return in - n * dot(in, n) * (R)2; }

// AObject.hpp:35-45: cube-normalised direction, flipped into n's hemisphere.  Draws 0,1,2 of the event's block.
template <typename R> __device__ __forceinline__ V3<R> diffuse_dir(V3<R> n, uint4 rnd)
{
    V3<R> v = normalize(mk<R>(s24<R>(rnd.x), s24<R>(rnd.y), s24<R>(rnd.z)));   // never (0,0,0): s24 is an odd multiple of 2^-24
    return dot(v, n) < (R)0 ? -v : v;
}

// "some component is not zero" (+0 and -0 are zero, NaN is not): for fp32 one OR of the three bit patterns and one masked test
// instead of three compares
__device__ __forceinline__ bool any_nonzero(V3<float> v) { return ((__float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z)) & 0x7FFFFFFFu) != 0u; }
__device__ __forceinline__ bool any_nonzero(V3<double> v) { return v.x != 0.0 || v.y != 0.0 || v.z != 0.0; }

// AObject.hpp:47-60: eta = 1/1.5 in both directions, not normalised; returns false on total internal reflection.
template <typename R> __device__ __forceinline__ bool refract_dir(V3<R> in, V3<R> n, V3<R>& out)
{
    const R index = (R)(1.0 / 1.5);
    const R cosI = fabs(dot(n, in));
    const R sin2 = (index * index) * ((R)1 - cosI * cosI);
    if (sin2 > (R)1) return false;
    const R cosT = sqrt_((R)1 - sin2);
    out = in * index + n * (index * cosI - cosT);
    return any_nonzero(out);                                     // AObject.hpp:117 compares the result with Vec3()
}

// Result of shading one hit: up to two continuation rays (the reference's RayData, RayData.hpp:12-28).
template <typename R> struct Spawn {
    V3<R> d0, d1;
    R w0, w1;
    bool has0, has1;
    bool teleport;   // unknown material: the continuation is the reference's default RayData ray, origin = direction = 0
};

// Sphere.cu:41-56, Plane.cu:70-84, AObject.hpp:83-135.  `g0` is the first geometry vector of the hit slot.
template <typename R>
__device__ __forceinline__ Spawn<R> scatter(bool isRect, const R4<R> g0, int reflection, V3<R> P, V3<R> in, uint32_t depth, uint4 rnd)
{
    V3<R> raw, n;
    if (isRect) {
        const V3<R> pn = xyz(g0);
        n = dot(in, pn) < (R)0 ? pn : -pn;      // Plane.cu:73: opposes the incoming ray
        raw = n;                                // Plane.cu:79
    } else {
        raw = normalize(P - xyz(g0));           // Sphere.cu:44
        n = dot(in, raw) < (R)0 ? -raw : raw;   // Sphere.cu:45: points ALONG the incoming ray (into the surface)
    }
    Spawn<R> s;
    s.has0 = true; s.has1 = false; s.w0 = (R)1; s.w1 = (R)0; s.teleport = false;
    s.d1 = mk<R>(0, 0, 0);
    if (reflection == 0) {                      // AObject.hpp:104-108
        s.d0 = diffuse_dir(n, rnd);
    } else if (reflection == 1) {               // AObject.hpp:83-102
        const V3<R> spec = reflect_dir(in, n);
        const V3<R> diff = diffuse_dir(n, rnd);
        if (depth < 2) { s.d0 = spec; s.w0 = (R)0.92; s.d1 = diff; s.w1 = (R)0.08; s.has1 = true; }
        else s.d0 = (u23<R>(rnd.w) > (R)0.9) ? diff : spec;
    } else if (reflection == 2) {               // AObject.hpp:110-135
        const V3<R> spec = reflect_dir(in, n);
        V3<R> refr;
        if (!refract_dir(in, raw, refr)) s.d0 = spec;
        else if (depth < 2) { s.d0 = refr; s.w0 = (R)0.95; s.d1 = spec; s.w1 = (R)0.05; s.has1 = true; }
        else s.d0 = (u23<R>(rnd.w) > (R)0.95) ? spec : refr;
    } else {
        // "Uknown reflection type" (Sphere.cu:52-55): RayData{} = ray (0,0,0)->(0,0,0) with power 0.  firstLayer and
        // secondLayer multiply by that power (nothing to trace), but deepLayers ignores it (Renderer.cu:208-209): from
        // depth 2 on the path continues, at full weight, from the origin with a zero direction (it can only hit a
        // sphere that contains the origin).
        s.has0 = depth >= 2;
        s.teleport = true;
        s.d0 = mk<R>(0, 0, 0);
    }
    return s;
}


// ============================================================================================== fp32 fast path
// Brute-force scenes in fp32 (the product path for the three shipped scenes and BASELINE config 4): the scene is
// re-laid out by ipt_ctx_set_scene into typed lists so that each list's test is a short branch-free sequence:
//   spheres                  float4 {c.xyz, r}
//   axis-aligned rectangles  per axis K (normal = +-e_K): float4 {p_K, c_I, c_J, h_I}, float4 {h_J, obj bits, 0, 0}
//                            with I, J the two other axes in increasing order — t = (p_K - o_K) / d_K, inside iff
//                            |P_I - c_I| <= h_I and |P_J - c_J| <= h_J.  Same test as Plane.cu:47-100 for n = +-e_K
//                            (all rectangles of spheres/mirrors/maze.json are of this kind).
//   general rectangles       the 4 x float4 slot of the generic path
// Every list is kept in object (JSON) order and updated with a strict '<', so among primitives of one list the lowest
// object index wins exact ties in t, as in Renderer.cu:235 (coplanar overlapping rectangles share a list; ties across
// lists need a ray through a geometric edge).  A hit on an axis-aligned rectangle sets the hit point's K coordinate to
// p_K exactly, so the next ray's test against that plane (and any coplanar one) gives t = 0 and fails `t > 1e-4`
// exactly like the reference's fp64 arithmetic does — no per-primitive self comparison is needed in these lists.
// Blob layout in 16-byte words: {n_sph, n_x, n_y, n_z} {n_gen, n_objects, 0, 0} sph[] sph_obj[] x[] y[] z[] gen[] gen_obj[] mat[]
struct FastScene {
    const float4* sph; const uint32_t* sph_obj; uint32_t n_sph;
    const float4* axs;                       // the three axis lists back to back; list K starts at record ax0[K]
    uint32_t ax0_x, ax0_y, ax0_z, n_x, n_y, n_z;
    uint32_t g_x, g_y, g_z;                  // leading records of each list that share one plane (fast_axis_group), 0 = none
    const R4<float>* gen; const uint32_t* gen_obj; uint32_t n_gen;
    const float4* mat;
    bool box_pairs;                          // see fast_axis_pair
    float blo_x, blo_y, blo_z, bhi_x, bhi_y, bhi_z;   // plane coordinates of the lower / upper wall per axis (kernel parameters:
                                             // the 'origin between the walls' test compares against the constant bank)
};
__host__ __device__ inline uint32_t fast_blob_words(uint32_t n_sph, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t n_gen, uint32_t n_obj)
{
    return 2 + n_sph + (n_sph + 3) / 4 + 2 * (nx + ny + nz) + 4 * n_gen + (n_gen + 3) / 4 + 2 * n_obj;
}
// The list sizes come in as kernel parameters (uniform registers / constant bank), not from the blob in shared memory.
struct FastHeader {
    uint32_t n_sph, n_x, n_y, n_z, n_gen, n_obj;
    uint32_t box_pairs;     // box room whose two rectangles per axis are stored lower plane first (fast_axis_pair)
    uint32_t any_unknown;   // some object has an unknown reflection value: the typed-list kernel is not used (the generic one knows the case)
    float box_lo[3], box_hi[3];   // their plane coordinates
    uint32_t group[3];            // per axis list: its first group[K] records lie on one plane (0: no such group)
    uint32_t off_sphobj, off_axs, off_gen, off_genobj, off_mat;   // 16-byte word offsets of the lists inside the blob
};
__host__ __device__ inline FastHeader fast_header(uint32_t n_sph, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t n_gen, uint32_t n_obj)
{
    FastHeader h;
    h.n_sph = n_sph; h.n_x = nx; h.n_y = ny; h.n_z = nz; h.n_gen = n_gen; h.n_obj = n_obj; h.box_pairs = 0; h.any_unknown = 1;
    for (int k = 0; k < 3; k++) { h.box_lo[k] = 0.f; h.box_hi[k] = 0.f; h.group[k] = 0; }
    h.off_sphobj = 2 + n_sph;
    h.off_axs = h.off_sphobj + (n_sph + 3) / 4;
    h.off_gen = h.off_axs + 2 * (nx + ny + nz);
    h.off_genobj = h.off_gen + 4 * n_gen;
    h.off_mat = h.off_genobj + (n_gen + 3) / 4;
    return h;
}
__device__ __forceinline__ FastScene fast_view(const uint4* blob, const FastHeader& hd)
{
    FastScene f;
    f.n_sph = hd.n_sph; f.n_x = hd.n_x; f.n_y = hd.n_y; f.n_z = hd.n_z; f.n_gen = hd.n_gen; f.box_pairs = hd.box_pairs != 0;
    f.blo_x = hd.box_lo[0]; f.blo_y = hd.box_lo[1]; f.blo_z = hd.box_lo[2]; f.bhi_x = hd.box_hi[0]; f.bhi_y = hd.box_hi[1]; f.bhi_z = hd.box_hi[2];
    f.sph = reinterpret_cast<const float4*>(blob + 2);
    f.sph_obj = reinterpret_cast<const uint32_t*>(blob + hd.off_sphobj);
    f.axs = reinterpret_cast<const float4*>(blob + hd.off_axs);
    f.ax0_x = 0; f.ax0_y = hd.n_x; f.ax0_z = hd.n_x + hd.n_y;
    f.g_x = hd.group[0]; f.g_y = hd.group[1]; f.g_z = hd.group[2];
    f.gen = reinterpret_cast<const R4<float>*>(blob + hd.off_gen);
    f.gen_obj = reinterpret_cast<const uint32_t*>(blob + hd.off_genobj);
    f.mat = reinterpret_cast<const float4*>(blob + hd.off_mat);
    return f;
}

// hit code: kind[28:32) (0 sphere, 1..3 axis-aligned rectangle with normal e_(kind-1), 4 general rectangle) | index in
// the sphere list / the concatenated axis lists / the general list
struct FastHit { float t; uint32_t code; };

__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int K> __device__ __forceinline__ float comp(const V3<float>& v) { return K == 0 ? v.x : (K == 1 ? v.y : v.z); }

template <int K>
__device__ __forceinline__ void fast_axis_list(const float4* __restrict__ axs, uint32_t first, uint32_t n, const V3<float>& o,
                                               const V3<float>& d, float inv_dk, FastHit& best)
{
    constexpr int I = K == 0 ? 1 : 0, J = K == 2 ? 1 : 2;
    const float ok = comp<K>(o), oi = comp<I>(o), oj = comp<J>(o), di = comp<I>(d), dj = comp<J>(d);
    const float4* rec = axs + 2 * first;
    for (uint32_t s = first; s < first + n; s += 2, rec += 4) {   // n is even: lists are padded with a never-hit record
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const float4 a = rec[2 * u];
            const float hj = rec[2 * u + 1].x;
            const float t = (a.x - ok) * inv_dk;                 // d_K == 0: +-inf or NaN, both rejected below (Plane.cu:55)
            const float ei = fabsf(fmaf(di, t, oi) - a.y), ej = fabsf(fmaf(dj, t, oj) - a.z);
            const bool hit = t > (float)IPT_MARGIN && t < best.t && ei <= a.w && ej <= hj;
            best.t = hit ? t : best.t;
            best.code = hit ? (((uint32_t)(K + 1) << 28) + (s + u)) : best.code;   // '+', not '|': one predicated add with an immediate per record
        }
    }
}

// The records [first, first + g) of a list lie on ONE plane: t and the hit point are computed once, a record costs its two edge
// tests.  Walked backwards with an unconditional overwrite, so the lowest record that contains the point wins - what the loop
// above does for equal t (after the first hit `t < best.t` is false for the rest of the plane); same arithmetic per record.
template <int K>
__device__ __forceinline__ void fast_axis_group(const float4* __restrict__ axs, uint32_t first, uint32_t g, const V3<float>& o,
                                                const V3<float>& d, float inv_dk, FastHit& best)
{
    constexpr int I = K == 0 ? 1 : 0, J = K == 2 ? 1 : 2;
    const float ok = comp<K>(o), oi = comp<I>(o), oj = comp<J>(o), di = comp<I>(d), dj = comp<J>(d);
    const float t = (axs[2 * first].x - ok) * inv_dk;
    const float pi = fmaf(di, t, oi), pj = fmaf(dj, t, oj);
    uint32_t found = NO_OBJECT;
    const float4* rec = axs + 2 * (first + g);
#pragma unroll 4
    for (uint32_t s = first + g; s-- > first;) {
        rec -= 2;
        const float4 a = rec[0];
        const float hj = rec[1].x;
        found = (fabsf(pi - a.y) <= a.w && fabsf(pj - a.z) <= hj) ? s : found;
    }
    const bool hit = t > (float)IPT_MARGIN && t < best.t && found != NO_OBJECT;
    best.t = hit ? t : best.t;
    best.code = hit ? (((uint32_t)(K + 1) << 28) + found) : best.code;
}

// The same for a list whose position and (even) length are compile-time constants: straight-line code, record loads
// at constant offsets, hit codes as immediates.
template <int K, int FIRST_REC, int N>
__device__ __forceinline__ void fast_axis_fixed(const float4* __restrict__ axs, const V3<float>& o, const V3<float>& d, float inv_dk, FastHit& best)
{
    constexpr int I = K == 0 ? 1 : 0, J = K == 2 ? 1 : 2;
    const float ok = comp<K>(o), oi = comp<I>(o), oj = comp<J>(o), di = comp<I>(d), dj = comp<J>(d);
#pragma unroll
    for (int u = 0; u < N; u++) {
        const float4 a = axs[2 * (FIRST_REC + u)];
        const float hj = axs[2 * (FIRST_REC + u) + 1].x;
        const float t = (a.x - ok) * inv_dk;
        const float ei = fabsf(fmaf(di, t, oi) - a.y), ej = fabsf(fmaf(dj, t, oj) - a.z);
        const bool hit = t > (float)IPT_MARGIN && t < best.t && ei <= a.w && ej <= hj;
        best.t = hit ? t : best.t;
        best.code = hit ? (((uint32_t)(K + 1) << 28) | (uint32_t)(FIRST_REC + u)) : best.code;
    }
}

// The two parallel walls of a box room along axis K, for a ray whose origin lies between them (FastHeader::box_pairs: the
// host has stored the lower wall of every pair first).  The wall the ray travels away from has t = (p - o_K) / d_K <= 0
// (or -inf / NaN for d_K = 0) and fails `t > 1e-4` in the scan of both, so testing only the wall ahead - its record picked
// by the sign of d_K - gives the bit-identical result with half the tests.
template <int K, int FIRST_REC>
__device__ __forceinline__ void fast_axis_pair(const float4* __restrict__ axs, const V3<float>& o, const V3<float>& d, float inv_dk, FastHit& best)
{
    constexpr int I = K == 0 ? 1 : 0, J = K == 2 ? 1 : 2;
    const float ok = comp<K>(o), oi = comp<I>(o), oj = comp<J>(o), di = comp<I>(d), dj = comp<J>(d);
    const uint32_t up = comp<K>(d) > 0.f ? 1u : 0u;
    const float4* rec = axs + 2 * FIRST_REC + 2 * up;
    const float4 a = rec[0];
    const float hj = rec[1].x;
    const float t = (a.x - ok) * inv_dk;
    const float ei = fabsf(fmaf(di, t, oi) - a.y), ej = fabsf(fmaf(dj, t, oj) - a.z);
    const bool hit = t > (float)IPT_MARGIN && t < best.t && ei <= a.w && ej <= hj;
    best.t = hit ? t : best.t;
    best.code = hit ? ((((uint32_t)(K + 1) << 28) | (uint32_t)FIRST_REC) + up) : best.code;
}

// Renderer.cu:227-243 for the fp32 brute-force layout; self-hit rule as in SelfRule<float>.  `self` is the hit CODE of
// the surface the ray starts on (the fast kernel's queues are private to it), NO_OBJECT for camera rays.
__device__ __forceinline__ void fast_sphere(const float4 sp, uint32_t s, uint32_t self_sphere, const V3<float>& o, const V3<float>& d, FastHit& best)
{
    const bool selfS = s == self_sphere;                         // start point lies ON this sphere: exact second root -2b
    const V3<float> op = mk<float>(o.x - sp.x, o.y - sp.y, o.z - sp.z);
    const float b = dot(op, d);
    const float delta = fmaf(b, b, fmaf(sp.w, sp.w, -dot(op, op)));           // b*b - op.op + r*r   (Sphere.cu:31)
    const float sq = sqrt_fast(fmaxf(delta, 0.f));
    const float t1 = -b - sq, t2 = sq - b;
    float t = t1 > (float)IPT_MARGIN ? t1 : t2;
    t = selfS ? -2.f * b : t;
    const bool hit = (delta >= 0.f || selfS) && t > (float)IPT_MARGIN && t < best.t;
    best.t = hit ? t : best.t;
    best.code = hit ? s : best.code;
}

// SHAPE: 0 = list lengths at run time (loops); -1 = the same with a coplanar group in front of some list (fast_axis_group);
// k = 1..5 = "box room": exactly two rectangles per axis list (a closed
// axis-aligned box, padded records included), no general rectangles and k-1 spheres.  Those lengths being compile-time
// constants, the whole scan is straight-line code (spheres.json and every Cornell-box-like scene).
__host__ __device__ inline int fast_shape(uint32_t n_sph, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t n_gen)
{
    return (nx == 2 && ny == 2 && nz == 2 && n_gen == 0 && n_sph <= 4) ? (int)n_sph + 1 : 0;
}

template <int SHAPE>
__device__ __forceinline__ FastHit nearest_fast(const FastScene& f, const V3<float>& o, const V3<float>& d, uint32_t self, bool onSurf)
{
    FastHit best;
    best.t = (float)IPT_INF; best.code = NO_OBJECT;
    const uint32_t self_sphere = onSurf ? self : NO_OBJECT;      // sphere codes are plain list indices (kind 0)
    if (SHAPE > 0) {
#pragma unroll
        for (int s = 0; s < SHAPE - 1; s++) fast_sphere(f.sph[s], (uint32_t)s, self_sphere, o, d, best);
        // origin between the walls of every pair (on a wall counts): only the wall ahead can be hit (fast_axis_pair)
        const bool between = f.box_pairs && o.x >= f.blo_x && o.x <= f.bhi_x && o.y >= f.blo_y && o.y <= f.bhi_y && o.z >= f.blo_z && o.z <= f.bhi_z;
        if (between) {
            fast_axis_pair<0, 0>(f.axs, o, d, rcp_fast(d.x), best);
            fast_axis_pair<1, 2>(f.axs, o, d, rcp_fast(d.y), best);
            fast_axis_pair<2, 4>(f.axs, o, d, rcp_fast(d.z), best);
        } else {
            fast_axis_fixed<0, 0, 2>(f.axs, o, d, rcp_fast(d.x), best);
            fast_axis_fixed<1, 2, 2>(f.axs, o, d, rcp_fast(d.y), best);
            fast_axis_fixed<2, 4, 2>(f.axs, o, d, rcp_fast(d.z), best);
        }
        return best;
    }
#pragma unroll 2
    for (uint32_t s = 0; s < f.n_sph; s++) fast_sphere(f.sph[s], s, self_sphere, o, d, best);
    if (SHAPE < 0) {    // some list starts with a coplanar group (its own instantiation: carried by every list scene, the extra
                        // branches and loop bounds cost mirrors.json 5 %)
        const float ix = rcp_fast(d.x), iy = rcp_fast(d.y), iz = rcp_fast(d.z);
        if (f.g_x) fast_axis_group<0>(f.axs, f.ax0_x, f.g_x, o, d, ix, best);
        fast_axis_list<0>(f.axs, f.ax0_x + f.g_x, f.n_x - f.g_x, o, d, ix, best);
        if (f.g_y) fast_axis_group<1>(f.axs, f.ax0_y, f.g_y, o, d, iy, best);
        fast_axis_list<1>(f.axs, f.ax0_y + f.g_y, f.n_y - f.g_y, o, d, iy, best);
        if (f.g_z) fast_axis_group<2>(f.axs, f.ax0_z, f.g_z, o, d, iz, best);
        fast_axis_list<2>(f.axs, f.ax0_z + f.g_z, f.n_z - f.g_z, o, d, iz, best);
    } else {
        fast_axis_list<0>(f.axs, f.ax0_x, f.n_x, o, d, rcp_fast(d.x), best);
        fast_axis_list<1>(f.axs, f.ax0_y, f.n_y, o, d, rcp_fast(d.y), best);
        fast_axis_list<2>(f.axs, f.ax0_z, f.n_z, o, d, rcp_fast(d.z), best);
    }
    for (uint32_t s = 0; s < f.n_gen; s++) {
        Hit<float> h;
        h.t = best.t; h.obj = NO_OBJECT; h.slot = NO_OBJECT;     // obj = max: strict '<' within this (ordered) list
        test_rect<float>(f.gen + 4 * s, s, (4u << 28) | s, o, d, self, h);   // compared with `self` only: codes on both sides
        if (h.slot != NO_OBJECT) { best.t = h.t; best.code = (4u << 28) | s; }
    }
    return best;
}

// Object id (with RECT_BIT) of a hit code, and the hit point with the axis-aligned snap described above.
__device__ __forceinline__ uint32_t fast_hit_object(const FastScene& f, uint32_t code)
{
    const uint32_t kind = code >> 28, idx = code & 0x0FFFFFFFu;
    if (kind == 0) return f.sph_obj[idx];
    if (kind == 4) return f.gen_obj[idx];
    return __float_as_uint(f.axs[2 * idx + 1].y);
}
__device__ __forceinline__ V3<float> fast_hit_point(const FastScene& f, uint32_t code, const V3<float>& o, const V3<float>& d, float t)
{
    V3<float> P = mk<float>(fmaf(d.x, t, o.x), fmaf(d.y, t, o.y), fmaf(d.z, t, o.z));
    const uint32_t kind = code >> 28;
    if (kind >= 1 && kind <= 3) {
        const float pk = f.axs[2 * (code & 0x0FFFFFFFu)].x;
        P.x = kind == 1 ? pk : P.x; P.y = kind == 2 ? pk : P.y; P.z = kind == 3 ? pk : P.z;
    }
    return P;
}

// Branch-free variant of scatter<float>: all candidate directions are computed, the material selects.
// NO_GEN: the scene has no general (non axis-aligned) rectangle, so hit kind 4 cannot occur.
template <bool NO_GEN = false>
__device__ __forceinline__ Spawn<float> scatter_fast(const FastScene& f, uint32_t code, int reflection, V3<float> P, V3<float> in,
                                                     uint32_t depth, uint4 rnd)
{
    const uint32_t kind = code >> 28, idx = code & 0x0FFFFFFFu;
    V3<float> raw, n;
    {   // axis-aligned rectangle (most hits; computed for every lane, a few selects): normal +-e_K - Plane.cu:73's test
        // n.d < 0 is the sign of the ray's K component
        const float ik = kind == 1 ? in.x : (kind == 2 ? in.y : in.z);
        const float sg = ik < 0.f ? 1.f : -1.f;
        n = mk<float>(kind == 1 ? sg : 0.f, kind == 2 ? sg : 0.f, kind == 3 ? sg : 0.f);
        raw = n;                                                              // Plane.cu:79
    }
    if (kind == 0) {
        const float4 sp = f.sph[idx];
        raw = normalize(mk<float>(P.x - sp.x, P.y - sp.y, P.z - sp.z));       // Sphere.cu:44
        n = dot(in, raw) < 0.f ? -raw : raw;                                  // Sphere.cu:45
    } else if (!NO_GEN && kind == 4) {
        const V3<float> pn = xyz(f.gen[4 * idx]);
        n = dot(in, pn) < 0.f ? pn : -pn;                                     // Plane.cu:73
        raw = n;
    }
    const V3<float> diff = diffuse_dir(n, rnd);                               // AObject.hpp:35-45
    const V3<float> spec = reflect_dir(in, n);                                // AObject.hpp:30-33
    V3<float> refr = spec;
    const bool refr_ok = refract_dir(in, raw, refr);                          // AObject.hpp:47-60
    const float u = u23<float>(rnd.w);
    const bool early = depth < 2;
    const bool isSpec = reflection == 1, isRefr = reflection == 2;
    // main ray: diffuse -> diff; specular -> spec (depth<2, or u <= 0.9), else diff; refractive -> spec on TIR or
    // (depth>=2 and u > 0.95), else refr
    const bool pickSpec = (isSpec && (early || !(u > 0.9f))) || (isRefr && (!refr_ok || (!early && u > 0.95f)));
    const bool pickRefr = isRefr && !pickSpec;
    Spawn<float> s;
    s.d0.x = pickSpec ? spec.x : (pickRefr ? refr.x : diff.x);
    s.d0.y = pickSpec ? spec.y : (pickRefr ? refr.y : diff.y);
    s.d0.z = pickSpec ? spec.z : (pickRefr ? refr.z : diff.z);
    // no unknown materials here (the 'teleport' ray of scatter<R>): a scene that has one is rendered by the generic kernel
    // (FastHeader::any_unknown, ipt_render.cu) - as a flag inside this kernel the case cost eight selects per bounce
    s.teleport = false; s.has0 = true;
    s.has1 = early && (isSpec || (isRefr && refr_ok));                        // AObject.hpp:91-94, :122-125
    s.w0 = s.has1 ? (isSpec ? 0.92f : 0.95f) : 1.f;
    s.w1 = isSpec ? 0.08f : 0.05f;
    s.d1.x = isSpec ? diff.x : spec.x; s.d1.y = isSpec ? diff.y : spec.y; s.d1.z = isSpec ? diff.z : spec.z;
    return s;
}

}  // namespace ipt
