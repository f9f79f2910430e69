// TEST INFRASTRUCTURE ONLY — the CPU oracle. Nothing in the product path may include, link or call this.
//
// CPU restatement (FP64, scalar, one pixel at a time) of the per-pixel radiance routine of
// AdamStudies-PWR/Improved-Path-Tracer.  Every function cites the reference file:line it follows
// (paths relative to /root/reference).  Written from the algorithm, not copied: own
// vector type, own object records, RNG behind an interface with two interchangeable streams.
//
// PINNING: with OR_RNG_REFERENCE this file reproduces oracle/_ref/libref_host.so (the reference's own
// Renderer.cu compiled for the host) to <= 1e-12 on every pixel of spheres/mirrors/maze
// (tests/test_oracle_pin.py; function-level KATs in the same file).  The reference ships no tests and no
// golden vectors of its own, so outputs of the reference run here are the anchor.
//
// "Clean semantics": the reference reads objectEmissions[maxDepth-2], one element past the end, on every
// full-length path (Renderer.cu:216-219).  Here — as in oracle/_ref, whose operator new[] zero-fills — that slot is 0.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_runtime.h>        // leaves __host__/__device__ empty for g++
#define QUALIFIERS static inline
#include <curand_kernel.h>       // NVIDIA's XORWOW (toolkit header, not reference code): host-callable this way

#include "restate.h"

namespace {

// ------------------------------------------------------------------------------------------------ vector
// Vec3.hpp:18-84 (value semantics only)
struct V {
    double x = 0, y = 0, z = 0;
    V() {}
    V(double a, double b, double c) : x(a), y(b), z(c) {}
    explicit V(const double* p) : x(p[0]), y(p[1]), z(p[2]) {}
};
inline V operator+(const V& a, const V& b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }      // Vec3.hpp:53-56
inline V operator-(const V& a, const V& b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }      // :58-61
inline V operator*(const V& a, double s) { return V(a.x * s, a.y * s, a.z * s); }              // :63-66
inline double dot(const V& a, const V& b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }  // :27-30
inline V mult(const V& a, const V& b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }           // :32-35
inline V cross(const V& a, const V& b)                                                          // :69-72
{
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline V norm(const V& a) { return a * (1 / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); }   // :48-51
inline double dist(const V& a, const V& b)                                                      // :37-40 (pow(x,2) == x*x)
{
    return std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}
inline bool eq(const V& a, const V& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }      // :74-77

struct RayT { V o, d; };

const double MARGIN = 1e-4;           // scene/cuda/objects/Constants.hpp:8
const double INF = 1e20;              // Renderer.cu:29
const float FOV_SCALE = 0.0009f;      // Renderer.cu:27 (a float constant, promoted to double where used)
const double VIEWPORT_DISTANCE = 140; // Renderer.cu:28
const uint32_t BLOCK = 22;            // renderer/Constants.hpp:11

// ------------------------------------------------------------------------------------------------ RNG
// Philox4x32-R (Salmon et al., SC'11) — the counter-based generator of the B200 path, which uses R = 7 rounds
// (OR_PHILOX_ROUNDS; Random123's known answers for R = 7 and R = 10 pin this function, tests/test_oracle_pin.py).
const int OR_PHILOX_ROUNDS = 7;
inline void philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4], int rounds = OR_PHILOX_ROUNDS)
{
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
    for (int r = 0; r < rounds; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// Random reals of the counter stream, exactly representable in fp32 and fp64 (the B200 kernels compute the same values):
inline double s24(uint32_t x) { return ((double)((int)(x >> 8) - 8388608) + 0.5) * (1.0 / 8388608.0); }  // (-1,1), odd multiple of 2^-24
inline double u23(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }                    // (0,1)

const uint32_t NODE_CAMERA = 0xFFFFu;
const uint32_t CTR_TAG = 0x49505442u;  // "IPTB"

struct Rng {
    int mode = OR_RNG_REFERENCE;
    curandState st;                    // reference stream (Renderer.cu:95-97)
    uint32_t key[2] = {0, 0}, pixel = 0, sample = 0, blk[4] = {0, 0, 0, 0};

    void node(uint32_t lane, uint32_t depth) { block((lane << 8) | depth); }
    void block(uint32_t node_id)
    {
        if (mode != OR_RNG_COUNTER) return;
        const uint32_t c[4] = {pixel, sample, node_id, CTR_TAG};
        philox(c, key, blk);
    }
    // CudaUtils.hpp:14-17 one_one = 2u-1 over curand_uniform_double (two 32-bit draws each)
    double one_one_ref() { return (curand_uniform_double(&st) * 2) - 1; }
    // Renderer.cu:133-134: xFactor first, then zFactor
    void jitter(double& jx, double& jz)
    {
        if (mode == OR_RNG_COUNTER) { block(NODE_CAMERA); jx = s24(blk[0]); jz = s24(blk[1]); }
        else { jx = one_one_ref(); jz = one_one_ref(); }
    }
    // AObject.hpp:40 Vec3(one_one, one_one, one_one): g++ evaluates the arguments right to left (zz, yy, xx),
    // which is what oracle/_ref does; the counter stream assigns draws 0,1,2 to xx,yy,zz by definition.
    V dir3()
    {
        V v;
        if (mode == OR_RNG_COUNTER) { v.x = s24(blk[0]); v.y = s24(blk[1]); v.z = s24(blk[2]); }
        else { v.z = one_one_ref(); v.y = one_one_ref(); v.x = one_one_ref(); }
        return v;
    }
    // AObject.hpp:94,127 curand_uniform_double for the stochastic lobe pick; counter stream: draw 3 of the block
    double choice() { return mode == OR_RNG_COUNTER ? u23(blk[3]) : curand_uniform_double(&st); }
};

// ------------------------------------------------------------------------------------------------ objects
struct Obj {
    int type, reflection;
    double radius;
    V position, emission, color;
    // rectangle ("Plane") state computed by its constructor, Plane.cu:32-45
    V planeVector, bottomLeft, bottomRight, topLeft, topRight;
    double distanceHorizontal = 0, distanceVertical = 0;
};

Obj make_obj(const or_object& s)
{
    Obj o;
    o.type = s.type; o.reflection = s.reflection; o.radius = s.radius;
    o.position = V(s.position); o.emission = V(s.emission); o.color = V(s.color);
    if (s.type == 1) {
        const V north(s.north), east(s.east);
        o.planeVector = norm(cross(north, east));                    // Plane.cu:36
        o.bottomRight = o.position + (north * -1) + east;            // :38
        o.bottomLeft = o.position + (north * -1) + (east * -1);      // :39
        o.topLeft = o.position + north + (east * -1);                // :40
        o.topRight = o.position + north + east;                      // :41
        o.distanceHorizontal = dist(o.bottomLeft, o.bottomRight);    // :43
        o.distanceVertical = dist(o.bottomLeft, o.topLeft);          // :44
    }
    return o;
}

// Sphere.cu:25-39
double sphere_intersect(const Obj& s, const RayT& r)
{
    const V op = r.o - s.position;
    const double b = dot(op, r.d);
    double delta = b * b - dot(op, op) + s.radius * s.radius;
    if (delta < 0) return 0.0;
    delta = std::sqrt(delta);
    double t;
    return ((t = -b - delta) > MARGIN) ? t : (((t = -b + delta) > MARGIN) ? t : 0.0);
}

// Plane.cu:16-26: perpendicular distance from `impact` to the line through `origin` along `border`
double distance_to_border(const V& origin, const V& border, const V& impact)
{
    const V ref = impact - origin;
    const double top = border.x * ref.x + border.y * ref.y + border.z * ref.z;
    const double bottom = border.x * border.x + border.y * border.y + border.z * border.z;
    if (bottom == 0.0) return 0.0;
    const double d = top / bottom;
    return dist(origin + border * d, impact);
}

// Plane.cu:87-100: inside iff the distances to the two opposite edge lines add up to the side length (+-MARGIN)
bool in_bounds(const Obj& p, const V& impact)
{
    double vertical = distance_to_border(p.bottomLeft, norm(p.bottomLeft - p.bottomRight), impact);
    if (p.distanceVertical - vertical < -MARGIN) return false;
    vertical = vertical + distance_to_border(p.topLeft, norm(p.topLeft - p.topRight), impact);
    if (p.distanceVertical - vertical < -MARGIN || p.distanceVertical - vertical > MARGIN) return false;

    double horizontal = distance_to_border(p.bottomLeft, norm(p.bottomLeft - p.topLeft), impact);
    if (p.distanceHorizontal - horizontal < -MARGIN) return false;
    horizontal = horizontal + distance_to_border(p.bottomRight, norm(p.bottomRight - p.topRight), impact);
    if (p.distanceHorizontal - horizontal < -MARGIN || p.distanceHorizontal - horizontal > MARGIN) return false;
    return true;
}

// Plane.cu:47-68
double plane_intersect(const Obj& p, const RayT& r)
{
    const V ref = p.position - r.o;
    const double top = p.planeVector.x * ref.x + p.planeVector.y * ref.y + p.planeVector.z * ref.z;
    const double bottom = p.planeVector.x * r.d.x + p.planeVector.y * r.d.y + p.planeVector.z * r.d.z;
    if (bottom == 0.0) return 0.0;
    const double t = top / bottom;
    if (t <= MARGIN) return 0.0;
    const V impact = r.o + (r.d * t);
    if (!in_bounds(p, impact)) return 0.0;
    return t;
}

inline double intersect(const Obj& o, const RayT& r) { return o.type == 0 ? sphere_intersect(o, r) : plane_intersect(o, r); }

// RayData.hpp:12-28
struct Scatter {
    RayT ray, second;
    double power = 0.0, secondPower = 0.0;
    bool useSecond = false;
};

// AObject.hpp:30-33  (normal * dot) * 2, in that order
V reflect_dir(const V& in, const V& n) { return in - n * dot(in, n) * 2; }

// AObject.hpp:35-45: three uniforms in (-1,1), normalised (a cube-normalised direction), flipped to n's side
V diffuse_dir(const V& n, Rng& rng)
{
    V d(0, 0, 0);
    while (eq(d, V(0, 0, 0))) d = rng.dir3();
    d = norm(d);
    return (dot(d, n) < 0) ? d * -1 : d;
}

// AObject.hpp:47-60: eta = 1/1.5 whichever way the ray travels; result NOT normalised
V refract_dir(const V& in, const V& n)
{
    const double index = 1.0 / 1.5;
    const double cosI = std::fabs(dot(n, in));
    const double sin2 = (index * index) * (1.0 - cosI * cosI);
    if (sin2 > 1.0) return V();
    const double cosT = std::sqrt(1.0 - sin2);
    return in * index + n * (index * cosI - cosT);
}

// AObject.hpp:83-102
Scatter handle_specular(const V& P, const V& in, const V& n, Rng& rng, int depth)
{
    const V spec = reflect_dir(in, n);
    const V diff = diffuse_dir(n, rng);
    Scatter s;
    if (depth < 2) { s.ray = {P, spec}; s.power = 0.92; s.second = {P, diff}; s.secondPower = 0.08; s.useSecond = true; return s; }
    if (rng.choice() > 0.9) { s.ray = {P, diff}; s.power = 1.0; }
    else { s.ray = {P, spec}; s.power = 1.0; }
    return s;
}

// AObject.hpp:104-108
Scatter handle_diffuse(const V& P, const V& n, Rng& rng)
{
    Scatter s;
    s.ray = {P, diffuse_dir(n, rng)};
    s.power = 1.0;
    return s;
}

// AObject.hpp:110-135
Scatter handle_refractive(const V& P, const V& in, const V& raw, const V& n, Rng& rng, int depth)
{
    const V spec = reflect_dir(in, n);
    const V refr = refract_dir(in, raw);
    Scatter s;
    if (eq(refr, V())) { s.ray = {P, spec}; s.power = 1.0; return s; }
    if (depth < 2) { s.ray = {P, refr}; s.power = 0.95; s.second = {P, spec}; s.secondPower = 0.05; s.useSecond = true; return s; }
    if (rng.choice() > 0.95) { s.ray = {P, spec}; s.power = 1.0; }
    else { s.ray = {P, refr}; s.power = 1.0; }
    return s;
}

// Sphere.cu:41-56 and Plane.cu:70-84
Scatter scatter(const Obj& o, const V& P, const V& in, Rng& rng, int depth)
{
    V raw, n;
    if (o.type == 0) {
        raw = norm(P - o.position);                                   // Sphere.cu:44
        n = dot(in, raw) < 0 ? raw * -1 : raw;                        // :45  (points along the incoming ray)
    } else {
        n = (dot(in, o.planeVector) < 0 ? o.planeVector * -1 : o.planeVector) * -1;   // Plane.cu:73 (opposes it)
        raw = n;                                                      // Plane.cu:79 passes `normal` twice
    }
    switch (o.reflection) {
        case 1: return handle_specular(P, in, n, rng, depth);
        case 0: return handle_diffuse(P, n, rng);
        case 2: return handle_refractive(P, in, raw, n, rng, depth);
        default: return Scatter();                                     // "Uknown reflection type": zero ray, weight 0
    }
}

// ------------------------------------------------------------------------------------------------ renderer
struct Accel;
struct Ctx {
    std::shared_ptr<const std::vector<Obj>> objs_shared;   // one copy for all the threads of a render (a million objects: 200 MB)
    const std::vector<Obj>& objs_ref() const { return *objs_shared; }
    uint32_t W, H, samples, maxDepth;
    V camO, camD, camX, vecZ;
    uint64_t casts_reference = 0, casts_needed = 0;
    std::shared_ptr<const Accel> accel;   // null: the reference's scan (the default); see Route 3 below
};

struct Hit { int index; double t; };

// ---- SURVEY.md §8(c) Route 3: the same nearest hit for scenes the scan cannot finish (config 5: a million objects).
// NOT in the reference (it has no acceleration structure); an addition of the oracle, opt-in (or_render_accel,
// or_nearest_hit_accel), and checked against the scan below: same object and bit-identical t on random rays and whole
// frames (tests/test_oracle_pin.py).  A median-split tree of padded boxes, deliberately unrelated to the product's SAH
// tree / uniform grid (host/bvh.cpp, host/grid.cpp).  It only decides WHICH objects are tested; the tests themselves are
// intersect() above and the winner is the (t, index)-smallest, i.e. what the scan's strict '<' yields.
// Why culling is exact: a rectangle's accepted hit point lies within MARGIN of the rectangle (Plane.cu:87-100), inside
// its box grown by PAD; a sphere's reported t is never in front of the point where the ray's line enters the sphere,
// also for the shorter-than-unit directions refraction produces (AObject.hpp:59 does not normalise): with l = |d| <= 1,
// x = -(op.d)/l and c = op.op - r^2 > 0 the reported root is f(l x), f(y) = y - sqrt(y^2 - c) = c / (y + sqrt(y^2 - c)),
// the geometric entry in units of d is f(x) / l, and l f(l x) = c / (x + sqrt(x^2 - c / l^2)) >= f(x); c <= 0 means the
// origin is inside the sphere, hence inside its box (entry 0).  So a box is skipped only if the line misses it or enters
// it behind the best t so far (strictly: ties are kept).
struct Accel {
    struct Node { double lo[3], hi[3]; int left, right; uint32_t first, count; };   // leaf: left < 0
    std::vector<Node> nodes;
    std::vector<uint32_t> order;      // leaf ranges point into this
    std::vector<uint32_t> big;        // objects as large as the scene (walls, the light): tested for every ray
};

void object_box(const Obj& o, double lo[3], double hi[3])
{
    const double PAD = 1e-2;
    const V corners[4] = {o.bottomLeft, o.bottomRight, o.topLeft, o.topRight};
    for (int k = 0; k < 3; k++) { lo[k] = 1e300; hi[k] = -1e300; }
    auto grow = [&](const V& p) {
        const double q[3] = {p.x, p.y, p.z};
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], q[k]); hi[k] = std::max(hi[k], q[k]); }
    };
    if (o.type == 0) {
        const double r = std::fabs(o.radius);
        grow(o.position + V(r, r, r)); grow(o.position - V(r, r, r));
    } else {
        for (const V& c : corners) grow(c);
    }
    for (int k = 0; k < 3; k++) {
        const double pad = PAD + 1e-9 * std::max(std::fabs(lo[k]), std::fabs(hi[k]));
        lo[k] -= pad; hi[k] += pad;
    }
}

std::shared_ptr<const Accel> build_accel(const std::vector<Obj>& objs)
{
    auto A = std::make_shared<Accel>();
    const size_t n = objs.size();
    std::vector<double> lo(3 * n), hi(3 * n), ctr(3 * n);
    double slo[3] = {1e300, 1e300, 1e300}, shi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = 0; i < n; i++) {
        object_box(objs[i], &lo[3 * i], &hi[3 * i]);
        for (int k = 0; k < 3; k++) {
            ctr[3 * i + k] = 0.5 * (lo[3 * i + k] + hi[3 * i + k]);
            slo[k] = std::min(slo[k], lo[3 * i + k]); shi[k] = std::max(shi[k], hi[3 * i + k]);
        }
    }
    const double scene_extent = std::max({shi[0] - slo[0], shi[1] - slo[1], shi[2] - slo[2]});
    for (size_t i = 0; i < n; i++) {
        const double ext = std::max({hi[3 * i] - lo[3 * i], hi[3 * i + 1] - lo[3 * i + 1], hi[3 * i + 2] - lo[3 * i + 2]});
        if (!(ext <= 0.1 * scene_extent)) A->big.push_back((uint32_t)i); else A->order.push_back((uint32_t)i);   // (a NaN box is 'big')
    }
    if (A->order.empty()) return A;
    struct Range { uint32_t first, count; int node; };
    std::vector<Range> todo;
    A->nodes.emplace_back();
    todo.push_back({0u, (uint32_t)A->order.size(), 0});
    while (!todo.empty()) {
        const Range rg = todo.back(); todo.pop_back();
        Accel::Node nd;
        double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
        for (int k = 0; k < 3; k++) { nd.lo[k] = 1e300; nd.hi[k] = -1e300; }
        for (uint32_t j = rg.first; j < rg.first + rg.count; j++) {
            const size_t i = A->order[j];
            for (int k = 0; k < 3; k++) {
                nd.lo[k] = std::min(nd.lo[k], lo[3 * i + k]); nd.hi[k] = std::max(nd.hi[k], hi[3 * i + k]);
                clo[k] = std::min(clo[k], ctr[3 * i + k]); chi[k] = std::max(chi[k], ctr[3 * i + k]);
            }
        }
        nd.left = nd.right = -1; nd.first = rg.first; nd.count = rg.count;
        if (rg.count > 4) {
            int ax = 0;
            for (int k = 1; k < 3; k++) if (chi[k] - clo[k] > chi[ax] - clo[ax]) ax = k;
            const uint32_t half = rg.count / 2;
            std::nth_element(A->order.begin() + rg.first, A->order.begin() + rg.first + half, A->order.begin() + rg.first + rg.count,
                             [&](uint32_t a, uint32_t b) { return ctr[3 * (size_t)a + ax] < ctr[3 * (size_t)b + ax] || (ctr[3 * (size_t)a + ax] == ctr[3 * (size_t)b + ax] && a < b); });
            nd.left = (int)A->nodes.size(); nd.right = nd.left + 1;
            A->nodes.emplace_back(); A->nodes.emplace_back();
            todo.push_back({rg.first, half, nd.left});
            todo.push_back({rg.first + half, rg.count - half, nd.right});
        }
        A->nodes[rg.node] = nd;
    }
    return A;
}

// Parameter at which the ray's line enters the box, or a negative value if it misses it (or leaves it behind the origin).
inline double box_entry(const Accel::Node& nd, const RayT& r)
{
    const double o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    double t0 = 0.0, t1 = 1e300;
    for (int k = 0; k < 3; k++) {
        if (d[k] == 0.0) { if (o[k] < nd.lo[k] || o[k] > nd.hi[k]) return -1.0; continue; }
        double a = (nd.lo[k] - o[k]) / d[k], b = (nd.hi[k] - o[k]) / d[k];
        if (a > b) std::swap(a, b);
        // one ulp of slack either way: the boxes are padded by 1e-2, the quotients are good to 1e-16 relative
        t0 = std::max(t0, a - 1e-9 * std::fabs(a)); t1 = std::min(t1, b + 1e-9 * std::fabs(b));
    }
    return t0 <= t1 ? t0 : -1.0;
}

Hit nearest_accel(const std::vector<Obj>& objs, const Accel& A, const RayT& r)
{
    Hit h{-1, INF};
    auto offer = [&](uint32_t i) {
        const double t = intersect(objs[i], r);
        if (t != 0.0 && (t < h.t || (t == h.t && (int)i < h.index))) { h.t = t; h.index = (int)i; }
    };
    for (uint32_t i : A.big) offer(i);
    if (A.nodes.empty()) return h;
    int stack[128], sp = 0;                       // median splits: depth <= log2(n / 4) + 2
    stack[sp++] = 0;
    while (sp) {
        const Accel::Node& nd = A.nodes[stack[--sp]];
        const double e = box_entry(nd, r);
        if (e < 0.0 || e > h.t) continue;
        if (nd.left < 0) { for (uint32_t j = nd.first; j < nd.first + nd.count; j++) offer(A.order[j]); continue; }
        if (sp + 2 > 128) { for (uint32_t j = nd.first; j < nd.first + nd.count; j++) offer(A.order[j]); continue; }   // cannot happen; stays exact
        stack[sp++] = nd.left; stack[sp++] = nd.right;
    }
    return h;
}

// Renderer.cu:227-243: linear scan, strict '<', so the lowest index wins ties; t == 0 means "no hit"
Hit nearest(const Ctx& c, const RayT& r)
{
    if (c.accel) return nearest_accel(c.objs_ref(), *c.accel, r);
    Hit h{-1, INF};
    const std::vector<Obj>& objs = c.objs_ref();
    for (size_t i = 0; i < objs.size(); i++) {
        const double t = intersect(objs[i], r);
        if (t != 0.0 && t < h.t) { h.t = t; h.index = (int)i; }
    }
    return h;
}

inline bool is_zero(const V& v) { return v.x == 0 && v.y == 0 && v.z == 0; }

// Renderer.cu:196-225.  `lane`, `thr` and `needed` are bookkeeping of the restatement (RNG node ids for the
// counter stream, cast accounting); they do not influence the value computed.
V deep_layers(Ctx& c, RayT ray, uint8_t depth, Rng& rng, uint32_t lane, V thr, bool needed)
{
    // new Vec3[maxDepth-2] twice (:198-199), zero-initialised by Vec3's default constructor, plus the clean OOB slot
    static thread_local V emissions[257], colors[257];
    for (uint32_t i = 0; i < c.maxDepth; i++) { emissions[i] = V(); colors[i] = V(); }
    for (; depth < c.maxDepth; depth++) {
        c.casts_reference++;
        if (needed && !is_zero(thr)) c.casts_needed++;
        const Hit h = nearest(c, ray);
        if (h.index == -1) break;
        const Obj& o = c.objs_ref()[h.index];
        const V P = ray.o + ray.d * h.t;
        rng.node(lane, depth);
        const Scatter s = scatter(o, P, ray.d, rng, depth);
        ray = s.ray;
        emissions[depth - 2] = o.emission;
        colors[depth - 2] = o.color;
        thr = mult(thr, o.color);
    }
    V pixel;
    for (int8_t i = (int8_t)(depth - 2); i >= 0; i--)               // :216 int8_t index: skipped when depth-2 > 127
        pixel = emissions[i] + mult(colors[i], pixel);
    return pixel;
}

// Renderer.cu:173-194.  `depth` BY REFERENCE: the second call made by first_layer enters with depth 2.
V second_layer(Ctx& c, const RayT& ray, uint8_t& depth, Rng& rng, uint32_t lane, V thr, bool probeOnly)
{
    c.casts_reference++;
    if (!is_zero(thr)) c.casts_needed++;
    const Hit h = nearest(c, ray);
    if (h.index == -1) return V();
    const Obj& o = c.objs_ref()[h.index];
    const V P = ray.o + ray.d * h.t;
    rng.node(lane, depth);
    const Scatter s = scatter(o, P, ray.d, rng, depth);
    depth++;
    V back;
    if (depth < c.maxDepth) {
        const V t2 = mult(thr, o.color);
        back = deep_layers(c, s.ray, depth, rng, lane, t2 * s.power, !probeOnly) * s.power;
        if (s.useSecond) back = back + deep_layers(c, s.second, depth, rng, 1, t2 * s.secondPower, !probeOnly) * s.secondPower;
    }
    return o.emission + mult(o.color, back);
}

// Renderer.cu:149-171
V first_layer(Ctx& c, const RayT& ray, Rng& rng)
{
    uint8_t depth = 0;
    c.casts_reference++;
    c.casts_needed++;
    const Hit h = nearest(c, ray);
    if (h.index == -1) return V();
    const Obj& o = c.objs_ref()[h.index];
    const V P = ray.o + ray.d * h.t;
    rng.node(0, depth);
    const Scatter s = scatter(o, P, ray.d, rng, depth);
    depth++;
    V back;
    if (depth < c.maxDepth) {
        back = second_layer(c, s.ray, depth, rng, 0, o.color * s.power, false) * s.power;
        // the second branch enters with depth == 2: its deep_layers(.,3) always folds to 0 (slot 0 never written),
        // so it is an emission probe of the first object hit
        if (s.useSecond) back = back + second_layer(c, s.second, depth, rng, 2, o.color * s.secondPower, true) * s.secondPower;
    }
    return o.emission + mult(o.color, back);
}

// Renderer.cu:112-147
V sample_pixel(Ctx& c, uint32_t px, uint32_t pz, Rng& rng)
{
    const double corrX = (c.W % 2 == 0) ? 0.5 : 0.0;
    const double corrZ = (c.W % 2 == 0) ? 0.5 : 0.0;               // :119 tests width_ for the z axis too
    const double stepX = (px < c.W / 2) ? (c.W / 2 - px) - corrX
                                         : ((double)c.W / 2 - px - 1.0) + ((corrX == 0.0) ? 1.0 : corrX);
    const double stepZ = (pz < c.H / 2) ? (c.H / 2 - pz) - corrZ
                                         : ((double)c.H / 2 - pz - 1.0) + ((corrZ == 0.0) ? 1.0 : corrZ);
    const V gaze = norm(c.camD + c.camX * stepX * (double)FOV_SCALE + c.vecZ * stepZ * (double)FOV_SCALE);   // :127
    V pixel;
    rng.pixel = pz * c.W + px;
    for (uint32_t i = 0; i < c.samples; i++) {
        rng.sample = i;
        double jx, jz;
        rng.jitter(jx, jz);
        const V tent = c.camX * jx + c.vecZ * jz;                                                          // :135
        const V origin = c.camO + c.camX * stepX + c.vecZ * stepZ + tent;                                  // :138
        pixel = pixel + first_layer(c, RayT{origin + c.camD * VIEWPORT_DISTANCE, gaze}, rng);              // :139
    }
    pixel.x = pixel.x / c.samples; pixel.y = pixel.y / c.samples; pixel.z = pixel.z / c.samples;          // :142-144
    return pixel;
}

// Renderer.cu:33-53: the pixel rectangle of reference thread (idX = threadIdx.x, idZ = blockIdx.x)
void cell_rect(uint32_t idX, uint32_t idZ, uint32_t W, uint32_t H, uint32_t& x0, uint32_t& z0, uint32_t& nx, uint32_t& nz)
{
    if (W <= BLOCK && H <= BLOCK) { x0 = idX; z0 = idZ; nx = 0; nz = 0; return; }   // :36-39 (renders nothing)
    const uint32_t xAdd = W % BLOCK, zAdd = H % BLOCK;
    uint32_t xStep = W / BLOCK, zStep = H / BLOCK;
    x0 = idX * xStep + ((idX >= xAdd) ? xAdd : idX);
    z0 = idZ * zStep + ((idZ >= zAdd) ? zAdd : idZ);
    nx = xStep + ((xAdd <= 0) ? 0 : ((idX < xAdd) ? 1 : 0));
    nz = zStep + ((zAdd <= 0) ? 0 : ((idZ < zAdd) ? 1 : 0));
}

Ctx make_ctx(const or_scene* s, uint32_t samples, uint32_t maxDepth)
{
    Ctx c;
    c.W = s->width; c.H = s->height; c.samples = samples; c.maxDepth = maxDepth;
    c.camO = V(s->camera); c.camD = V(s->camera + 3); c.camX = V(s->camera + 6);
    c.vecZ = norm(cross(c.camD, c.camX));                          // RenderController.cu:39
    auto objs = std::make_shared<std::vector<Obj>>();
    objs->reserve(s->n_objects);
    for (uint32_t i = 0; i < s->n_objects; i++) objs->push_back(make_obj(s->objects[i]));
    c.objs_shared = objs;
    return c;
}

}  // namespace

extern "C" {

static int render_impl(const or_scene* scene, uint32_t samples, uint32_t max_depth, int rng_mode, uint64_t seed, int begin,
                       int end, int nthreads, double* out_rgb, or_counts* counts, bool accel)
{
    if (!scene || !out_rgb || max_depth < 1 || max_depth > 255 || samples < 1) return -1;
    Ctx base0 = make_ctx(scene, samples, max_depth);
    if (accel) base0.accel = build_accel(base0.objs_ref());
    const Ctx& base = base0;
    const uint32_t W = base.W, H = base.H;
    const uint32_t nT = W <= BLOCK ? W : BLOCK, nB = H <= BLOCK ? H : BLOCK;   // RenderController.cu:53-54
    const int nUnits = rng_mode == OR_RNG_REFERENCE ? (int)(nT * nB) : (int)H;
    if (begin < 0) begin = 0;
    if (end < 0 || end > nUnits) end = nUnits;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > end - begin) nthreads = std::max(1, end - begin);
    std::atomic<int> next{begin};
    std::atomic<uint64_t> castsRef{0}, castsNeeded{0}, nSamples{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
        pool.emplace_back([&] {
            Ctx c = base;
            uint64_t ns = 0;
            for (;;) {
                const int u = next.fetch_add(1);
                if (u >= end) break;
                Rng rng;
                rng.mode = rng_mode;
                uint32_t x0, z0, nx, nz;
                if (rng_mode == OR_RNG_REFERENCE) {
                    const uint32_t idZ = (uint32_t)u / nT, idX = (uint32_t)u % nT;
                    cell_rect(idX, idZ, W, H, x0, z0, nx, nz);
                    curand_init(123456ULL, idX + idZ * nT, 0ULL, &rng.st);      // Renderer.cu:95-97
                } else {
                    x0 = 0; nx = W; z0 = (uint32_t)u; nz = 1;
                    rng.key[0] = (uint32_t)seed; rng.key[1] = (uint32_t)(seed >> 32);
                }
                for (uint32_t z = z0; z < z0 + nz; z++)                          // Renderer.cu:99-108
                    for (uint32_t x = x0; x < x0 + nx; x++) {
                        const V p = sample_pixel(c, x, z, rng);
                        double* o = out_rgb + 3 * ((size_t)z * W + x);
                        o[0] = p.x; o[1] = p.y; o[2] = p.z;
                        ns += samples;
                    }
            }
            castsRef += c.casts_reference; castsNeeded += c.casts_needed; nSamples += ns;
        });
    for (auto& th : pool) th.join();
    if (counts) { counts->samples = nSamples; counts->casts_reference = castsRef; counts->casts_needed = castsNeeded; }
    return 0;
}

int or_render(const or_scene* scene, uint32_t samples, uint32_t max_depth, int rng_mode, uint64_t seed, int begin,
              int end, int nthreads, double* out_rgb, or_counts* counts)
{
    return render_impl(scene, samples, max_depth, rng_mode, seed, begin, end, nthreads, out_rgb, counts, false);
}

int or_render_accel(const or_scene* scene, uint32_t samples, uint32_t max_depth, int rng_mode, uint64_t seed, int begin,
                    int end, int nthreads, double* out_rgb, or_counts* counts)
{
    return render_impl(scene, samples, max_depth, rng_mode, seed, begin, end, nthreads, out_rgb, counts, true);
}

static void nearest_hit_impl(const or_scene* scene, const double* rays, uint32_t n_rays, int32_t* out_index, double* out_t, bool accel);
void or_nearest_hit(const or_scene* scene, const double* rays, uint32_t n_rays, int32_t* out_index, double* out_t)
{
    nearest_hit_impl(scene, rays, n_rays, out_index, out_t, false);
}
void or_nearest_hit_accel(const or_scene* scene, const double* rays, uint32_t n_rays, int32_t* out_index, double* out_t)
{
    nearest_hit_impl(scene, rays, n_rays, out_index, out_t, true);
}

static void nearest_hit_impl(const or_scene* scene, const double* rays, uint32_t n_rays, int32_t* out_index, double* out_t, bool accel)
{
    Ctx c = make_ctx(scene, 1, 3);
    if (accel) c.accel = build_accel(c.objs_ref());
    // rays are independent: spread over the host cores (the scan of a million-primitive scene takes ~10 ms per ray)
    const unsigned hw = std::thread::hardware_concurrency();
    const uint32_t nt = std::max(1u, std::min<uint32_t>(hw ? hw : 1u, n_rays / 16u + 1u));
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < nt; t++)
        pool.emplace_back([&, t] {
            for (uint32_t i = t; i < n_rays; i += nt) {
                const RayT r{V(rays + 6 * i), V(rays + 6 * i + 3)};
                const Hit h = nearest(c, r);
                out_index[i] = h.index;
                out_t[i] = h.t;
            }
        });
    for (auto& th : pool) th.join();
}

double or_sphere_intersect(double radius, const double* c, const double* o, const double* d)
{
    or_object s{};
    s.type = 0; s.radius = radius; std::memcpy(s.position, c, 24);
    return sphere_intersect(make_obj(s), RayT{V(o), V(d)});
}

double or_plane_intersect(const double* north, const double* east, const double* c, const double* o, const double* d)
{
    or_object s{};
    s.type = 1; std::memcpy(s.north, north, 24); std::memcpy(s.east, east, 24); std::memcpy(s.position, c, 24);
    return plane_intersect(make_obj(s), RayT{V(o), V(d)});
}

void or_scatter(int kind, const double* geom, int reflection, const double* P, const double* incoming, int depth,
                unsigned long long subsequence, double* out)
{
    or_object s{};
    s.type = kind; s.reflection = reflection;
    if (kind == 0) { s.radius = geom[0]; std::memcpy(s.position, geom + 1, 24); }
    else { std::memcpy(s.north, geom, 24); std::memcpy(s.east, geom + 3, 24); std::memcpy(s.position, geom + 6, 24); }
    Rng rng;
    rng.mode = OR_RNG_REFERENCE;
    curand_init(123456ULL, subsequence, 0ULL, &rng.st);
    curandState ref = rng.st;
    const Scatter sc = scatter(make_obj(s), V(P), V(incoming), rng, depth);
    const V* v[4] = {&sc.ray.o, &sc.ray.d, &sc.second.o, &sc.second.d};
    for (int i = 0; i < 4; i++) { out[3 * i] = v[i]->x; out[3 * i + 1] = v[i]->y; out[3 * i + 2] = v[i]->z; }
    out[12] = sc.power; out[13] = sc.secondPower; out[14] = sc.useSecond ? 1.0 : 0.0;
    int draws = 0;
    while (std::memcmp(&ref, &rng.st, sizeof(ref)) != 0 && draws < 64) { curand(&ref); draws++; }
    out[15] = draws;
}

void or_philox4x32(const uint32_t* counter, const uint32_t* key, int rounds, uint32_t* out) { philox(counter, key, out, rounds > 0 ? rounds : OR_PHILOX_ROUNDS); }
double or_sym24(uint32_t x) { return s24(x); }
double or_uniform23(uint32_t x) { return u23(x); }

// Image.cpp:19-22
int or_to_rgb(double x) { return std::clamp(int(x * 255), 0, 255); }

}  // extern "C"
