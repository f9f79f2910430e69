// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Route-1 checker: compiles the reference's UNMODIFIED hot path
// (/root/reference/src/renderer/Renderer.cu and everything it #includes:
// Plane.cu, Sphere.cu, AObject.hpp, Coordinates.cu, HitData.cu) for the HOST with
// g++ and exposes it through a small C interface.  No reference source is copied
// into this repository: the files are compiled where they lie (see oracle/Makefile,
// REF=/root/reference) and only the resulting library lands in oracle/_ref/.
//
// What the shim supplies so that device code builds for the CPU:
//   * <cuda_runtime.h>            -> __device__/__host__/__global__ are empty for g++
//   * QUALIFIERS static inline    -> <curand_kernel.h> becomes a bit-exact host XORWOW
//   * threadIdx/blockIdx/blockDim -> thread_local structs (one "CUDA thread" = one cell)
//   * atomicAdd                   -> __atomic_fetch_add
//   * printf                      -> swallowed (the reference prints progress per row)
//   * operator new / new[]        -> zero-filled, over-allocated by 64 B, so that the
//                                    reference's out-of-bounds read
//                                    objectEmissions[maxDepth-2] (Renderer.cu:216-219)
//                                    deterministically returns 0 ("clean semantics").
//     The replacement operators are kept local to this library by the linker
//     version script (oracle/ref_shim.map).
#include <cstdlib>
#include <new>
void* operator new[](std::size_t n) { void* p = std::calloc(1, n + 64); if (!p) throw std::bad_alloc(); return p; }
void* operator new(std::size_t n)   { void* p = std::calloc(1, n + 64); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#define QUALIFIERS static inline
#include <curand_kernel.h>

#include "scene/SceneData.hpp"

struct ShimDim3 { unsigned x, y, z; };
static thread_local ShimDim3 shimThreadIdx, shimBlockIdx, shimBlockDim;
#define threadIdx shimThreadIdx
#define blockIdx  shimBlockIdx
#define blockDim  shimBlockDim
#undef  __global__
#define __global__
template <class T, class U> static T atomicAddShim(T* p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }
#define atomicAdd atomicAddShim
#define printf(...) ((void)0)
#include "Renderer.cu"   // the reference file, unmodified (-I$(REF)/src/renderer)
#undef printf
#undef atomicAdd
#undef threadIdx
#undef blockIdx
#undef blockDim

namespace {
using tracer::containers::Vec3;
using tracer::containers::Ray;
using tracer::scene::SceneData;
using tracer::scene::objects::ObjectData;
using tracer::scene::objects::Camera;
using tracer::scene::objects::EReflectionType;
using tracer::scene::objects::RayData;
using tracer::scene::objects::Sphere;
using tracer::scene::objects::Plane;

struct Loaded {
    bool ok = false;
    uint32_t W = 0, H = 0;
    Camera cam;
    std::vector<ObjectData> objs;
};

Loaded load(const char* path, int W_override, int H_override, std::string* printed = nullptr)
{
    Loaded l;
    // SceneData prints "Loading Scene Data..." etc. on std::cout: silence it.  Swapping cout's buffer is not thread
    // safe and callers may render cells from several threads: one load at a time.
    static std::mutex load_mutex;
    std::lock_guard<std::mutex> lock(load_mutex);
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    try {
        SceneData sd{std::string(path)};
        l.ok = sd.initScene();
        if (l.ok) {
            l.W = sd.getWidth(); l.H = sd.getHeight();
            l.cam = sd.getCamera();
            l.objs = sd.getObjectsData();
        }
    } catch (...) { l.ok = false; }
    std::cout.rdbuf(old);
    if (printed) *printed = sink.str();
    if (W_override > 0) l.W = (uint32_t)W_override;
    if (H_override > 0) l.H = (uint32_t)H_override;
    return l;
}

void run_cell(const Loaded& l, uint32_t cell, uint32_t nT, uint32_t samples, uint8_t depth, Vec3* img, Vec3 vecZ)
{
    shimBlockIdx  = {cell / nT, 0, 0};
    shimThreadIdx = {cell % nT, 0, 0};
    shimBlockDim  = {nT, 1, 1};
    tracer::renderer::Renderer r(samples, l.W, l.H, depth, l.cam);
    r.setUp(const_cast<ObjectData*>(l.objs.data()), (uint32_t)l.objs.size());
    r.start(img, vecZ);
}
}  // namespace

extern "C" {

// Scene dimensions as the reference's own SceneData parses them. Returns 0 on success.
int ref_scene_dims(const char* path, int* W, int* H, int* nObjects)
{
    Loaded l = load(path, 0, 0);
    if (!l.ok) return -1;
    *W = (int)l.W; *H = (int)l.H; *nObjects = (int)l.objs.size();
    return 0;
}

// The reference's OWN in-memory scene, byte for byte: the storage of SceneData::getObjectsData() (std::vector<ObjectData>,
// ObjectData.hpp:15-31) and getCamera() (Camera.hpp:8-16) - what RenderContoller::start() uploads
// (RenderController.cu:47-50) and what ipt_render_objects() takes.  Returns the number of objects, or -1; copies at
// most max_objects records of 144 bytes into out_objects and 72 bytes into out_camera.
static_assert(sizeof(ObjectData) == 144, "ObjectData layout changed: include/ipt_abi.h (ipt_render_objects) assumes 144 bytes");
static_assert(sizeof(Camera) == 72, "Camera layout changed: include/ipt_abi.h assumes 9 doubles");
int ref_scene_objects(const char* path, void* out_objects, int max_objects, void* out_camera, int* W, int* H)
{
    Loaded l = load(path, 0, 0);
    if (!l.ok) return -1;
    const int n = (int)l.objs.size();
    if (out_objects) std::memcpy(out_objects, l.objs.data(), sizeof(ObjectData) * (size_t)(n < max_objects ? n : max_objects));
    if (out_camera) std::memcpy(out_camera, &l.cam, sizeof(Camera));
    if (W) *W = (int)l.W;
    if (H) *H = (int)l.H;
    return n;
}

// What SceneData::initScene (SceneData.cpp:61-96) prints for `path` and whether it accepts the file: the differential test of
// the host layer's scene loader compares its verdict and message with this (tests/test_host.py).
int ref_scene_load_text(const char* path, char* out_text, int out_len)
{
    std::string printed;
    Loaded l = load(path, 0, 0, &printed);
    if (out_text && out_len > 0) std::snprintf(out_text, (size_t)out_len, "%s", printed.c_str());
    return l.ok ? (int)l.objs.size() : -1;
}

// Number of "CUDA threads" (cells) the reference launches: min(H,22) blocks x min(W,22) threads
// (RenderController.cu:53-56).
int ref_num_cells(int W, int H)
{
    const int nT = W <= 22 ? W : 22, nB = H <= 22 ? H : 22;
    return nT * nB;
}

// Runs cudaMain's per-thread body (Renderer.cu:254-265) for cells [cell_begin, cell_end) on
// `nthreads` host threads and writes the W*H*3 float64 image (untouched pixels stay 0).
// W_override/H_override > 0 replace the scene's width/height (config 4: spheres.json at 3840x2160).
int ref_render_cells(const char* path, int samples, int depth, int W_override, int H_override,
                     int cell_begin, int cell_end, int nthreads, double* out_rgb)
{
    Loaded l = load(path, W_override, H_override);
    if (!l.ok) return -1;
    Vec3 vecZ = (l.cam.direction_ % l.cam.orientation_).norm();   // RenderController.cu:39
    const uint32_t nT = l.W <= 22 ? l.W : 22, nB = l.H <= 22 ? l.H : 22;
    const uint32_t nCells = nT * nB;
    if (cell_begin < 0) cell_begin = 0;
    if (cell_end < 0 || (uint32_t)cell_end > nCells) cell_end = (int)nCells;
    std::vector<Vec3> img((size_t)l.W * l.H);
    std::atomic<int> next{cell_begin};
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
        pool.emplace_back([&] {
            for (;;) {
                int c = next.fetch_add(1);
                if (c >= cell_end) break;
                run_cell(l, (uint32_t)c, nT, (uint32_t)samples, (uint8_t)depth, img.data(), vecZ);
            }
        });
    for (auto& th : pool) th.join();
    for (size_t i = 0; i < img.size(); i++) {
        out_rgb[3 * i + 0] = img[i].xx_;
        out_rgb[3 * i + 1] = img[i].yy_;
        out_rgb[3 * i + 2] = img[i].zz_;
    }
    return 0;
}

// The same for an explicit list of cells (bench.py's bounded sample: cells strided over the frame), one call,
// `nthreads` host threads.  out_rgb may be NULL (timing only).
int ref_render_cell_list(const char* path, int samples, int depth, int W_override, int H_override,
                         const int* cells, int n_cells, int nthreads, double* out_rgb)
{
    Loaded l = load(path, W_override, H_override);
    if (!l.ok) return -1;
    Vec3 vecZ = (l.cam.direction_ % l.cam.orientation_).norm();
    const uint32_t nT = l.W <= 22 ? l.W : 22, nB = l.H <= 22 ? l.H : 22;
    std::vector<Vec3> img((size_t)l.W * l.H);
    std::atomic<int> next{0};
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
        pool.emplace_back([&] {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n_cells) break;
                if (cells[i] >= 0 && (uint32_t)cells[i] < nT * nB) run_cell(l, (uint32_t)cells[i], nT, (uint32_t)samples, (uint8_t)depth, img.data(), vecZ);
            }
        });
    for (auto& th : pool) th.join();
    if (out_rgb)
        for (size_t i = 0; i < img.size(); i++) { out_rgb[3 * i] = img[i].xx_; out_rgb[3 * i + 1] = img[i].yy_; out_rgb[3 * i + 2] = img[i].zz_; }
    return 0;
}

// ---- function-level known-answer access to the reference classes -------------------------------

// Sphere::intersect (Sphere.cu:25-39)
double ref_sphere_intersect(double radius, const double* c, const double* o, const double* d)
{
    Sphere s(radius, Vec3(c[0], c[1], c[2]), Vec3(), Vec3(), EReflectionType::Diffuse);
    return s.intersect(Ray(Vec3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2])));
}

// Plane::intersect (Plane.cu:47-68 with checkIfInBounds :87-100)
double ref_plane_intersect(const double* north, const double* east, const double* c, const double* o, const double* d)
{
    Plane p(Vec3(north[0], north[1], north[2]), Vec3(east[0], east[1], east[2]), Vec3(c[0], c[1], c[2]), Vec3(), Vec3(),
            EReflectionType::Diffuse);
    return p.intersect(Ray(Vec3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2])));
}

// calculateReflections (Sphere.cu:41-56 / Plane.cu:70-84 and AObject.hpp:83-135) with a fresh XORWOW
// state curand_init(123456, subsequence, 0).  kind 0 = sphere (geom = radius, centre), 1 = plane
// (geom = north, east, centre).  out = {ray.o(3), ray.d(3), second.o(3), second.d(3), power, secondPower,
// useSecond, draws consumed (32-bit curand() calls)} = 16 doubles.
void ref_scatter(int kind, const double* geom, int reflection, const double* P, const double* incoming, int depth,
                 unsigned long long subsequence, double* out)
{
    curandState st;
    curand_init(123456ULL, subsequence, 0ULL, &st);
    curandState ref = st;
    RayData rd;
    const Vec3 p(P[0], P[1], P[2]), in(incoming[0], incoming[1], incoming[2]);
    if (kind == 0) {
        Sphere s(geom[0], Vec3(geom[1], geom[2], geom[3]), Vec3(), Vec3(), (EReflectionType)reflection);
        rd = s.calculateReflections(p, in, st, (uint8_t)depth);
    } else {
        Plane pl(Vec3(geom[0], geom[1], geom[2]), Vec3(geom[3], geom[4], geom[5]), Vec3(geom[6], geom[7], geom[8]),
                 Vec3(), Vec3(), (EReflectionType)reflection);
        rd = pl.calculateReflections(p, in, st, (uint8_t)depth);
    }
    const Vec3* v[4] = {&rd.ray_.origin_, &rd.ray_.direction_, &rd.secondRay_.origin_, &rd.secondRay_.direction_};
    for (int i = 0; i < 4; i++) { out[3 * i] = v[i]->xx_; out[3 * i + 1] = v[i]->yy_; out[3 * i + 2] = v[i]->zz_; }
    out[12] = rd.power_; out[13] = rd.secondPower_; out[14] = rd.useSecond_ ? 1.0 : 0.0;
    // count the 32-bit draws consumed by replaying the untouched copy until the states agree
    int draws = 0;
    while (std::memcmp(&ref, &st, sizeof(st)) != 0 && draws < 64) { curand(&ref); draws++; }
    out[15] = draws;
}

// First values of the stream the reference seeds per thread (Renderer.cu:95-97): two raw curand() draws
// and one curand_uniform_double() — the RNG known-answer test of SURVEY.md §4.2.
void ref_xorwow_kat(unsigned long long subsequence, unsigned int* raw2, double* u)
{
    curandState st;
    curand_init(123456ULL, subsequence, 0ULL, &st);
    raw2[0] = curand(&st);
    raw2[1] = curand(&st);
    *u = curand_uniform_double(&st);
}

// toRgb is in Image.cpp (needs Magick++, absent) — not compiled here.

}  // extern "C"
