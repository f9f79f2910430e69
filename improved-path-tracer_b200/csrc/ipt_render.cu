// ipt_render.cu — implementation of the C ABI in include/ipt_abi.h on top of the kernels in ipt_kernels.cuh.
// Host side of the boundary the reference crosses in RenderContoller::start() (RenderController.cu:36-70):
// allocate, upload, launch, copy back — here as a resident context per GPU, with no synchronisation between
// the launches of a render and no CPU fallback.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ipt_abi.h"
#include "ipt_kernels.cuh"

using namespace ipt;

// ------------------------------------------------------------------------------------------------ errors
// The text of the most recent failure: of this thread if it has one that is at least as recent as any other thread's
// (ipt_render runs one worker thread per GPU; their messages must reach the caller's thread), under a lock.
static std::mutex g_err_lock;
static uint64_t g_err_clock = 0;                 // orders the messages of all threads
static std::string g_err_shared;                 // most recent message of any thread
static uint64_t g_err_shared_at = 0;
static thread_local std::string g_err;           // most recent message of this thread
static thread_local uint64_t g_err_at = 0;
static void set_err(const std::string& s)
{
    std::lock_guard<std::mutex> lock(g_err_lock);
    g_err = s; g_err_at = ++g_err_clock;
    g_err_shared = s; g_err_shared_at = g_err_at;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                                \
            return e_ == cudaErrorMemoryAllocation ? IPT_ERR_OUT_OF_MEMORY : IPT_ERR_CUDA;              \
        }                                                                                               \
    } while (0)

extern "C" int ipt_abi_version(void) { return IPT_ABI_VERSION; }

// Progress of a render (Renderer.cu:105-107 prints "\rRendering %.2f%%" per finished pixel row from the device): the share of
// wavefront batches that have FINISHED on the device, reported from the host thread that renders rank 0's tiles.
static std::atomic<ipt_progress_fn> g_progress_fn{nullptr};
static std::atomic<void*> g_progress_user{nullptr};
extern "C" void ipt_set_progress(ipt_progress_fn fn, void* user) { g_progress_user = user; g_progress_fn = fn; }
extern "C" const char* ipt_last_error(void)
{
    static thread_local std::string copy;        // stays valid until this thread asks again
    std::lock_guard<std::mutex> lock(g_err_lock);
    copy = g_err_at >= g_err_shared_at ? g_err : g_err_shared;
    return copy.c_str();
}

extern "C" int ipt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* ipt_device_name(int device)
{
    static thread_local char name[256];
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return ""; }
    std::snprintf(name, sizeof(name), "%s", prop.name);
    return name;
}

// Static interleaved tile schedule: diagonal interleave, so every tile row and every tile column is spread over
// all ranks (the literal 4K config puts all the work in the middle third of the frame).
extern "C" uint32_t ipt_tile_owner(uint32_t tile_x, uint32_t tile_y, uint32_t /*tiles_x*/, uint32_t world)
{
    return world <= 1 ? 0u : (tile_x + tile_y) % world;
}

// ------------------------------------------------------------------------------------------------ context
struct ipt_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // scene
    bool have_scene = false;
    uint32_t W = 0, H = 0, n_slots = 0, n_spheres = 0, n_objects = 0, n_nodes = 0;
    double cam[9] = {0};
    double max_emission = 0, max_color = 0;
    double scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};   // bounding box of everything a ray can hit
    bool scene_box_valid = false;
    void *geom32 = nullptr, *geom64 = nullptr, *mat32 = nullptr, *mat64 = nullptr;
    // The fp64 copy of the scene is packed with the fp32 one but crosses PCIe only when an fp64 render or trace asks for
    // it (upload_fp64): it stays in the pinned staging buffer until then.
    bool fp64_pending = false;
    size_t fp64_geom_bytes = 0, fp64_mat_off = 0, fp64_mat_bytes = 0;
    uint32_t* slot_obj = nullptr;
    float4* nodes = nullptr;
    float4* bslot = nullptr;         // fp32 BVH leaf records (2 x float4 per slot), built when the scene has a BVH
    // uniform grid (ipt_scene::grid_*): {first reference, count} per cell, the references (slots), the big primitives' slots
    uint2* grid_cells = nullptr;
    uint32_t* grid_refs = nullptr;
    GridHeader grid_hd = {};
    WideNode* wide = nullptr;        // the 8-wide quantised tree derived from the 2-wide one (ipt_wide.h), fp32 traversal
    uint32_t n_wide = 0, wide_depth = 0;
    uint2* wide_spill = nullptr;     // stack entries beyond the shared-memory part, per resident lane of k_extend_cw
    size_t wide_spill_bytes = 0;
    uint4* fast_blob = nullptr;      // fp32 brute-force layout (FastScene), built when the scene has no BVH
    uint32_t fast_words = 0;
    FastHeader fast_hd = {};
    size_t scene_bytes[8] = {0, 0, 0, 0, 0, 0, 0};   // sizes of the scene allocations (reused when unchanged)
    // render state
    uint4* q[2] = {nullptr, nullptr};
    size_t q_bytes = 0;
    uint2* hits = nullptr;           // split pipeline: one {t, slot} per queued ray
    size_t hits_bytes = 0;
    uint32_t* counters = nullptr;
    unsigned long long* traced = nullptr;
    double* lights = nullptr;               // IPT_FLAG_NEXT_EVENT: 8 doubles per emissive sphere {c.xyz, r, E.rgb, object index}
    uint32_t n_lights = 0;
    uint32_t* fast_hint = nullptr;          // fast_schedule: bounces per pass the last batch settled on (cleared by set_scene)
    unsigned long long* frame = nullptr;
    size_t frame_pixels = 0;
    float* out32 = nullptr;
    double* out64 = nullptr;
    uint8_t* out8 = nullptr;         // toRgb bytes of the whole frame (ipt_ctx_download_rgb8)
    size_t out8_bytes = 0;
    float* gather32 = nullptr;   // where resolve writes (own frame unless a gather target is set)
    double* gather64 = nullptr;
    void* ipc_mapped = nullptr;
    uint32_t* tile_ids = nullptr;
    size_t tile_cap = 0;
    uint32_t* mt_list = nullptr;     // active micro-tiles of this rank (see k_active_microtiles)
    uint32_t* mt_packed = nullptr;
    uint8_t* mt_flags = nullptr;
    size_t mt_cap = 0;
    uint32_t n_mt = 0;
    uint64_t mt_key = 0, active_pixels = 0;
    std::vector<uint32_t> tile_host;
    uint32_t tile_key[5] = {0, 0, 0, 0, 0};
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    ipt_stats last = {};
};

// Host-side packing of large scenes (a million primitives is ~250 MB of staging): index ranges over a few threads.
template <class F>
static void parallel_ranges(size_t n, F f)
{
    unsigned nt = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = std::getenv("IPT_HOST_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    if (n < (size_t)1 << 16 || nt <= 1) { f((size_t)0, n); return; }
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const size_t b = std::min(n, t * chunk), e = std::min(n, b + chunk);
        if (b < e) th.emplace_back([=, &f] { f(b, e); });
    }
    for (auto& t : th) t.join();
}

static int ensure_pinned(ipt_ctx* c, size_t bytes)
{
    if (c->pinned_bytes >= bytes) return IPT_OK;
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr; c->pinned_bytes = 0;
    CK(cudaMallocHost(&c->pinned, bytes));
    c->pinned_bytes = bytes;
    return IPT_OK;
}

extern "C" ipt_ctx* ipt_ctx_create(int device)
{
    int n = ipt_device_count();
    if (n <= 0) { set_err("CUDA capable device not found! Cannot continue"); return nullptr; }
    if (device < 0 || device >= n) { set_err("ipt_ctx_create: no such device"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); cudaGetLastError(); return nullptr; }
    ipt_ctx* c = new ipt_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        cudaMalloc(&c->counters, N_COUNTERS * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&c->traced, 64) != cudaSuccess ||
        cudaMalloc(&c->fast_hint, 4) != cudaSuccess || cudaMemset(c->fast_hint, 0, 4) != cudaSuccess) {
        set_err(std::string("ipt_ctx_create: ") + cudaGetErrorString(cudaGetLastError()));
        delete c;
        return nullptr;
    }
    c->sm_count = prop.multiProcessorCount;
    return c;
}

static void free_scene(ipt_ctx* c)
{
    cudaFree(c->geom32); cudaFree(c->geom64); cudaFree(c->mat32); cudaFree(c->mat64); cudaFree(c->slot_obj); cudaFree(c->nodes); cudaFree(c->fast_blob); cudaFree(c->bslot);
    cudaFree(c->wide); cudaFree(c->grid_cells); cudaFree(c->grid_refs);
    c->grid_cells = nullptr; c->grid_refs = nullptr; c->grid_hd = GridHeader{};
    c->bslot = nullptr; c->wide = nullptr; c->n_wide = 0;
    c->geom32 = c->geom64 = c->mat32 = c->mat64 = nullptr; c->slot_obj = nullptr; c->nodes = nullptr; c->fast_blob = nullptr; c->fast_words = 0;
    std::memset(c->scene_bytes, 0, sizeof(c->scene_bytes));
    c->have_scene = false;
}

extern "C" void ipt_ctx_destroy(ipt_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_scene(c);
    cudaFree(c->wide_spill);
    cudaFree(c->q[0]); cudaFree(c->q[1]); cudaFree(c->hits); cudaFree(c->counters); cudaFree(c->traced); cudaFree(c->fast_hint); cudaFree(c->lights); cudaFree(c->frame);
    cudaFree(c->out32); cudaFree(c->out8);   // out64 lives in out32's allocation cudaFree(c->tile_ids); cudaFree(c->mt_list); cudaFree(c->mt_packed); cudaFree(c->mt_flags);
    if (c->ipc_mapped) cudaIpcCloseMemHandle(c->ipc_mapped);
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

// the resolved fp64 frame follows the fp32 one in the same allocation, at this offset
static size_t frame64_offset(size_t px) { return (px * 12 + 255) & ~(size_t)255; }

// fp32 brute-force layout (see FastScene in ipt_device.cuh): spheres, axis-aligned rectangles per normal axis,
// general rectangles, materials.  Built in fp64 from the flattened scene, rounded once to fp32.
static std::vector<uint32_t> build_fast_blob(const ipt_scene* s, uint32_t group[3])
{
    group[0] = group[1] = group[2] = 0;
    struct AxRect { float pk, cI, cJ, hI, hJ; uint32_t obj; };
    std::vector<AxRect> ax[3];
    std::vector<uint32_t> gen;
    auto axis_of = [](const double* v, int& k, double& sign) {
        k = -1;
        for (int i = 0; i < 3; i++) {
            if (std::fabs(std::fabs(v[i]) - 1.0) <= 1e-12) { if (k >= 0) return false; k = i; sign = v[i]; }
            else if (std::fabs(v[i]) > 1e-12) return false;
        }
        return k >= 0;
    };
    for (uint32_t j = 0; j < s->n_rects; j++) {
        const double *pl = s->rect_plane + 4 * (size_t)j, *u = s->rect_u + 4 * (size_t)j, *v = s->rect_v + 4 * (size_t)j, *b = s->rect_bounds + 4 * (size_t)j;
        int K, iu, iv; double sk, su, sv;
        const bool aa = axis_of(pl, K, sk) && axis_of(u, iu, su) && axis_of(v, iv, sv) && b[0] == 0.0 && b[2] == 0.0 && iu != iv && iu != K && iv != K;
        if (!aa) { gen.push_back(j); continue; }
        const int I = K == 0 ? 1 : 0;
        const double cu = u[3] / su, cv = v[3] / sv;   // centre coordinates along the two in-plane axes
        AxRect r;
        r.pk = (float)(pl[3] / sk);
        if (iu == I) { r.cI = (float)cu; r.hI = (float)b[1]; r.cJ = (float)cv; r.hJ = (float)b[3]; }
        else { r.cI = (float)cv; r.hI = (float)b[3]; r.cJ = (float)cu; r.hJ = (float)b[1]; }
        r.obj = s->rect_object[j] | RECT_BIT;
        ax[K].push_back(r);
    }
    // box room (two rectangles per axis, at different plane coordinates): the lower wall of a pair goes first and the kernel
    // tests only the wall a ray travels towards when its origin lies between the two (fast_axis_pair); IPT_NO_PAIR=1: A/B runs
    bool box_pairs = gen.empty() && s->n_spheres <= 4 && !std::getenv("IPT_NO_PAIR");
    for (int k = 0; k < 3 && box_pairs; k++) {
        if (ax[k].size() != 2) { box_pairs = false; break; }
        if (ax[k][1].pk < ax[k][0].pk) std::swap(ax[k][0], ax[k][1]);
        box_pairs = ax[k][0].pk < ax[k][1].pk;
    }
    // Coplanar groups (maze.json: 23 of its 25 z rectangles lie on z = 200): the largest set of rectangles of a list that share
    // their plane goes to the front of the list, in its JSON order; the kernel computes t and the hit point of that plane once
    // and only the edge tests per rectangle (fast_axis_group).  Rectangles on different planes of one axis never tie in t, so
    // moving the group in front of the others changes no answer (Renderer.cu:235's tie rule concerns equal t).  IPT_NO_GROUP=1: A/B.
    for (int k = 0; k < 3 && !box_pairs && !std::getenv("IPT_NO_GROUP"); k++) {
        size_t best_n = 0; float best_p = 0.f;
        for (const AxRect& r : ax[k]) {
            size_t n_same = 0;
            for (const AxRect& q : ax[k]) n_same += q.pk == r.pk;
            if (n_same > best_n) { best_n = n_same; best_p = r.pk; }
        }
        if (best_n < 4) continue;
        std::stable_partition(ax[k].begin(), ax[k].end(), [&](const AxRect& r) { return r.pk == best_p; });
        group[k] = (uint32_t)best_n;
    }
    // the part of an axis list that the two-records-per-trip loop walks is padded to an even length with a record that can never be hit
    for (int k = 0; k < 3; k++)
        if ((ax[k].size() - group[k]) & 1) ax[k].push_back(AxRect{3.0e38f, 0.f, 0.f, -1.f, -1.f, NO_OBJECT});
    const uint32_t ns_real = s->n_spheres, ns = ns_real, ng = (uint32_t)gen.size(), no = s->n_objects;
    const uint32_t words = fast_blob_words(ns, (uint32_t)ax[0].size(), (uint32_t)ax[1].size(), (uint32_t)ax[2].size(), ng, no);
    std::vector<uint32_t> blob((size_t)words * 4, 0u);
    auto F = [](double x) { float f = (float)x; uint32_t u; std::memcpy(&u, &f, 4); return u; };
    auto Ff = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; };
    uint32_t* p = blob.data();
    p[0] = ns; p[1] = (uint32_t)ax[0].size(); p[2] = (uint32_t)ax[1].size(); p[3] = (uint32_t)ax[2].size(); p[4] = ng; p[5] = no;
    p[6] = box_pairs ? 1u : 0u;
    {   // p[7]: does any object carry a reflection value outside 0..2 (unknown material)?
        bool unk = false;
        for (uint32_t k = 0; k < no; k++) unk = unk || s->mat_reflection[k] < 0 || s->mat_reflection[k] > 2;
        p[7] = unk ? 1u : 0u;
    }
    p += 8;
    for (uint32_t i = 0; i < ns; i++)
        for (int k = 0; k < 4; k++) *p++ = i < ns_real ? F(s->sphere_cxyzr[4 * (size_t)i + k]) : F(k < 3 ? (double)NAN : 0.0);   // pad: NaN centre -> delta is NaN -> never a hit
    for (uint32_t i = 0; i < (ns + 3) / 4 * 4; i++) *p++ = i < ns_real ? s->sphere_object[i] : (NO_OBJECT - 1u);
    for (int k = 0; k < 3; k++)
        for (const AxRect& r : ax[k]) { *p++ = Ff(r.pk); *p++ = Ff(r.cI); *p++ = Ff(r.cJ); *p++ = Ff(r.hI); *p++ = Ff(r.hJ); *p++ = r.obj; *p++ = 0; *p++ = 0; }
    for (uint32_t j : gen) {
        for (int k = 0; k < 4; k++) *p++ = F(s->rect_plane[4 * (size_t)j + k]);
        for (int k = 0; k < 4; k++) *p++ = F(s->rect_u[4 * (size_t)j + k]);
        for (int k = 0; k < 4; k++) *p++ = F(s->rect_v[4 * (size_t)j + k]);
        for (int k = 0; k < 4; k++) *p++ = F(s->rect_bounds[4 * (size_t)j + k]);
    }
    for (uint32_t i = 0; i < (ng + 3) / 4 * 4; i++) *p++ = i < ng ? (s->rect_object[gen[i]] | RECT_BIT) : NO_OBJECT;
    for (uint32_t k = 0; k < no; k++) {
        bool anyE = false;
        for (int j = 0; j < 3; j++) anyE = anyE || s->mat_emission[3 * (size_t)k + j] != 0.0;
        for (int j = 0; j < 3; j++) *p++ = F(s->mat_color[3 * (size_t)k + j]);
        *p++ = F((double)s->mat_reflection[k]);
        for (int j = 0; j < 3; j++) *p++ = F(s->mat_emission[3 * (size_t)k + j]);
        *p++ = F(anyE ? 1.0 : 0.0);
    }
    return blob;
}

// Host -> device copy of the scene.  Builds the slot arrays (BVH leaf order, or spheres then rectangles), the
// fp32 and fp64 copies, and the packed BVH nodes in one pinned staging buffer, then issues the copies.
extern "C" int ipt_ctx_set_scene(ipt_ctx* c, const ipt_scene* s)
{
    if (!c || !s) { set_err("ipt_ctx_set_scene: null argument"); return IPT_ERR_BAD_ARGUMENT; }
    if (s->n_objects == 0 || s->n_objects != s->n_spheres + s->n_rects || s->width == 0 || s->height == 0) {
        set_err("ipt_ctx_set_scene: empty or inconsistent scene");
        return IPT_ERR_BAD_ARGUMENT;
    }
    if (s->n_objects >= 0x7FFFFFFFu) { set_err("ipt_ctx_set_scene: too many objects"); return IPT_ERR_BAD_ARGUMENT; }
    const bool bvh = s->n_bvh_nodes > 0;
    if (bvh && (s->n_bvh_slots != s->n_objects || !s->bvh_nodes || !s->bvh_slot_prim)) {
        set_err("ipt_ctx_set_scene: BVH must reference every primitive exactly once");
        return IPT_ERR_BAD_ARGUMENT;
    }
    CK(cudaSetDevice(c->device));
    cudaEvent_t e0 = c->ev0, e1 = c->ev1;
    CK(cudaEventRecord(e0, c->stream));
    const uint32_t n = s->n_objects, ns = s->n_spheres;
    // The 2-wide tree must be emitted parents first (children at higher indices: what host/bvh.cpp does; no cycles) and no
    // deeper than the traversal stacks of the generic kernels (ipt_device.cuh: int stack[64]).
    WideTree wt;
    if (bvh) {
        std::vector<uint8_t> depth(s->n_bvh_nodes, 0);
        uint32_t deepest = 1;
        depth[0] = 1;
        for (uint32_t i = 0; i < s->n_bvh_nodes; i++)
            for (int k = 0; k < 2; k++) {
                const int32_t ch = s->bvh_nodes[i].child[k];
                if (ch < 0) continue;
                if ((uint32_t)ch <= i) { set_err("ipt_ctx_set_scene: BVH nodes must be ordered parents first (child index > parent index)"); return IPT_ERR_BAD_ARGUMENT; }
                depth[ch] = (uint8_t)std::min(255, depth[i] + 1);
                deepest = std::max<uint32_t>(deepest, depth[ch]);
            }
        if (deepest > 60) { set_err("ipt_ctx_set_scene: BVH deeper than 60 levels (traversal stacks hold 64 entries)"); return IPT_ERR_BAD_ARGUMENT; }
        // IPT_BVH8=1 (A/B runs): the 8-wide quantised form of the tree (ipt_wide.h) for k_extend_cw.  Measured on the
        // 1M-primitive scene it loses to the 2-wide traversal, 1.2 against 1.6 Gbounces/s (profiles/README.md, round 2), so
        // the 2-wide tree is what the fp32 pipeline walks by default.
        if (std::getenv("IPT_BVH8")) {
            const uint32_t leaf_max = std::getenv("IPT_WIDE_LEAF") ? (uint32_t)std::atoi(std::getenv("IPT_WIDE_LEAF")) : 4u;
            if (const char* e = wide_collapse(s->bvh_nodes, s->n_bvh_nodes, n, leaf_max, wt)) {
                // a tree this form cannot express (leaves of more than 4 primitives) is walked 2-wide; a broken one is refused
                if (std::strncmp(e, "unsupported", 11) != 0) { set_err(std::string("ipt_ctx_set_scene: ") + e); return IPT_ERR_BAD_ARGUMENT; }
                wt = WideTree();
            }
        }
    }
    // One slot order for everything on the device: the 8-wide tree's when there is one (ipt_wide.h: whole leaves move, the
    // leaves of the 2-wide tree stay consecutive ranges), else the caller's.
    const uint32_t* slot_perm = wt.perm.empty() ? nullptr : wt.perm.data();
    const uint32_t* slot_inv = wt.inv.empty() ? nullptr : wt.inv.data();
    const size_t slot_pad = ((size_t)n + 3) / 4 * 4;
    const size_t b_geom64 = (size_t)n * 16 * 8, b_geom32 = (size_t)n * 16 * 4, b_mat64 = (size_t)n * 8 * 8, b_mat32 = (size_t)n * 8 * 4;
    const size_t b_slot = slot_pad * 4, b_nodes = (size_t)s->n_bvh_nodes * 64;
    const size_t total = b_geom64 + b_geom32 + b_mat64 + b_mat32 + b_slot + b_nodes;
    // uniform grid of the caller (ipt_scene::grid_*), checked here: indices in range, cell starts ascending
    const bool grid = bvh && s->grid_res[0] > 0 && !std::getenv("IPT_NO_GRID");
    uint64_t grid_n_cells = 0;
    if (grid) {
        grid_n_cells = (uint64_t)s->grid_res[0] * s->grid_res[1] * s->grid_res[2];
        const bool sane = s->grid_res[0] <= 1024 && s->grid_res[1] <= 1024 && s->grid_res[2] <= 1024 && s->grid_res[1] > 0 && s->grid_res[2] > 0 &&
                          grid_n_cells <= (1ull << 26) && s->grid_cell_start && s->n_grid_big <= GRID_MAX_BIG && (s->n_grid_big == 0 || s->grid_big) &&
                          (s->n_grid_refs == 0 || s->grid_refs) && s->grid_cell[0] > 0.f && s->grid_cell[1] > 0.f && s->grid_cell[2] > 0.f &&
                          s->n_grid_refs < (1u << 31);
        if (!sane) { set_err("ipt_ctx_set_scene: inconsistent grid"); return IPT_ERR_BAD_ARGUMENT; }
    }
    const size_t b_cells = grid ? (size_t)grid_n_cells * 8 : 0, b_refs = grid ? ((size_t)s->n_grid_refs + 4) * 4 : 0;
    const size_t b_bs = bvh ? (size_t)n * 32 : 0;      // typed 32-byte records of the fp32 traversal kernels
    int rc = ensure_pinned(c, total + b_cells + b_refs + b_bs);
    if (rc) return rc;
    // IPT_VERBOSE: where the time of an upload goes (host side)
    const bool verbose_up = std::getenv("IPT_VERBOSE") != nullptr;
    auto T_up = std::chrono::steady_clock::now();
    auto lap_up = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        if (verbose_up && n > 100000) std::fprintf(stderr, "[ipt] set_scene %-26s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - T_up).count());
        T_up = now;
    };
    char* pin = (char*)c->pinned;
    double* g64 = (double*)pin;
    float* g32 = (float*)(pin + b_geom64);
    double* m64 = (double*)(pin + b_geom64 + b_geom32);
    float* m32 = (float*)(pin + b_geom64 + b_geom32 + b_mat64);
    uint32_t* so = (uint32_t*)(pin + b_geom64 + b_geom32 + b_mat64 + b_mat32);
    float* nd = (float*)(pin + b_geom64 + b_geom32 + b_mat64 + b_mat32 + b_slot);
    std::memset(so + n, 0xFF, b_slot - (size_t)n * 4);
    std::atomic<const char*> bad{nullptr};              // first inconsistency a packing thread found
    uint2* gcells = (uint2*)(pin + total);
    uint32_t* grefs = (uint32_t*)(pin + total + b_cells);
    float* bs = (float*)(pin + total + b_cells + b_refs);
    if (grid) {
        parallel_ranges((size_t)grid_n_cells, [&](size_t lo_, size_t hi_) {
            for (size_t ci = lo_; ci < hi_; ci++) {
                const uint32_t a = s->grid_cell_start[ci], z = s->grid_cell_start[ci + 1];
                if (z < a || z > s->n_grid_refs) { bad = "ipt_ctx_set_scene: bad grid cell"; return; }
                gcells[ci] = make_uint2(a, z - a);
            }
        });
        parallel_ranges(s->n_grid_refs, [&](size_t lo_, size_t hi_) {
            for (size_t r = lo_; r < hi_; r++) {
                const uint32_t slot = s->grid_refs[r];
                if (slot >= n) { bad = "ipt_ctx_set_scene: bad grid reference"; return; }
                grefs[r] = slot_inv ? slot_inv[slot] : slot;
            }
        });
        for (uint32_t b = 0; b < s->n_grid_big; b++) if (s->grid_big[b] >= n) bad = "ipt_ctx_set_scene: bad grid reference";
    }
    parallel_ranges(n, [&](size_t lo_, size_t hi_) {
      for (size_t slot = lo_; slot < hi_; slot++) {
        uint32_t prim = bvh ? s->bvh_slot_prim[slot_perm ? slot_perm[slot] : slot] : (slot < ns ? (uint32_t)slot : (RECT_BIT | ((uint32_t)slot - ns)));
        const bool rect = (prim & RECT_BIT) != 0;
        const uint32_t idx = prim & ~RECT_BIT;
        if ((rect && idx >= s->n_rects) || (!rect && idx >= ns)) { bad = "ipt_ctx_set_scene: bad BVH slot"; return; }
        double* g = g64 + slot * 16;
        if (rect) {
            std::memcpy(g, s->rect_plane + 4 * (size_t)idx, 32);
            std::memcpy(g + 4, s->rect_u + 4 * (size_t)idx, 32);
            std::memcpy(g + 8, s->rect_v + 4 * (size_t)idx, 32);
            std::memcpy(g + 12, s->rect_bounds + 4 * (size_t)idx, 32);
            so[slot] = s->rect_object[idx] | RECT_BIT;
        } else {
            std::memcpy(g, s->sphere_cxyzr + 4 * (size_t)idx, 32);
            std::memset(g + 4, 0, 96);
            so[slot] = s->sphere_object[idx];
        }
        if ((so[slot] & ~RECT_BIT) >= n) { bad = "ipt_ctx_set_scene: object index out of range"; return; }
        for (int i = 0; i < 16; i++) g32[slot * 16 + i] = (float)g[i];
      }
    });
    if (bad.load()) { set_err(bad.load()); return IPT_ERR_BAD_ARGUMENT; }
    double maxE = 0, maxC = 0;
    std::mutex max_lock;
    parallel_ranges(n, [&](size_t lo_, size_t hi_) {
      double mE = 0, mC = 0;
      for (size_t k = lo_; k < hi_; k++) {
        double* m = m64 + k * 8;
        bool anyE = false;
        for (int j = 0; j < 3; j++) {
            m[j] = s->mat_color[3 * k + j];
            m[4 + j] = s->mat_emission[3 * k + j];
            anyE = anyE || m[4 + j] != 0.0;
            mE = std::max(mE, std::fabs(m[4 + j]));
            mC = std::max(mC, std::fabs(m[j]));
        }
        m[3] = (double)s->mat_reflection[k];
        m[7] = anyE ? 1.0 : 0.0;
        for (int i = 0; i < 8; i++) m32[k * 8 + i] = (float)m[i];
      }
      std::lock_guard<std::mutex> g(max_lock);
      maxE = std::max(maxE, mE); maxC = std::max(maxC, mC);
    });
    parallel_ranges(s->n_bvh_nodes, [&](size_t lo_, size_t hi_) {
      for (size_t i = lo_; i < hi_; i++) {
        const ipt_bvh_node& b = s->bvh_nodes[i];
        float* o = nd + (size_t)i * 16;
        o[0] = b.lo0[0]; o[1] = b.lo0[1]; o[2] = b.lo0[2]; o[3] = b.hi0[0];
        o[4] = b.hi0[1]; o[5] = b.hi0[2]; o[6] = b.lo1[0]; o[7] = b.lo1[1];
        o[8] = b.lo1[2]; o[9] = b.hi1[0]; o[10] = b.hi1[1]; o[11] = b.hi1[2];
        int32_t ch[2];
        for (int k = 0; k < 2; k++) {
            if (b.child[k] >= 0) {
                if ((uint32_t)b.child[k] >= s->n_bvh_nodes) { bad = "ipt_ctx_set_scene: bad BVH child"; return; }
                ch[k] = b.child[k];
            } else {
                const uint32_t first = (uint32_t)(~b.child[k]), cnt = b.count[k];
                if (cnt < 1 || cnt > 16 || (size_t)first + cnt > n || first >= (1u << 27)) { bad = "ipt_ctx_set_scene: bad BVH leaf"; return; }
                ch[k] = ~(int32_t)(((slot_inv ? slot_inv[first] : first) << 4) | (cnt - 1));
            }
        }
        std::memcpy(o + 12, ch, 8);
        o[14] = 0; o[15] = 0;
      }
    });
    if (bad.load()) { set_err(bad.load()); return IPT_ERR_BAD_ARGUMENT; }
    std::vector<uint32_t> blob;
    uint32_t fast_group[3] = {0, 0, 0};
    if (!bvh) { blob = build_fast_blob(s, fast_group); if (blob.size() * 4 > 200 * 1024) blob.clear(); }
    // fp32 BVH leaf records: typed 32-byte entries in leaf (slot) order
    lap_up("pack geometry + tree");
    if (bvh) {
        auto axis_of = [](const double* v, int& k, double& sign) {
            k = -1;
            for (int i = 0; i < 3; i++) {
                if (std::fabs(std::fabs(v[i]) - 1.0) <= 1e-12) { if (k >= 0) return false; k = i; sign = v[i]; }
                else if (std::fabs(v[i]) > 1e-12) return false;
            }
            return k >= 0;
        };
        parallel_ranges(n, [&](size_t lo_, size_t hi_) {
          for (size_t slot = lo_; slot < hi_; slot++) {
            const uint32_t prim = s->bvh_slot_prim[slot_perm ? slot_perm[slot] : slot], idx = prim & ~RECT_BIT;
            float* o8 = bs + slot * 8;
            for (int k = 0; k < 8; k++) o8[k] = 0.f;
            uint32_t kind, obj;
            if (!(prim & RECT_BIT)) {
                for (int k = 0; k < 4; k++) o8[k] = (float)s->sphere_cxyzr[4 * (size_t)idx + k];
                kind = 0; obj = s->sphere_object[idx];
            } else {
                const double *pl = s->rect_plane + 4 * (size_t)idx, *u = s->rect_u + 4 * (size_t)idx, *v = s->rect_v + 4 * (size_t)idx, *b = s->rect_bounds + 4 * (size_t)idx;
                int K, iu, iv; double sk, su, sv;
                obj = s->rect_object[idx] | RECT_BIT;
                if (axis_of(pl, K, sk) && axis_of(u, iu, su) && axis_of(v, iv, sv) && b[0] == 0.0 && b[2] == 0.0 && iu != iv && iu != K && iv != K) {
                    const int I = K == 0 ? 1 : 0;
                    const double cu = u[3] / su, cv = v[3] / sv;
                    o8[0] = (float)(pl[3] / sk);
                    if (iu == I) { o8[1] = (float)cu; o8[2] = (float)cv; o8[3] = (float)b[1]; o8[6] = (float)b[3]; }
                    else { o8[1] = (float)cv; o8[2] = (float)cu; o8[3] = (float)b[3]; o8[6] = (float)b[1]; }
                    kind = 1u + (uint32_t)K;
                } else kind = 4;
            }
            std::memcpy(o8 + 4, &kind, 4); std::memcpy(o8 + 5, &obj, 4);
          }
        });
    }
    lap_up("typed records");
    const size_t want[8] = {b_geom64, b_geom32, b_mat64, b_mat32, b_slot, b_nodes, blob.size() * 4 + b_bs + wt.nodes.size() * sizeof(WideNode), b_cells + b_refs};
    if (std::memcmp(want, c->scene_bytes, sizeof(want)) != 0 || !c->geom32) {
        free_scene(c);
        CK(cudaMalloc(&c->geom32, b_geom32));
        CK(cudaMalloc(&c->mat32, b_mat32));
        CK(cudaMalloc(&c->slot_obj, b_slot));
        if (b_nodes) CK(cudaMalloc(&c->nodes, b_nodes));
        if (!blob.empty()) CK(cudaMalloc(&c->fast_blob, blob.size() * 4));
        if (b_bs) CK(cudaMalloc(&c->bslot, b_bs));
        if (!wt.nodes.empty()) CK(cudaMalloc(&c->wide, wt.nodes.size() * sizeof(WideNode)));
        if (grid) { CK(cudaMalloc(&c->grid_cells, b_cells)); CK(cudaMalloc(&c->grid_refs, b_refs)); }
        std::memcpy(c->scene_bytes, want, sizeof(want));
    }
    c->have_scene = false;
    c->fp64_pending = true; c->fp64_geom_bytes = b_geom64; c->fp64_mat_off = b_geom64 + b_geom32; c->fp64_mat_bytes = b_mat64;
    CK(cudaMemcpyAsync(c->geom32, g32, b_geom32, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->mat32, m32, b_mat32, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->slot_obj, so, b_slot, cudaMemcpyHostToDevice, c->stream));
    if (b_nodes) CK(cudaMemcpyAsync(c->nodes, nd, b_nodes, cudaMemcpyHostToDevice, c->stream));
    if (b_bs) CK(cudaMemcpyAsync(c->bslot, bs, b_bs, cudaMemcpyHostToDevice, c->stream));
    if (!wt.nodes.empty()) CK(cudaMemcpyAsync(c->wide, wt.nodes.data(), wt.nodes.size() * sizeof(WideNode), cudaMemcpyHostToDevice, c->stream));
    c->n_wide = (uint32_t)wt.nodes.size(); c->wide_depth = wt.depth;
    c->grid_hd = GridHeader{};
    if (grid) {
        CK(cudaMemcpyAsync(c->grid_cells, gcells, b_cells, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->grid_refs, grefs, b_refs, cudaMemcpyHostToDevice, c->stream));
        GridHeader& g = c->grid_hd;
        for (int k = 0; k < 3; k++) {
            g.res[k] = s->grid_res[k]; g.lo[k] = s->grid_lo[k]; g.cs[k] = s->grid_cell[k]; g.inv_cs[k] = 1.f / s->grid_cell[k];
            g.hi[k] = s->grid_lo[k] + s->grid_cell[k] * (float)s->grid_res[k];
        }
        g.n_big = s->n_grid_big;
        for (uint32_t b = 0; b < s->n_grid_big; b++) g.big[b] = slot_inv ? slot_inv[s->grid_big[b]] : s->grid_big[b];
        if (std::getenv("IPT_VERBOSE"))
            std::fprintf(stderr, "[ipt] grid %u x %u x %u, %u references, %u big primitives: %.1f MB\n", g.res[0], g.res[1], g.res[2], s->n_grid_refs, g.n_big, (b_cells + b_refs) / 1e6);
    }
    if (std::getenv("IPT_VERBOSE") && bvh)
        std::fprintf(stderr, "[ipt] 8-wide tree: %zu nodes (%.2f children per node), depth %u, from %u 2-wide nodes\n", wt.nodes.size(),
                     wt.nodes.empty() ? 0.0 : wt.sum_children / wt.nodes.size(), wt.depth, s->n_bvh_nodes);
    c->fast_words = 0;
    if (!blob.empty()) {
        CK(cudaMemcpyAsync(c->fast_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice, c->stream));
        c->fast_words = (uint32_t)(blob.size() / 4);
        c->fast_hd = fast_header(blob[0], blob[1], blob[2], blob[3], blob[4], blob[5]);
        c->fast_hd.box_pairs = blob[6]; c->fast_hd.any_unknown = blob[7];
        for (int k = 0; k < 3; k++) c->fast_hd.group[k] = fast_group[k];
        if (blob[6]) {   // plane coordinates of the wall pairs: record r of the axis lists starts at word off_axs * 4 + r * 8
            const uint32_t* ax = blob.data() + (size_t)c->fast_hd.off_axs * 4;
            for (int k = 0; k < 3; k++) { std::memcpy(&c->fast_hd.box_lo[k], ax + (2 * k) * 8, 4); std::memcpy(&c->fast_hd.box_hi[k], ax + (2 * k + 1) * 8, 4); }
        }
        if (std::getenv("IPT_VERBOSE"))
            std::fprintf(stderr, "[ipt] typed lists: %u spheres, %u + %u + %u axis-aligned rectangles, %u general; box-room instantiation %d, wall pairs %s\n", blob[0], blob[1],
                         blob[2], blob[3], blob[4], fast_shape(blob[0], blob[1], blob[2], blob[3], blob[4]), blob[6] ? "yes" : "no");
    }
    CK(cudaEventRecord(e1, c->stream));
    lap_up("enqueue copies");
    CK(cudaStreamSynchronize(c->stream));
    lap_up("wait for copies");
    struct LapAtExit { decltype(lap_up)& f; ~LapAtExit() { f("scene box, lights, frame buffers"); } } lap_at_exit{lap_up};
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    c->last.upload_ms = ms;
    c->last.h2d_bytes = total - b_geom64 - b_mat64 + blob.size() * 4 + b_bs + wt.nodes.size() * sizeof(WideNode) + b_cells + b_refs;
    c->W = s->width; c->H = s->height; c->n_slots = n; c->n_spheres = bvh ? 0 : ns; c->n_objects = n; c->n_nodes = s->n_bvh_nodes;
    std::memcpy(c->cam, s->cam_origin, 24); std::memcpy(c->cam + 3, s->cam_dir, 24); std::memcpy(c->cam + 6, s->cam_orient, 24);
    c->max_emission = maxE; c->max_color = maxC;
    {   // bounding box of all primitives (for the exact camera-ray pruning): spheres c +- |r|; rectangles: the four
        // corners of the accepted parallelogram, solved from n.r = 0, u.r = +-u_hi, v.r = +-v_hi (degenerate ones,
        // whose normal is NaN, can never be hit and are skipped)
        // (one partial box per host thread, merged under a lock: a million primitives are 10 ms on one core)
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        bool ok = true;
        std::mutex box_lock;
        struct Part {
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            bool ok = true;
            void grow(const double* q) { for (int k = 0; k < 3; k++) { if (!(q[k] == q[k])) ok = false; lo[k] = std::min(lo[k], q[k]); hi[k] = std::max(hi[k], q[k]); } }
        };
        auto merge = [&](const Part& pt) {
            std::lock_guard<std::mutex> g(box_lock);
            ok = ok && pt.ok;
            for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], pt.lo[k]); hi[k] = std::max(hi[k], pt.hi[k]); }
        };
        parallel_ranges(s->n_spheres, [&](size_t lo_, size_t hi_) {
            Part pt;
            for (size_t i = lo_; i < hi_; i++) {
                const double* sp = s->sphere_cxyzr + 4 * i;
                const double r = std::fabs(sp[3]);
                const double a[3] = {sp[0] - r, sp[1] - r, sp[2] - r}, b[3] = {sp[0] + r, sp[1] + r, sp[2] + r};
                pt.grow(a); pt.grow(b);
            }
            merge(pt);
        });
        parallel_ranges(s->n_rects, [&](size_t lo_, size_t hi_) {
            Part pt;
            for (size_t j = lo_; j < hi_; j++) {
                const double *n3 = s->rect_plane + 4 * j, *u = s->rect_u + 4 * j, *v = s->rect_v + 4 * j, *b = s->rect_bounds + 4 * j;
                if (!(n3[0] == n3[0]) || !(u[0] == u[0]) || !(v[0] == v[0])) continue;   // NaN normal: never hit
                // solve [n; u; v] x = rhs by Cramer's rule
                const double det = n3[0] * (u[1] * v[2] - u[2] * v[1]) - n3[1] * (u[0] * v[2] - u[2] * v[0]) + n3[2] * (u[0] * v[1] - u[1] * v[0]);
                if (!(std::fabs(det) > 1e-12)) { pt.ok = false; continue; }
                for (int su = -1; su <= 1; su += 2)
                    for (int sv = -1; sv <= 1; sv += 2) {
                        const double r0 = n3[3], r1 = u[3] + su * b[1], r2 = v[3] + sv * b[3];
                        double q[3];
                        q[0] = (r0 * (u[1] * v[2] - u[2] * v[1]) - n3[1] * (r1 * v[2] - u[2] * r2) + n3[2] * (r1 * v[1] - u[1] * r2)) / det;
                        q[1] = (n3[0] * (r1 * v[2] - u[2] * r2) - r0 * (u[0] * v[2] - u[2] * v[0]) + n3[2] * (u[0] * r2 - r1 * v[0])) / det;
                        q[2] = (n3[0] * (u[1] * r2 - r1 * v[1]) - n3[1] * (u[0] * r2 - r1 * v[0]) + r0 * (u[0] * v[1] - u[1] * v[0])) / det;
                        pt.grow(q);
                    }
            }
            merge(pt);
        });
        c->scene_box_valid = ok && lo[0] <= hi[0];
        for (int k = 0; k < 3; k++) { c->scene_lo[k] = lo[k]; c->scene_hi[k] = hi[k]; }
        c->mt_key = 0;   // a new scene invalidates the active micro-tile list
    }
    {   // emissive spheres, for the next-event extension
        std::vector<double> L;
        for (uint32_t i = 0; i < s->n_spheres; i++) {
            const uint32_t obj = s->sphere_object[i];
            const double* E = s->mat_emission + 3 * (size_t)obj;
            if (E[0] == 0.0 && E[1] == 0.0 && E[2] == 0.0) continue;
            const double* sp = s->sphere_cxyzr + 4 * (size_t)i;
            const double rec[8] = {sp[0], sp[1], sp[2], std::fabs(sp[3]), E[0], E[1], E[2], (double)obj};
            L.insert(L.end(), rec, rec + 8);
        }
        cudaFree(c->lights); c->lights = nullptr;
        c->n_lights = (uint32_t)(L.size() / 8);
        if (c->n_lights) {
            CK(cudaMalloc(&c->lights, L.size() * sizeof(double)));
            CK(cudaMemcpy(c->lights, L.data(), L.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    c->have_scene = true;
    CK(cudaMemsetAsync(c->fast_hint, 0, 4, c->stream));   // a new scene: no survival measurement yet
    // frame buffers
    const size_t px = (size_t)c->W * c->H;
    if (px != c->frame_pixels) {
        if (c->gather32 == c->out32) { c->gather32 = nullptr; c->gather64 = nullptr; }
        cudaFree(c->frame); cudaFree(c->out32);
        c->frame = nullptr; c->out32 = nullptr; c->out64 = nullptr; c->frame_pixels = 0;
        // the resolved fp32 and fp64 frames share ONE allocation (fp64 at frame64_offset): one CUDA IPC handle then gives
        // another process both, so an fp64 render gathers across processes like an fp32 one
        CK(cudaMalloc(&c->frame, px * 24)); CK(cudaMalloc(&c->out32, frame64_offset(px) + px * 24));
        c->out64 = (double*)((char*)c->out32 + frame64_offset(px));
        CK(cudaMemsetAsync(c->out32, 0, frame64_offset(px) + px * 24, c->stream));
        c->frame_pixels = px;
        if (!c->gather32) { c->gather32 = c->out32; c->gather64 = c->out64; }
    }
    return IPT_OK;
}

template <typename R> static V3<R> hv(const double* p) { V3<R> v; v.x = (R)p[0]; v.y = (R)p[1]; v.z = (R)p[2]; return v; }

template <bool FIRST, int SHAPE, bool RR>
static int launch_bounce_fast(ipt_ctx* c, const KParams<float>& kp, int* grid_cache)
{
    auto kern = k_bounce_fast<FIRST, SHAPE, RR>;
    constexpr int threads = FastCfg<FIRST>::THREADS;
    const size_t smem = (size_t)kp.fast_words * 16 + (size_t)2 * 3 * threads * 16;   // scene lists + double-buffered ray staging
    if (*grid_cache == 0) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        if (per_sm < 1) { set_err("fast kernel does not fit on an SM"); return IPT_ERR_BAD_ARGUMENT; }
        *grid_cache = per_sm * c->sm_count;
    }
    kern<<<*grid_cache, threads, smem, c->stream>>>(kp);
    return IPT_OK;
}

// One batch of the split pipeline (fp32, BVH): raygen, then per bounce k_extend_bvh and k_bounce<MODE_SHADE>.
// Grids of the split pipeline's kernels (once per render) and, for k_extend_cw, the spill area of its stacks.
static int prepare_extend(ipt_ctx* c, KParams<float>& kp, size_t smem_top, int* grids)
{
    auto shade = k_bounce<float, MODE_SHADE, false, false>;
    if (kp.wide && grids[2] == 0) {
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extend_cw, BLOCK_THREADS, CW_SMEM));
        if (per_sm < 1) { set_err("k_extend_cw does not fit on an SM"); return IPT_ERR_BAD_ARGUMENT; }
        grids[2] = per_sm * c->sm_count;
        // stack entries beyond a lane's shared-memory part: a ray has at most one entry pending per level of the tree
        kp.wide_spill_cap = std::max(1u, c->wide_depth);
        const size_t need = (size_t)grids[2] * BLOCK_THREADS * kp.wide_spill_cap * sizeof(uint2);
        if (need > c->wide_spill_bytes) {
            cudaFree(c->wide_spill); c->wide_spill = nullptr; c->wide_spill_bytes = 0;
            CK(cudaMalloc(&c->wide_spill, need));
            c->wide_spill_bytes = need;
        }
        kp.wide_spill = c->wide_spill;
    }
    if (kp.grid_cells && grids[3] == 0) {
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extend_grid, BLOCK_THREADS, 0));
        if (per_sm < 1) { set_err("k_extend_grid does not fit on an SM"); return IPT_ERR_BAD_ARGUMENT; }
        grids[3] = per_sm * c->sm_count;
    }
    if (grids[0] == 0) {
        CK(cudaFuncSetAttribute(k_extend_bvh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_top));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_extend_bvh, BLOCK_THREADS, smem_top));
        if (per_sm < 1) { set_err("k_extend_bvh does not fit on an SM"); return IPT_ERR_BAD_ARGUMENT; }
        grids[0] = per_sm * c->sm_count;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shade, BLOCK_THREADS, 0));
        grids[1] = std::max(per_sm, 1) * c->sm_count;
    }
    return IPT_OK;
}

static int launch_split_batch(ipt_ctx* c, KParams<float>& kp, uint32_t max_depth, uint32_t cap, size_t smem_top, int* grids, uint64_t* launches)
{
    auto shade = k_bounce<float, MODE_SHADE, false, false>;
    int rc_prep = prepare_extend(c, kp, smem_top, grids);
    if (rc_prep) return rc_prep;
    kp.depth = 0;
    kp.qout = Queue{c->q[0], cap};
    k_raygen<float><<<c->sm_count * 8, BLOCK_THREADS, 0, c->stream>>>(kp);
    (*launches)++;
    for (uint32_t d = 0; d < max_depth; d++) {
        kp.depth = d;
        kp.qin = Queue{c->q[d & 1], cap};
        kp.qout = Queue{c->q[(d + 1) & 1], cap};
        if (kp.grid_cells) k_extend_grid<<<grids[3], BLOCK_THREADS, 0, c->stream>>>(kp);
        else if (kp.wide) k_extend_cw<<<grids[2], BLOCK_THREADS, CW_SMEM, c->stream>>>(kp);
        else k_extend_bvh<<<grids[0], BLOCK_THREADS, smem_top, c->stream>>>(kp);
        shade<<<grids[1], BLOCK_THREADS, 0, c->stream>>>(kp);
        *launches += 2;
    }
    return IPT_OK;
}

template <typename R, int MODE, bool FIRST, bool DEFER = false>
static int launch_bounce(ipt_ctx* c, const KParams<R>& kp, size_t smem, int* grid_cache)
{
    auto kern = k_bounce<R, MODE, FIRST, DEFER>;
    if (*grid_cache == 0) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK_THREADS, smem));
        if (per_sm < 1) { set_err("kernel does not fit on an SM (scene too large for shared memory: build a BVH)"); return IPT_ERR_BAD_ARGUMENT; }
        *grid_cache = per_sm * c->sm_count;
    }
    kern<<<*grid_cache, BLOCK_THREADS, smem, c->stream>>>(kp);
    return IPT_OK;
}

static int build_tiles(ipt_ctx* c, uint32_t tile_w, uint32_t tile_h, uint32_t rank, uint32_t world, uint32_t* tiles_x_out)
{
    const uint32_t tiles_x = (c->W + tile_w - 1) / tile_w, tiles_y = (c->H + tile_h - 1) / tile_h;
    *tiles_x_out = tiles_x;
    const uint32_t key[5] = {tile_w, tile_h, rank, world, tiles_x * 65536u + tiles_y};
    if (std::memcmp(key, c->tile_key, sizeof(key)) == 0 && c->tile_ids) return IPT_OK;
    c->tile_host.clear();
    for (uint32_t ty = 0; ty < tiles_y; ty++)
        for (uint32_t tx = 0; tx < tiles_x; tx++)
            if (ipt_tile_owner(tx, ty, tiles_x, world) == rank) c->tile_host.push_back(ty * tiles_x + tx);
    const size_t nb = std::max<size_t>(c->tile_host.size(), 1) * 4;
    if (nb > c->tile_cap) {
        cudaFree(c->tile_ids); c->tile_ids = nullptr; c->tile_cap = 0;
        CK(cudaMalloc(&c->tile_ids, nb));
        c->tile_cap = nb;
    }
    if (!c->tile_host.empty())
        CK(cudaMemcpyAsync(c->tile_ids, c->tile_host.data(), c->tile_host.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));   // tile_host may be rebuilt by the next call
    std::memcpy(c->tile_key, key, sizeof(key));
    return IPT_OK;
}

// The active micro-tile list of this rank (exact camera-ray pruning, see k_active_microtiles).
static int build_active_microtiles(ipt_ctx* c, uint32_t tile_w, uint32_t tile_h, uint32_t tiles_x, uint64_t key)
{
    if (c->mt_key == key && c->mt_list) return IPT_OK;
    const uint32_t mt_x = tile_w / 8, mt_per_tile = mt_x * (tile_h / 4);
    const size_t total = c->tile_host.size() * (size_t)mt_per_tile;
    if (total > c->mt_cap || !c->mt_list) {
        cudaFree(c->mt_list); cudaFree(c->mt_packed); cudaFree(c->mt_flags);
        c->mt_list = c->mt_packed = nullptr; c->mt_flags = nullptr; c->mt_cap = 0;
        const size_t cap = std::max<size_t>(total, 1);
        CK(cudaMalloc(&c->mt_list, cap * 4)); CK(cudaMalloc(&c->mt_packed, cap * 4)); CK(cudaMalloc(&c->mt_flags, cap));
        c->mt_cap = cap;
    }
    c->n_mt = 0;
    if (total == 0) { c->mt_key = key; return IPT_OK; }
    ActiveParams ap;
    std::memset(&ap, 0, sizeof(ap));
    const double* D = c->cam + 3; const double* X = c->cam + 6;
    double Z[3] = {D[1] * X[2] - D[2] * X[1], D[2] * X[0] - D[0] * X[2], D[0] * X[1] - D[1] * X[0]};
    const double zl = 1.0 / std::sqrt(Z[0] * Z[0] + Z[1] * Z[1] + Z[2] * Z[2]);
    const bool cull = c->scene_box_valid && !std::getenv("IPT_NO_CULL");
    for (int k = 0; k < 3; k++) {
        ap.camO[k] = c->cam[k]; ap.camD[k] = D[k]; ap.camX[k] = X[k]; ap.camZ[k] = Z[k] * zl;
        // grown by the +-1 pixel jitter along X and Z (<= sqrt 2 per axis) plus slack for fp32 ray generation
        const double ext = std::max(std::fabs(c->scene_lo[k]), std::fabs(c->scene_hi[k]));
        ap.lo[k] = cull ? c->scene_lo[k] - 2.0 - 1e-5 * ext : -1e300;
        ap.hi[k] = cull ? c->scene_hi[k] + 2.0 + 1e-5 * ext : 1e300;
    }
    ap.fov = (double)0.0009f;
    ap.W = c->W; ap.H = c->H; ap.tile_ids = c->tile_ids; ap.n_tiles_local = (uint32_t)c->tile_host.size(); ap.tiles_x = tiles_x;
    ap.tile_w = tile_w; ap.tile_h = tile_h; ap.mt_x = mt_x; ap.mt_per_tile = mt_per_tile;
    ap.packed = c->mt_packed; ap.flags = c->mt_flags;
    k_active_microtiles<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, c->stream>>>(ap);
    std::vector<uint32_t> packed(total);
    std::vector<uint8_t> flags(total);
    CK(cudaMemcpyAsync(packed.data(), c->mt_packed, total * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(flags.data(), c->mt_flags, total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> list;
    list.reserve(total);
    for (size_t i = 0; i < total; i++) if (flags[i]) list.push_back(packed[i]);   // order kept: tile-major
    if (!list.empty()) CK(cudaMemcpyAsync(c->mt_list, list.data(), list.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->n_mt = (uint32_t)list.size();
    c->active_pixels = 0;
    for (uint32_t xy : list) {
        const uint32_t x0 = (xy & 0xFFFFu) * 8, z0 = (xy >> 16) * 4;
        c->active_pixels += (uint64_t)(std::min(c->W, x0 + 8) - std::min(c->W, x0)) * (std::min(c->H, z0 + 4) - std::min(c->H, z0));
    }
    c->mt_key = key;
    return IPT_OK;
}

// fp64 copy of the scene: from the staging buffer of the last ipt_ctx_set_scene to the device, on first use.
static int upload_fp64(ipt_ctx* c)
{
    if (!c->fp64_pending) return IPT_OK;
    if (!c->geom64) { CK(cudaMalloc(&c->geom64, c->fp64_geom_bytes)); CK(cudaMalloc(&c->mat64, c->fp64_mat_bytes)); }
    const char* pin = (const char*)c->pinned;
    CK(cudaMemcpyAsync(c->geom64, pin, c->fp64_geom_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->mat64, pin + c->fp64_mat_off, c->fp64_mat_bytes, cudaMemcpyHostToDevice, c->stream));
    c->last.h2d_bytes += c->fp64_geom_bytes + c->fp64_mat_bytes;
    c->fp64_pending = false;
    return IPT_OK;
}

template <typename R>
static int render_typed(ipt_ctx* c, const ipt_params& prm, uint32_t tile_w, uint32_t tile_h, uint32_t tiles_x, ipt_stats* st)
{
    if (sizeof(R) == 8) { int rc64 = upload_fp64(c); if (rc64) return rc64; }
    const bool bvh = c->n_nodes > 0;
    KParams<R> kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.sc.geom = (const R4<R>*)(sizeof(R) == 4 ? c->geom32 : c->geom64);
    kp.sc.mat = (const R4<R>*)(sizeof(R) == 4 ? c->mat32 : c->mat64);
    kp.sc.slot_obj = c->slot_obj;
    kp.sc.n_slots = c->n_slots; kp.sc.n_spheres = c->n_spheres; kp.sc.n_objects = c->n_objects;
    kp.sc.nodes = c->nodes; kp.sc.n_nodes = c->n_nodes; kp.sc.bslot = std::getenv("IPT_GENERIC_KERNEL") ? nullptr : c->bslot;
    // the next-event extension lives in the generic fused kernels only
    const bool nee = (prm.flags & IPT_FLAG_NEXT_EVENT) != 0 && c->n_lights > 0;
    kp.sc.lights = nee ? c->lights : nullptr; kp.sc.n_lights = nee ? c->n_lights : 0;
    // camera: vecZ = normalize(direction x orientation) in fp64 on the host (RenderController.cu:39)
    const double* D = c->cam + 3; const double* X = c->cam + 6;
    double Z[3] = {D[1] * X[2] - D[2] * X[1], D[2] * X[0] - D[0] * X[2], D[0] * X[1] - D[1] * X[0]};
    const double zl = 1.0 / std::sqrt(Z[0] * Z[0] + Z[1] * Z[1] + Z[2] * Z[2]);
    for (double& z : Z) z *= zl;
    kp.camO = hv<R>(c->cam); kp.camD = hv<R>(D); kp.camX = hv<R>(X); kp.camZ = hv<R>(Z);
    kp.fov = (R)0.0009f;
    kp.W = c->W; kp.H = c->H; kp.spp = prm.samples; kp.maxDepth = prm.max_depth;
    kp.flags = prm.flags;
    kp.strat_n = (uint32_t)std::floor(std::sqrt((double)prm.samples));
    kp.keys = philox_expand((uint32_t)prm.seed, (uint32_t)(prm.seed >> 32));
    kp.tile_ids = c->tile_ids; kp.n_tiles_local = (uint32_t)c->tile_host.size(); kp.tiles_x = tiles_x;
    kp.tile_w = tile_w; kp.tile_h = tile_h; kp.mt_x = tile_w / 8; kp.mt_per_tile = (tile_w / 8) * (tile_h / 4);
    kp.mt_list = c->mt_list; kp.n_mt = c->n_mt;
    kp.counters = c->counters; kp.traced = c->traced; kp.frame = c->frame;

    // fixed-point scale: largest power of two such that spp * (bound on one sample's radiance) fits in 62 bits
    bool float_accum = (prm.flags & IPT_FLAG_FLOAT_ACCUM) != 0;
    double scale = 1.0;
    {
        const double mc = std::max(1.0, c->max_color);
        double bound = std::max(c->max_emission, 1e-30) * 2.0 * (prm.max_depth + 1.0) * std::pow(mc, (double)prm.max_depth);
        bound *= (double)prm.samples;
        // next-event estimation adds, per diffuse hit, up to lobe (0.433) x 1/q (2 pi) x n_lights x E of explicit light
        if (nee) bound *= std::max(1.0, 2.7 * (double)c->n_lights);
        const int bits = 62 - (int)std::ceil(std::log2(bound));
        // below 2^-30 resolution the quantisation would show in dim pixels: use fp64 atomics instead (not bit-reproducible)
        if (!std::isfinite(bound) || bits < 30) float_accum = true;
        else scale = std::ldexp(1.0, std::min(bits, 40));
    }
    if (float_accum) kp.flags |= IPT_FLAG_FLOAT_ACCUM; else kp.flags &= ~IPT_FLAG_FLOAT_ACCUM;
    kp.fixed_scale = scale;

    // batches
    const uint64_t total_mt = c->n_mt;   // only the micro-tiles whose camera rays can reach the scene
    const uint64_t total_groups = total_mt * prm.samples;
    // default batch: a quarter of the render, between 16 Mi and 256 Mi camera rays (2 x 26 GB of fp32 ray queues; HBM capacity is
    // not a constraint at 180 GB) - measured on B200 on the 4K config: 4 Mi 27.9, 16 Mi 34.4, 64 Mi 37.3 Gbounces/s in round 1
    // (fewer launches, shorter tails), 32 / 64 / 128 / 256 Mi 65.4 / 66.2 / 66.4 / 66.8 in round 2.  A one-shot render
    // (ipt_render: queues allocated for this call only, ~9 ms per GB) stops at 64 Mi.
    const uint64_t total_samples = total_groups * 32;
    uint64_t B = prm.batch_samples;
    if (!B) {
        const uint64_t cap_b = (prm.reserved[0] & 1u) ? (1ull << 26) : (1ull << 28);
        B = 1ull << 24;
        while (B < cap_b && B * 4 < total_samples) B <<= 1;
    }
    B = std::max<uint64_t>(32, std::min<uint64_t>(B, 1u << 28) / 32 * 32);
    B = std::min<uint64_t>(B, std::max<uint64_t>(32, total_groups * 32));
    // maxDepth >= 130: deep paths carry their deferred radiance in extra queue planes (see k_bounce, DEFER)
    const bool defer = prm.max_depth >= 130;
    const size_t ray_bytes = (size_t)16 * (QPlanes<R>::N + (defer ? QPlanes<R>::ACC : 0));
    // a sample has at most two live rays; + room for the dead tails of the fast kernel's per-warp output blocks
    // (at most OUT_BLOCK - 1 dead slots per resident warp of k_bounce_fast, whatever its launch configuration: 64 warps per SM)
    const uint32_t tail_slack = (uint32_t)c->sm_count * 64u * OUT_BLOCK;
    auto cap_of = [tail_slack](uint64_t b) { return (uint32_t)(2 * b) + tail_slack; };
    if ((size_t)cap_of(B) * ray_bytes > c->q_bytes) {   // (queues of that size already allocated: nothing to decide, and no driver call)
        // the default batch shrinks on devices (or in processes) where two queues of that size do not fit comfortably
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = ~(size_t)0; }
        const size_t avail = free_b / 2 + 2 * c->q_bytes;
        while (!prm.batch_samples && B > (1u << 20) && 2 * (size_t)cap_of(B) * ray_bytes > avail) B = B / 2 / 32 * 32;
    }
    const uint32_t cap = cap_of(B);
    const size_t q_bytes = (size_t)cap * ray_bytes;
    if (q_bytes > c->q_bytes) {
        cudaFree(c->q[0]); cudaFree(c->q[1]); c->q[0] = c->q[1] = nullptr; c->q_bytes = 0;
        CK(cudaMalloc(&c->q[0], q_bytes)); CK(cudaMalloc(&c->q[1], q_bytes));
        c->q_bytes = q_bytes;
    }
    const size_t smem = bvh ? (size_t)std::min<uint32_t>(c->n_nodes, BVH_TOP_NODES) * 64
                            : ((size_t)c->n_slots * 4 + (size_t)c->n_objects * 2) * sizeof(R4<R>) + ((size_t)c->n_slots + 3) / 4 * 16;
    if (smem > 227 * 1024) { set_err("scene too large for the shared-memory path: pass a BVH"); return IPT_ERR_BAD_ARGUMENT; }

    // fp32 + BVH: split pipeline (raygen -> [extend with lane refill -> shade + compact] per bounce)
    const bool use_split = sizeof(R) == 4 && bvh && kp.sc.bslot != nullptr && !defer && !nee && !std::getenv("IPT_FUSED_BVH");
    if (use_split && (size_t)cap * 8 > c->hits_bytes) {
        cudaFree(c->hits); c->hits = nullptr; c->hits_bytes = 0;
        CK(cudaMalloc(&c->hits, (size_t)cap * 8));
        c->hits_bytes = (size_t)cap * 8;
    }
    kp.hits = c->hits;
    // 8-wide traversal: only with IPT_BVH8=1 at ipt_ctx_set_scene time (and leaves of at most 4 primitives)
    kp.wide = use_split ? c->wide : nullptr;
    // uniform grid: when the host layer built one for the scene (host/grid.cpp) it replaces the tree in the fp32 pipeline
    kp.grid_cells = use_split ? c->grid_cells : nullptr; kp.grid_refs = c->grid_refs; kp.grid = c->grid_hd;
    kp.descend_min = std::getenv("IPT_DESCEND_MIN") ? (uint32_t)std::atoi(std::getenv("IPT_DESCEND_MIN")) : 12u;
    // (measured on the 1M-primitive scene, profiles/README.md round 2: the grid walk likes 12 / 16, the tree walks 8 / 8)
    const bool grid_walk = use_split && c->grid_cells != nullptr;
    kp.refill_min = std::getenv("IPT_REFILL_MIN") ? (uint32_t)std::atoi(std::getenv("IPT_REFILL_MIN")) : (grid_walk ? 12u : 8u);
    kp.leaf_min = std::getenv("IPT_LEAF_MIN") ? (uint32_t)std::atoi(std::getenv("IPT_LEAF_MIN")) : (grid_walk ? 16u : 8u);
    kp.static_slices = std::getenv("IPT_STATIC_SLICES") ? 1u : 0u;
    // fp32 + no BVH: the typed-list kernel (k_bounce_fast); IPT_GENERIC_KERNEL=1 forces the generic one (A/B runs)
    // (a scene with an unknown reflection value - none shipped has one - goes through the generic kernel, which carries that case)
    const bool use_fast = sizeof(R) == 4 && !bvh && !defer && !nee && c->fast_blob && c->fast_words > 0 && !c->fast_hd.any_unknown && !std::getenv("IPT_GENERIC_KERNEL");
    kp.fast_blob = c->fast_blob; kp.fast_words = c->fast_words; kp.fast_hd = c->fast_hd;
    const bool sync_passes = std::getenv("IPT_SYNC_PASSES") != nullptr;   // diagnostic: drain the GPU between passes
    // A/B knob: fixed number of bounces per pass for the fast kernel's passes from depth 2 on (default 0 = adaptive)
    const uint32_t fast_k = std::getenv("IPT_FAST_K") ? (uint32_t)std::min(64, std::max(0, std::atoi(std::getenv("IPT_FAST_K")))) : 0u;
    // fast kernel: pass 0 (camera rays, at least the bounces at depth 0 and 1), then launches that advance several
    // bounces each (fast_schedule); with a fixed IPT_FAST_K the host knows how many are needed
    uint32_t passes_per_batch = prm.max_depth;
    if (use_fast) {
        const uint32_t first = std::min(2u, prm.max_depth);
        passes_per_batch = 1 + (fast_k ? (prm.max_depth - first + fast_k - 1) / fast_k
                                       : std::min(prm.max_depth - first, FAST_LATER_LAUNCHES));
    }
    kp.fast_hint = c->fast_hint;
    int shape = std::getenv("IPT_NO_SHAPE") ? 0 : fast_shape(c->fast_hd.n_sph, c->fast_hd.n_x, c->fast_hd.n_y, c->fast_hd.n_z, c->fast_hd.n_gen);   // A/B knob
    if (shape == 0 && (c->fast_hd.group[0] | c->fast_hd.group[1] | c->fast_hd.group[2])) shape = -1;   // list scene with a coplanar group (build_fast_blob)
    CK(cudaMemsetAsync(c->frame, 0, c->frame_pixels * 24, c->stream));
    CK(cudaMemsetAsync(c->traced, 0, 64, c->stream));   // [0] casts, [1] queue records moved, [2..6] traversal work (add_work)
    CK(cudaEventRecord(c->ev0, c->stream));
    int grid_first = 0, grid_next = 0, split_grids[4] = {0, 0, 0, 0};
    // IPT_PASS_TIMES=1 (diagnostic): an event before every pass of the fused pipelines, per-depth sums on stderr
    const bool pass_times = std::getenv("IPT_PASS_TIMES") != nullptr && !use_split;
    struct PassDiag {                 // released on every return path
        std::vector<cudaEvent_t> events;
        float* mhz = nullptr;         // SM MHz sampled after each of the first 4096 passes (IPT_PASS_CLOCKS)
        ~PassDiag() { for (cudaEvent_t e : events) cudaEventDestroy(e); if (mhz) cudaFree(mhz); }
    } diag;
    std::vector<cudaEvent_t>& pass_events = diag.events;
    float*& clock_probe = diag.mhz;
    if (pass_times && std::getenv("IPT_PASS_CLOCKS")) CK(cudaMallocManaged(&clock_probe, 4096 * sizeof(float)));
    uint64_t launches = 0, batches = 0;
    // a batch = mt_per_batch micro-tiles x smp_per_batch samples (sample ids walk the micro-tiles first, decode_sample)
    const uint64_t groups_per_batch = B / 32;
    const uint32_t smp_per_batch = (uint32_t)std::min<uint64_t>(prm.samples, groups_per_batch);
    uint32_t mt_per_batch = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::max<uint64_t>(1, total_mt), groups_per_batch / smp_per_batch));
    // equal shares of the micro-tiles: full batches and a remainder leave the last launches of a render half empty (at 8 GPUs
    // a rank of the 4K bench has ~3.6 default batches)
    if (total_mt > mt_per_batch) {
        const uint64_t nb = (total_mt + mt_per_batch - 1) / mt_per_batch;
        mt_per_batch = (uint32_t)((total_mt + nb - 1) / nb);
    }
    // progress (ipt_set_progress): an event after (at most 200 of) the batches; the host reports how many have finished
    const ipt_progress_fn progress = prm.rank == 0 ? g_progress_fn.load() : nullptr;
    struct BatchEvents {
        std::vector<cudaEvent_t> ev; size_t done = 0;
        ~BatchEvents() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
    } marks;
    const uint64_t n_batches_total = total_mt ? ((total_mt + mt_per_batch - 1) / mt_per_batch) * ((prm.samples + smp_per_batch - 1) / smp_per_batch) : 0;
    const uint64_t mark_every = std::max<uint64_t>(1, (n_batches_total + 199) / 200);
    auto report = [&]() {
        while (marks.done < marks.ev.size() && cudaEventQuery(marks.ev[marks.done]) == cudaSuccess) marks.done++;
        cudaGetLastError();   // cudaErrorNotReady is not an error
        const double total_marks = (double)((n_batches_total + mark_every - 1) / mark_every);
        progress(total_marks > 0 ? std::min(1.0, (double)marks.done / total_marks) : 1.0, g_progress_user.load());
    };
    // Renderer.cu:36-39: when width and height are both <= BLOCK_SIZE (22) every reference thread gets an empty pixel
    // rectangle and the frame stays black.  Kept: such frames are resolved from zeroed accumulators without tracing.
    const bool tiny = c->W <= 22 && c->H <= 22;
    for (uint64_t bi = 0; bi < n_batches_total && !tiny; bi++) {
        const uint64_t s_blocks = (prm.samples + smp_per_batch - 1) / smp_per_batch;
        kp.base_mt = (uint32_t)((bi / s_blocks) * mt_per_batch);
        kp.mt_count = (uint32_t)std::min<uint64_t>(mt_per_batch, total_mt - kp.base_mt);
        kp.base_sample = (uint32_t)((bi % s_blocks) * smp_per_batch);
        kp.n_first = kp.mt_count * std::min<uint32_t>(smp_per_batch, prm.samples - kp.base_sample) * 32u;
        CK(cudaMemsetAsync(c->counters, 0, N_COUNTERS * sizeof(uint32_t), c->stream));
        if (use_split) {
            int rc = launch_split_batch(c, (KParams<float>&)kp, prm.max_depth, cap, smem, split_grids, &launches);
            if (rc) return rc;
            k_batch_stats<<<1, 32, 0, c->stream>>>(c->counters, c->traced);
            launches++;
            batches++;
            if (progress && batches % mark_every == 0) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); CK(cudaEventRecord(e, c->stream));
                marks.ev.push_back(e);
                report();
            }
            continue;
        }
        const uint32_t n_passes = passes_per_batch;
        kp.n_passes = n_passes;
        for (uint32_t pass = 0; pass < n_passes; pass++) {
            const uint32_t d = pass;                       // the depth of the pass, except for the fast kernel (fast_schedule)
            if (pass_times) { cudaEvent_t e; CK(cudaEventCreate(&e)); CK(cudaEventRecord(e, c->stream)); pass_events.push_back(e); }
            kp.depth = d;
            kp.pass = pass;
            kp.fast_k = fast_k;
            kp.qin = Queue{c->q[(pass + 1) & 1], cap};
            kp.qout = Queue{c->q[pass & 1], cap};
            int rc;
            if (use_fast) {
                const KParams<float>& kf = (const KParams<float>&)kp;
                // box rooms (fast_shape): the scan is straight-line code, one instantiation per sphere count
                // (Russian roulette, an extension, has its own instantiations: the default ones carry no code for it)
#define IPT_FAST_BY_SHAPE(FIRST, RR, GRID)                                                             \
    switch (shape) {                                                                                   \
        case 1: rc = launch_bounce_fast<FIRST, 1, RR>(c, kf, GRID); break;                             \
        case 2: rc = launch_bounce_fast<FIRST, 2, RR>(c, kf, GRID); break;                             \
        case 3: rc = launch_bounce_fast<FIRST, 3, RR>(c, kf, GRID); break;                             \
        case 4: rc = launch_bounce_fast<FIRST, 4, RR>(c, kf, GRID); break;                             \
        case 5: rc = launch_bounce_fast<FIRST, 5, RR>(c, kf, GRID); break;                             \
        case -1: rc = launch_bounce_fast<FIRST, -1, RR>(c, kf, GRID); break;                           \
        default: rc = launch_bounce_fast<FIRST, 0, RR>(c, kf, GRID); break;                            \
    }
                const bool roulette = (kf.flags & 0x8u) != 0;
                if (pass == 0 && roulette) IPT_FAST_BY_SHAPE(true, true, &grid_first)
                else if (pass == 0) IPT_FAST_BY_SHAPE(true, false, &grid_first)
                else if (roulette) IPT_FAST_BY_SHAPE(false, true, &grid_next)
                else IPT_FAST_BY_SHAPE(false, false, &grid_next)
#undef IPT_FAST_BY_SHAPE
            }
            else if (defer && d == 0) rc = bvh ? launch_bounce<R, MODE_BVH, true, true>(c, kp, smem, &grid_first) : launch_bounce<R, MODE_BRUTE, true, true>(c, kp, smem, &grid_first);
            else if (defer) rc = bvh ? launch_bounce<R, MODE_BVH, false, true>(c, kp, smem, &grid_next) : launch_bounce<R, MODE_BRUTE, false, true>(c, kp, smem, &grid_next);
            else if (d == 0) rc = bvh ? launch_bounce<R, MODE_BVH, true>(c, kp, smem, &grid_first) : launch_bounce<R, MODE_BRUTE, true>(c, kp, smem, &grid_first);
            else rc = bvh ? launch_bounce<R, MODE_BVH, false>(c, kp, smem, &grid_next) : launch_bounce<R, MODE_BRUTE, false>(c, kp, smem, &grid_next);
            if (rc) return rc;
            launches++;
            if (pass_times && clock_probe && pass_events.size() <= 4096) k_clock_probe<<<1, 32, 0, c->stream>>>(clock_probe + pass_events.size() - 1);
            if (sync_passes) CK(cudaStreamSynchronize(c->stream));
        }
        k_batch_stats<<<1, 32, 0, c->stream>>>(c->counters, c->traced);
        launches++;
        batches++;
        if (progress && batches % mark_every == 0) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); CK(cudaEventRecord(e, c->stream));
            marks.ev.push_back(e);
            report();
        }
    }
    if (pass_times) { cudaEvent_t e; CK(cudaEventCreate(&e)); CK(cudaEventRecord(e, c->stream)); pass_events.push_back(e); }
    // resolve (and gather: dst may be a peer GPU's frame)
    ResolveParams rp;
    rp.frame = c->frame; rp.dst32 = c->gather32; rp.dst64 = c->gather64; rp.tile_ids = c->tile_ids;
    rp.n_tiles_local = kp.n_tiles_local; rp.tiles_x = tiles_x; rp.tile_w = tile_w; rp.tile_h = tile_h; rp.W = c->W; rp.H = c->H;
    rp.inv = 1.0 / (scale * (double)prm.samples); rp.float_accum = float_accum ? 1u : 0u;
    if (float_accum) rp.inv = 1.0 / (double)prm.samples;
    k_resolve<<<c->sm_count * 4, 256, 0, c->stream>>>(rp);
    launches++;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaGetLastError());
    if (progress) {   // wait for the device while reporting the batches it finishes (20 reports a second)
        while (cudaEventQuery(c->ev1) == cudaErrorNotReady) { report(); std::this_thread::sleep_for(std::chrono::milliseconds(50)); }
        cudaGetLastError();
    }
    CK(cudaStreamSynchronize(c->stream));
    if (progress) report();
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (pass_times && pass_events.size() > 1) {
        std::vector<double> per_depth(passes_per_batch, 0.0);
        for (size_t i = 0; i + 1 < pass_events.size(); i++) {
            float t = 0;
            cudaEventElapsedTime(&t, pass_events[i], pass_events[i + 1]);
            per_depth[i % passes_per_batch] += t;
        }
        std::fprintf(stderr, "[ipt] pass times over %llu batch(es), ms per pass:", (unsigned long long)batches);
        for (uint32_t d = 0; d < passes_per_batch; d++) std::fprintf(stderr, " %.3f", per_depth[d]);
        std::fprintf(stderr, "  (total %.3f)\n", ms);
        if (use_fast) {   // the schedule the device chose for the last batch: bounces per pass, and queue lengths read
            std::vector<uint32_t> cnt(N_COUNTERS);
            CK(cudaMemcpy(cnt.data(), c->counters, N_COUNTERS * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            std::fprintf(stderr, "[ipt] last batch, bounces per pass (queue length read):");
            for (uint32_t q = 0; q < passes_per_batch; q++) std::fprintf(stderr, " %u (%u)", cnt[WORK_EXTEND + q], cnt[CNT + q]);
            std::fprintf(stderr, "\n");
        }
        if (clock_probe) {
            std::vector<double> mhz(passes_per_batch, 0.0); std::vector<int> cnt(passes_per_batch, 0);
            for (size_t i = 0; i + 1 < pass_events.size() && i < 4096; i++) { mhz[i % passes_per_batch] += clock_probe[i]; cnt[i % passes_per_batch]++; }
            std::fprintf(stderr, "[ipt] SM MHz after each pass:");
            for (uint32_t d = 0; d < passes_per_batch; d++) std::fprintf(stderr, " %.0f", cnt[d] ? mhz[d] / cnt[d] : 0.0);
            std::fprintf(stderr, "\n");
        }
    }
    unsigned long long traced[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CK(cudaMemcpy(traced, c->traced, 64, cudaMemcpyDeviceToHost));
    c->last.node_steps = traced[2]; c->last.box_tests = traced[3]; c->last.leaf_steps = traced[4];
    c->last.sphere_tests = traced[5]; c->last.rect_tests = traced[6];
    const size_t record_bytes = (sizeof(R) == 4 ? 48 : 96) + (defer ? (sizeof(R) == 4 ? 16 : 32) : 0);
    c->last.queue_bytes = traced[1] * record_bytes;
    c->last.render_ms = ms; c->last.traced_bounces = traced[0]; c->last.kernel_launches = launches; c->last.batches = batches;
    uint64_t npx = 0;
    for (uint32_t t : c->tile_host) {
        const uint32_t x0 = (t % tiles_x) * tile_w, z0 = (t / tiles_x) * tile_h;
        npx += (uint64_t)(std::min(c->W, x0 + tile_w) - x0) * (std::min(c->H, z0 + tile_h) - z0);
    }
    c->last.samples = npx * prm.samples;
    c->last.active_pixels = c->active_pixels;
    if (st) *st = c->last;
    return IPT_OK;
}

extern "C" int ipt_ctx_render(ipt_ctx* c, const ipt_params* prm, ipt_stats* st)
{
    if (!c || !prm) { set_err("ipt_ctx_render: null argument"); return IPT_ERR_BAD_ARGUMENT; }
    if (!c->have_scene) { set_err("ipt_ctx_render: no scene set"); return IPT_ERR_BAD_ARGUMENT; }
    if (prm->samples < 1 || prm->samples > 65535 || prm->max_depth < 1 || prm->max_depth > 255) {
        set_err("ipt_ctx_render: samples must be 1..65535 and max_depth 1..255");
        return IPT_ERR_BAD_ARGUMENT;
    }
    const uint32_t tile_w = prm->tile_w ? prm->tile_w : 64, tile_h = prm->tile_h ? prm->tile_h : 32;
    const uint32_t world = prm->world ? prm->world : 1;
    if (tile_w % 8 || tile_h % 4 || prm->rank >= world) { set_err("ipt_ctx_render: tile size must be a multiple of 8x4, rank < world"); return IPT_ERR_BAD_ARGUMENT; }
    CK(cudaSetDevice(c->device));
    const auto t_call = std::chrono::steady_clock::now();
    uint32_t tiles_x = 0;
    int rc = build_tiles(c, tile_w, tile_h, prm->rank, world, &tiles_x);
    if (rc) return rc;
    const uint64_t mt_key = 1 + ((uint64_t)tile_w << 48 | (uint64_t)tile_h << 32 | (uint64_t)prm->rank << 16 | world) + (std::getenv("IPT_NO_CULL") ? (1ull << 62) : 0);
    rc = build_active_microtiles(c, tile_w, tile_h, tiles_x, mt_key);
    if (rc) return rc;
    const auto t_tiles = std::chrono::steady_clock::now();
    rc = (prm->flags & IPT_FLAG_FP64) ? render_typed<double>(c, *prm, tile_w, tile_h, tiles_x, st)
                                      : render_typed<float>(c, *prm, tile_w, tile_h, tiles_x, st);
    if (std::getenv("IPT_VERBOSE")) {   // host view of one resident render: tile lists, then enqueue + wait, against the kernels' own time
        const auto t_end = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[ipt] rank %u render: tile and micro-tile lists %.2f ms, enqueue + wait %.2f ms, kernels %.2f ms\n", prm->rank,
                     std::chrono::duration<double, std::milli>(t_tiles - t_call).count(),
                     std::chrono::duration<double, std::milli>(t_end - t_tiles).count(), c->last.render_ms);
    }
    return rc;
}

extern "C" void* ipt_alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { set_err(std::string("ipt_alloc_pinned: ") + cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    return p;
}

extern "C" void ipt_free_pinned(void* p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int ipt_ctx_download(ipt_ctx* c, float* out32, double* out64)
{
    if (!c || !c->have_scene) { set_err("ipt_ctx_download: no frame"); return IPT_ERR_BAD_ARGUMENT; }
    CK(cudaSetDevice(c->device));
    const size_t px = c->frame_pixels;
    const size_t need = (out32 ? px * 12 : 0) + (out64 ? px * 24 : 0);
    // Straight into the caller's buffer: for pageable memory the driver pipelines its own pinned staging, which is
    // faster than a pinned bounce buffer plus a single-threaded host memcpy of a 100 MB 4K frame.
    CK(cudaEventRecord(c->ev0, c->stream));
    if (out32) CK(cudaMemcpyAsync(out32, c->out32, px * 12, cudaMemcpyDeviceToHost, c->stream));
    if (out64) CK(cudaMemcpyAsync(out64, c->out64, px * 24, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last.download_ms = ms;
    c->last.d2h_bytes = need;
    return IPT_OK;
}

extern "C" int ipt_ctx_download_rgb8(ipt_ctx* c, uint8_t* out8)
{
    if (!c || !c->have_scene || !out8) { set_err("ipt_ctx_download_rgb8: no frame"); return IPT_ERR_BAD_ARGUMENT; }
    CK(cudaSetDevice(c->device));
    const size_t n = c->frame_pixels * 3;
    if (c->out8_bytes < n) {
        cudaFree(c->out8); c->out8 = nullptr; c->out8_bytes = 0;
        CK(cudaMalloc(&c->out8, n));
        c->out8_bytes = n;
    }
    CK(cudaEventRecord(c->ev0, c->stream));
    k_to_rgb8<<<c->sm_count * 8, 256, 0, c->stream>>>(c->out32, c->out8, n);
    CK(cudaMemcpyAsync(out8, c->out8, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last.download_ms = ms;
    c->last.d2h_bytes = n;
    c->last.kernel_launches += 1;
    return IPT_OK;
}

// ------------------------------------------------------------------------------------------------ gather targets
extern "C" int ipt_ctx_export_frame(ipt_ctx* c, void* handle64)
{
    if (!c || !c->out32 || !handle64) { set_err("ipt_ctx_export_frame: no frame"); return IPT_ERR_BAD_ARGUMENT; }
    CK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->out32));
    static_assert(sizeof(h) == 64, "CUDA IPC handle is 64 bytes");
    std::memcpy(handle64, &h, 64);
    return IPT_OK;
}

extern "C" int ipt_ctx_set_gather_target_ipc(ipt_ctx* c, const void* handle64)
{
    if (!c || !handle64) { set_err("ipt_ctx_set_gather_target_ipc: null argument"); return IPT_ERR_BAD_ARGUMENT; }
    CK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    if (c->ipc_mapped) cudaIpcCloseMemHandle(c->ipc_mapped);
    c->ipc_mapped = p;
    c->gather32 = (float*)p;
    // the owner's fp64 frame follows its fp32 frame in the same allocation (same scene, hence the same frame size, on both sides)
    c->gather64 = c->frame_pixels ? (double*)((char*)p + frame64_offset(c->frame_pixels)) : nullptr;
    return IPT_OK;
}

extern "C" int ipt_ctx_set_gather_target(ipt_ctx* c, ipt_ctx* owner)
{
    if (!c || !owner || !owner->out32) { set_err("ipt_ctx_set_gather_target: owner has no frame"); return IPT_ERR_BAD_ARGUMENT; }
    if (c == owner || c->device == owner->device) { c->gather32 = owner->out32; c->gather64 = owner->out64; return IPT_OK; }
    CK(cudaSetDevice(c->device));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, c->device, owner->device));
    if (!can) { set_err("no peer access between the two devices"); return IPT_ERR_CUDA; }
    cudaError_t e = cudaDeviceEnablePeerAccess(owner->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_err(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return IPT_ERR_CUDA; }
    cudaGetLastError();
    c->gather32 = owner->out32; c->gather64 = owner->out64;
    return IPT_OK;
}

// ------------------------------------------------------------------------------------------------ one-shot render
static int render_impl(const ipt_scene* scene, const ipt_params* params, int n_gpus, float* out32, double* out64, uint8_t* out8, ipt_stats* stats)
{
    if (!scene || !params) { set_err("ipt_render: null argument"); return IPT_ERR_BAD_ARGUMENT; }
    const int ndev = ipt_device_count();
    if (ndev <= 0) { set_err("CUDA capable device not found! Cannot continue"); return IPT_ERR_NO_DEVICE; }
    if (n_gpus < 1) n_gpus = 1;
    if (n_gpus > ndev || n_gpus > 8) { set_err("ipt_render: n_gpus exceeds the devices present (max 8)"); return IPT_ERR_BAD_ARGUMENT; }
    std::vector<ipt_ctx*> ctx(n_gpus, nullptr);
    std::vector<int> rcs(n_gpus, 0);
    std::vector<ipt_stats> sts(n_gpus);
    int rc = IPT_OK;
    // IPT_VERBOSE: wall time of the phases of a one-shot render (what the reference's measure() region contains)
    const bool verbose = std::getenv("IPT_VERBOSE") != nullptr;
    auto T = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        if (verbose) std::fprintf(stderr, "[ipt] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - T).count());
        T = now;
    };
    for (int g = 0; g < n_gpus && rc == IPT_OK; g++) {
        ctx[g] = ipt_ctx_create(g);
        if (!ctx[g]) rc = IPT_ERR_CUDA;
    }
    lap("contexts");
    bool host_merge = false;
    if (rc == IPT_OK) {
        // upload (one host thread per GPU), set gather targets, render
        std::vector<std::thread> th;
        for (int g = 0; g < n_gpus; g++) th.emplace_back([&, g] { rcs[g] = ipt_ctx_set_scene(ctx[g], scene); });
        for (auto& t : th) t.join();
        for (int g = 0; g < n_gpus; g++) if (rcs[g]) rc = rcs[g];
    }
    lap("scene upload");
    if (rc == IPT_OK) {
        for (int g = 1; g < n_gpus; g++)
            if (ipt_ctx_set_gather_target(ctx[g], ctx[0]) != IPT_OK) host_merge = true;
        if (host_merge) for (int g = 1; g < n_gpus; g++) { ctx[g]->gather32 = ctx[g]->out32; ctx[g]->gather64 = ctx[g]->out64; }
        std::vector<std::thread> th;
        for (int g = 0; g < n_gpus; g++)
            th.emplace_back([&, g] {
                ipt_params p = *params;
                p.rank = (uint32_t)g; p.world = (uint32_t)n_gpus;
                p.reserved[0] |= 1u;   // one-shot: the ray queues live for this call only (render_typed's default batch size)
                rcs[g] = ipt_ctx_render(ctx[g], &p, &sts[g]);
            });
        for (auto& t : th) t.join();
        for (int g = 0; g < n_gpus; g++) if (rcs[g]) rc = rcs[g];
    }
    lap("queues + kernels");
    if (rc == IPT_OK && out8) {
        if (host_merge && n_gpus > 1) { set_err("ipt_render_rgb8: needs peer access between the GPUs"); rc = IPT_ERR_CUDA; }
        else rc = ipt_ctx_download_rgb8(ctx[0], out8);
    }
    if (rc == IPT_OK && (out32 || out64)) {
        rc = ipt_ctx_download(ctx[0], out32, out64);
        if (rc == IPT_OK && host_merge && n_gpus > 1) {
            // no peer access: fetch every other GPU's frame and copy its tiles on the host
            const size_t px = (size_t)scene->width * scene->height;
            std::vector<float> t32(out32 ? px * 3 : 0);
            std::vector<double> t64(out64 ? px * 3 : 0);
            const uint32_t tw = params->tile_w ? params->tile_w : 64, thh = params->tile_h ? params->tile_h : 32;
            const uint32_t tiles_x = (scene->width + tw - 1) / tw;
            for (int g = 1; g < n_gpus && rc == IPT_OK; g++) {
                rc = ipt_ctx_download(ctx[g], out32 ? t32.data() : nullptr, out64 ? t64.data() : nullptr);
                for (uint32_t z = 0; z < scene->height && rc == IPT_OK; z++)
                    for (uint32_t x = 0; x < scene->width; x++)
                        if (ipt_tile_owner(x / tw, z / thh, tiles_x, n_gpus) == (uint32_t)g) {
                            const size_t i = ((size_t)z * scene->width + x) * 3;
                            for (int k = 0; k < 3; k++) { if (out32) out32[i + k] = t32[i + k]; if (out64) out64[i + k] = t64[i + k]; }
                        }
            }
        }
    }
    if (rc == IPT_OK && stats) {
        *stats = sts[0];
        stats->samples = 0; stats->traced_bounces = 0; stats->kernel_launches = 0; stats->batches = 0; stats->render_ms = 0;
        stats->queue_bytes = 0;
        stats->node_steps = stats->box_tests = stats->leaf_steps = stats->sphere_tests = stats->rect_tests = 0;
        for (int g = 0; g < n_gpus; g++) {
            stats->queue_bytes += sts[g].queue_bytes;
            stats->node_steps += sts[g].node_steps; stats->box_tests += sts[g].box_tests; stats->leaf_steps += sts[g].leaf_steps;
            stats->sphere_tests += sts[g].sphere_tests; stats->rect_tests += sts[g].rect_tests;
            stats->samples += sts[g].samples; stats->traced_bounces += sts[g].traced_bounces;
            stats->kernel_launches += sts[g].kernel_launches; stats->batches += sts[g].batches;
            if (g > 0) stats->active_pixels += sts[g].active_pixels;
            stats->render_ms = std::max(stats->render_ms, sts[g].render_ms);
            stats->per_gpu_render_ms[g] = sts[g].render_ms; stats->per_gpu_bounces[g] = sts[g].traced_bounces;
        }
        stats->download_ms = ctx[0]->last.download_ms; stats->d2h_bytes = ctx[0]->last.d2h_bytes;
    }
    lap("download");
    for (int g = 0; g < n_gpus; g++) ipt_ctx_destroy(ctx[g]);
    lap("release");
    return rc;
}

extern "C" int ipt_render(const ipt_scene* scene, const ipt_params* params, int n_gpus, float* out32, double* out64, ipt_stats* stats)
{
    return render_impl(scene, params, n_gpus, out32, out64, nullptr, stats);
}

extern "C" int ipt_render_rgb8(const ipt_scene* scene, const ipt_params* params, int n_gpus, uint8_t* out8, ipt_stats* stats)
{
    if (!out8) { set_err("ipt_render_rgb8: null output"); return IPT_ERR_BAD_ARGUMENT; }
    return render_impl(scene, params, n_gpus, nullptr, nullptr, out8, stats);
}

// ------------------------------------------------------------------------------------------------ trace (tests)
extern "C" int ipt_ctx_trace(ipt_ctx* c, const double* rays, uint32_t n, uint32_t flags, int32_t* out_obj, double* out_t)
{
    if (!c || !c->have_scene || !rays || !out_obj || !out_t) { set_err("ipt_ctx_trace: bad argument"); return IPT_ERR_BAD_ARGUMENT; }
    if (n == 0) return IPT_OK;
    CK(cudaSetDevice(c->device));
    struct Temps {                  // released on every return path
        double* rays = nullptr; int32_t* obj = nullptr; double* t = nullptr;
        ~Temps() { cudaFree(rays); cudaFree(obj); cudaFree(t); }
    } tmp;
    double*& d_rays = tmp.rays; int32_t*& d_obj = tmp.obj; double*& d_t = tmp.t;
    if (flags & IPT_FLAG_FP64) { int rc64 = upload_fp64(c); if (rc64) return rc64; }
    CK(cudaMalloc(&d_rays, (size_t)n * 48)); CK(cudaMalloc(&d_obj, (size_t)n * 4)); CK(cudaMalloc(&d_t, (size_t)n * 8));
    CK(cudaMemcpyAsync(d_rays, rays, (size_t)n * 48, cudaMemcpyHostToDevice, c->stream));
    const bool bvh = c->n_nodes > 0, f64 = (flags & IPT_FLAG_FP64) != 0;
    if (!f64 && bvh && c->bslot && !std::getenv("IPT_GENERIC_KERNEL") && !std::getenv("IPT_FUSED_BVH")) {
        // fp32 + BVH: through the traversal kernel of the split pipeline itself (k_extend_bvh, or k_extend_cw with IPT_BVH8), fed
        // with a one-pass ray queue - what a render runs, not a restatement of it
        struct Q {
            uint4* rays = nullptr; uint2* hits = nullptr;
            ~Q() { cudaFree(rays); cudaFree(hits); }
        } q;
        CK(cudaMalloc(&q.rays, (size_t)n * 48)); CK(cudaMalloc(&q.hits, (size_t)n * 8));
        std::vector<uint4> rec((size_t)n * 3);
        auto fb = [](double x) { const float f = (float)x; uint32_t u; std::memcpy(&u, &f, 4); return u; };
        for (uint32_t i = 0; i < n; i++) {
            const double* r = rays + 6 * (size_t)i;
            rec[i] = make_uint4(fb(r[0]), fb(r[1]), fb(r[2]), fb(r[3]));
            rec[(size_t)n + i] = make_uint4(fb(r[4]), fb(r[5]), 0u, 0u);
            rec[2 * (size_t)n + i] = make_uint4(0u, 0u, 0u, NO_OBJECT);       // no flags, starts on no object
        }
        CK(cudaMemcpyAsync(q.rays, rec.data(), rec.size() * 16, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync(c->counters, 0, N_COUNTERS * sizeof(uint32_t), c->stream));
        CK(cudaMemcpyAsync(c->counters + CNT, &n, 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync(c->traced, 0, 64, c->stream));
        KParams<float> kp;
        std::memset(&kp, 0, sizeof(kp));
        kp.sc.geom = (const R4<float>*)c->geom32; kp.sc.mat = (const R4<float>*)c->mat32; kp.sc.slot_obj = c->slot_obj;
        kp.sc.n_slots = c->n_slots; kp.sc.n_spheres = c->n_spheres; kp.sc.n_objects = c->n_objects;
        kp.sc.nodes = c->nodes; kp.sc.n_nodes = c->n_nodes; kp.sc.bslot = c->bslot;
        kp.counters = c->counters; kp.traced = c->traced; kp.hits = q.hits; kp.depth = 0;
        kp.qin = Queue{q.rays, n};
        kp.wide = c->wide;
        kp.grid_cells = c->grid_cells; kp.grid_refs = c->grid_refs; kp.grid = c->grid_hd;
        kp.descend_min = 12; kp.refill_min = kp.grid_cells ? 12 : 8; kp.leaf_min = kp.grid_cells ? 16 : 8;
        const size_t smem_top = (size_t)std::min<uint32_t>(c->n_nodes, BVH_TOP_NODES) * 64;
        int grids[4] = {0, 0, 0, 0};
        int rc = prepare_extend(c, kp, smem_top, grids);
        if (rc) return rc;
        if (kp.grid_cells) k_extend_grid<<<grids[3], BLOCK_THREADS, 0, c->stream>>>(kp);
        else if (kp.wide) k_extend_cw<<<grids[2], BLOCK_THREADS, CW_SMEM, c->stream>>>(kp);
        else k_extend_bvh<<<grids[0], BLOCK_THREADS, smem_top, c->stream>>>(kp);
        CK(cudaGetLastError());
        std::vector<uint2> hits(n);
        std::vector<uint32_t> slot_obj(c->n_slots);
        CK(cudaMemcpyAsync(hits.data(), q.hits, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(slot_obj.data(), c->slot_obj, (size_t)c->n_slots * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (uint32_t i = 0; i < n; i++) {
            float t; std::memcpy(&t, &hits[i].x, 4);
            const bool hit = hits[i].y != NO_OBJECT && hits[i].y < c->n_slots;
            out_obj[i] = hit ? (int32_t)(slot_obj[hits[i].y] & ~RECT_BIT) : -1;
            out_t[i] = hit ? (double)t : 0.0;
        }
        return IPT_OK;
    }
    const size_t esz = f64 ? 32 : 16;
    const size_t smem = bvh ? (size_t)std::min<uint32_t>(c->n_nodes, BVH_TOP_NODES) * 64 : (size_t)c->n_slots * 4 * esz + ((size_t)c->n_slots + 3) / 4 * 16;
    const int grid = c->sm_count * 2;
    auto fill = [&](auto& sv, auto tag) {
        using R = decltype(tag);
        sv.geom = (const R4<R>*)(f64 ? c->geom64 : c->geom32); sv.mat = (const R4<R>*)(f64 ? c->mat64 : c->mat32);
        sv.slot_obj = c->slot_obj; sv.n_slots = c->n_slots; sv.n_spheres = c->n_spheres; sv.n_objects = c->n_objects;
        sv.nodes = c->nodes; sv.n_nodes = c->n_nodes; sv.bslot = c->bslot;
        sv.lights = nullptr; sv.n_lights = 0;
    };
    if (f64) {
        SceneView<double> sv; fill(sv, double());
        if (bvh) { CK(cudaFuncSetAttribute(k_trace<double, MODE_BVH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_trace<double, MODE_BVH><<<grid, BLOCK_THREADS, smem, c->stream>>>(sv, d_rays, n, d_obj, d_t); }
        else { CK(cudaFuncSetAttribute(k_trace<double, MODE_BRUTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_trace<double, MODE_BRUTE><<<grid, BLOCK_THREADS, smem, c->stream>>>(sv, d_rays, n, d_obj, d_t); }
    } else {
        SceneView<float> sv; fill(sv, float());
        if (bvh) { CK(cudaFuncSetAttribute(k_trace<float, MODE_BVH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_trace<float, MODE_BVH><<<grid, BLOCK_THREADS, smem, c->stream>>>(sv, d_rays, n, d_obj, d_t); }
        else { CK(cudaFuncSetAttribute(k_trace<float, MODE_BRUTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_trace<float, MODE_BRUTE><<<grid, BLOCK_THREADS, smem, c->stream>>>(sv, d_rays, n, d_obj, d_t); }
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_obj, d_obj, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(out_t, d_t, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return IPT_OK;
}
