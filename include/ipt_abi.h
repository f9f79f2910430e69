/* ipt_abi.h — C ABI of the B200-native per-pixel radiance path (libipt_b200.so).
 *
 * Drop-in boundary.  The reference (AdamStudies-PWR/Improved-Path-Tracer) has no plugin/FFI interface;
 * its seam is one C++ call made from main():
 *     tracer::renderer::RenderContoller(SceneData&, uint32_t samples, uint8_t maxDepth)   RenderController.hpp:11-24
 *     std::vector<containers::Vec3> RenderContoller::start()                              RenderController.cu:36-70
 * which allocates the frame, uploads the ObjectData array, launches
 *     cudaMain<<<22,22>>>(Vec3* image, ObjectData* objs, n, W, H, Camera, vecZ, samples, maxDepth)   Renderer.cu:254-265
 * and copies the frame back.  Everything below replaces exactly that call and that kernel:
 *   - ipt_render_objects()  takes the reference's own buffers (ObjectData[] AoS, Camera, W, H, samples, maxDepth)
 *                           and fills a W*H*3 float64 frame in the reference's layout (index z*W+x, row 0 = +vecZ
 *                           edge) — what RenderContoller::start() would call (binding shown in INTEGRATION.md);
 *   - ipt_render()          the same through the flattened structure-of-arrays scene (ipt_scene) the C++ host
 *                           code of this repository produces (include/ipt_host.h);
 *   - ipt_ctx_*             the same split into create / upload / render / download, for callers that keep the
 *                           scene resident (bench.py device-resident timing, one process per GPU under torchrun).
 * Plain pointers and sizes only; the caller owns every buffer it passes; no exceptions cross the boundary;
 * every function returns 0 on success or a negative ipt_status, and ipt_last_error() gives the text
 * (the reference prints CUDA errors and carries on, RenderController.cu:20-27; main polls cudaGetLastError,
 * main.cu:52-56 — the `tracer` host program of this repository maps a non-zero status to that exit code 1).
 * There is no CPU fallback: without a CUDA device every entry point fails with IPT_ERR_NO_DEVICE.
 */
#ifndef IPT_ABI_H
#define IPT_ABI_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define IPT_ABI_VERSION 3

typedef enum ipt_status {
    IPT_OK = 0,
    IPT_ERR_NO_DEVICE = -1,   /* "CUDA capable device not found! Cannot continue"  CudaUtils.cu:13-17 */
    IPT_ERR_BAD_ARGUMENT = -2,
    IPT_ERR_CUDA = -3,        /* a CUDA runtime call or kernel failed; text in ipt_last_error() */
    IPT_ERR_OUT_OF_MEMORY = -4
} ipt_status;

/* A node of the 2-wide bounding volume hierarchy: 64 bytes = four 128-bit loads.
 *   lo0/hi0, lo1/hi1 : boxes of child 0 and child 1 (fp32, padded outwards by the builder)
 *   child[i] >= 0    : index of an inner node
 *   child[i] <  0    : leaf; ~child[i] = first slot, count[i] = number of primitive slots (leaf order) */
typedef struct ipt_bvh_node {
    float lo0[3], hi0[3];
    float lo1[3], hi1[3];
    int32_t child[2];
    uint32_t count[2];
} ipt_bvh_node;

/* Flattened scene: one entry per primitive, primitives of each kind stored contiguously, every array fp64
 * (the library keeps an fp32 copy for the fast path).  Produced by ipt_host_load_scene() / ipt_host_from_objects() (include/ipt_host.h)
 * from the reference's scenes/<name>.json schema; object order is preserved in *_object[] because the reference's
 * nearest-hit scan lets the lowest object index win ties (Renderer.cu:235).
 *
 * Sphere i      : sphere_cxyzr[4i..]   = centre xyz, radius                       (Sphere.cu:25-39)
 * Rectangle j   : rect_plane[4j..]     = unit normal n = normalize(north x east), D = n.centre   (Plane.cu:36,47-58)
 *                 rect_u[4j..]         = unit in-plane axis perpendicular to `east`,  u.centre
 *                 rect_v[4j..]         = unit in-plane axis perpendicular to `north`, v.centre
 *                 rect_bounds[4j..]    = {u_lo, u_hi, v_lo, v_hi}: a plane hit P is inside iff
 *                                        u_lo <= |P.u - u.centre| <= u_hi and v_lo <= |P.v - v.centre| <= v_hi,
 *                                        which is Plane.cu:87-100 (sum of distances to opposite edge lines ==
 *                                        side length +- 1e-4) solved for the hit position.
 * Material k    : per OBJECT (JSON order): color, emission (Vec3 each), reflection 0/1/2 (EReflectionType.hpp:6-11).
 */
typedef struct ipt_scene {
    uint32_t width, height;          /* SceneData.cpp:98-111 */
    double cam_origin[3];            /* camera.position                                        */
    double cam_dir[3];               /* camera.direction, normalised (SceneData.cpp:143-145)   */
    double cam_orient[3];            /* camera.orientation, normalised                         */

    uint32_t n_objects;              /* = n_spheres + n_rects                                  */
    uint32_t n_spheres;
    uint32_t n_rects;
    uint32_t reserved0;
    const double* sphere_cxyzr;      /* [n_spheres*4]  */
    const uint32_t* sphere_object;   /* [n_spheres]    object (JSON) index of each sphere      */
    const double* rect_plane;        /* [n_rects*4]    */
    const double* rect_u;            /* [n_rects*4]    */
    const double* rect_v;            /* [n_rects*4]    */
    const double* rect_bounds;       /* [n_rects*4]    */
    const uint32_t* rect_object;     /* [n_rects]      */
    const double* mat_color;         /* [n_objects*3]  */
    const double* mat_emission;      /* [n_objects*3]  */
    const int32_t* mat_reflection;   /* [n_objects]    */

    /* Optional BVH (n_bvh_nodes == 0: every ray tests every primitive, staged in shared memory).
     * Slot s of the leaf order refers to primitive bvh_slot_prim[s]: bit 31 set = rectangle, low bits = index
     * into the sphere / rectangle arrays above. */
    uint32_t n_bvh_nodes;
    uint32_t n_bvh_slots;
    const ipt_bvh_node* bvh_nodes;   /* [n_bvh_nodes], node 0 = root */
    const uint32_t* bvh_slot_prim;   /* [n_bvh_slots]  */

    /* Optional uniform grid over the same slots (grid_res[0] == 0: none; needs the BVH above, which stays the structure of
     * the fp64 parity kernels).  Scenes whose primitives are small against their spacing and spread evenly (BASELINE config 5)
     * are walked far cheaper cell by cell than through a hierarchy: ~13 cells instead of ~49 inner nodes per ray on the
     * 1M-primitive scene.  Built by ipt_host_build_bvh() when the scene qualifies.
     *   cell (ix,iy,iz) = ix + res_x * (iy + res_y * iz) covers grid_lo + (i .. i+1) * grid_cell per axis;
     *   its primitives are the slots grid_refs[grid_cell_start[c] .. grid_cell_start[c+1]): every slot whose (padded)
     *   bounding box overlaps the cell;
     *   grid_big[]: slots too large for that (walls, big lights) - tested for every ray, never entered in a cell. */
    uint32_t grid_res[3];
    uint32_t n_grid_big;
    float grid_lo[3];
    float grid_cell[3];
    uint32_t n_grid_refs;
    uint32_t reserved1;
    const uint32_t* grid_cell_start; /* [res_x*res_y*res_z + 1] */
    const uint32_t* grid_refs;       /* [n_grid_refs]  */
    const uint32_t* grid_big;        /* [n_grid_big]   */
} ipt_scene;

/* flags */
#define IPT_FLAG_FP64          0x1u  /* compute in fp64 with the reference's literal self-hit tests (parity mode)    */
#define IPT_FLAG_FLOAT_ACCUM   0x4u  /* accumulate with floating-point atomics instead of deterministic fixed point  */
#define IPT_FLAG_RUSSIAN_ROULETTE 0x8u /* extension, OFF for parity: unbiased roulette on throughput from depth >= 3 */
#define IPT_FLAG_STRATIFIED    0x10u /* extension, OFF for parity: the two camera jitters of sample i are stratified on a
                                        floor(sqrt(spp))^2 grid (unbiased: every stratum is sampled uniformly)           */

#define IPT_FLAG_NEXT_EVENT    0x20u /* extension, OFF for parity: next-event estimation towards the emissive spheres at
                                        diffuse hits (one light, one visibility cast per hit); the estimator stays
                                        unbiased: the lobe's density is carried as a weight and a continuation ray does
                                        not count the emission of a sphere that was sampled explicitly              */

typedef struct ipt_params {
    uint32_t samples;        /* per pixel; reference CLI range 4..65535 (InputParser.cpp:14-24), any >= 1 accepted */
    uint32_t max_depth;      /* reference CLI range 3..255; 1..255 accepted                                        */
    uint64_t seed;           /* key of the counter-based generator (the reference's fixed seed is 123456)          */
    uint32_t flags;
    uint32_t tile_w, tile_h; /* multi-GPU tile size in pixels (multiples of 8 and 4); 0 = default 64x32            */
    uint32_t rank, world;    /* this context renders the tiles t with owner(t) == rank out of `world` (world 0 = 1) */
    uint32_t batch_samples;  /* camera rays generated per wavefront batch; 0 = default                             */
    uint32_t reserved[4];    /* set to 0 (reserved[0] bit 0 is used inside the library: one-shot render) */
} ipt_params;

typedef struct ipt_stats {
    uint64_t samples;            /* camera rays generated                                            */
    uint64_t traced_bounces;     /* nearest-hit queries executed (sum over wavefront passes)          */
    uint64_t kernel_launches;    /* kernels launched by the library for this render                   */
    uint64_t batches;
    double render_ms;            /* CUDA-event time of the kernels on the rendering stream            */
    double upload_ms, download_ms;
    uint64_t h2d_bytes, d2h_bytes;
    double per_gpu_render_ms[8]; /* ipt_render() with n_gpus > 1                                      */
    uint64_t per_gpu_bounces[8];
    uint64_t active_pixels;      /* pixels whose camera rays can reach the scene's bounding box; the others are exactly 0
                                    for every sample and no ray is generated for them (they still count in `samples`)  */
    uint64_t queue_bytes;        /* ray-queue bytes written + read back by the wavefront passes (records x record size): the
                                    traffic the design sends through HBM, counted on the device                         */
    /* Traversal work of BVH scenes in fp32 (the split pipeline), counted on the device; 0 for the brute-force scenes, whose
     * work per cast is fixed (every primitive).  bench.py turns these into the algorithmic flops of SURVEY.md §8d. */
    uint64_t node_steps;         /* inner nodes visited                                                */
    uint64_t box_tests;          /* child boxes tested (2 per node of the 2-wide tree, 8 per node of the 8-wide tree) */
    uint64_t leaf_steps;         /* leaves visited                                                     */
    uint64_t sphere_tests;       /* Sphere::intersect evaluations (Sphere.cu:25-39)                    */
    uint64_t rect_tests;         /* Plane::intersect evaluations (Plane.cu:47-100)                     */
} ipt_stats;

/* -- device probe (CudaUtils.cu:8-23) ------------------------------------------------------------------- */
int ipt_abi_version(void);
int ipt_device_count(void);
const char* ipt_device_name(int device);
const char* ipt_last_error(void);

/* -- progress ------------------------------------------------------------------------------------------------
 * The reference prints "\rRendering %.2f%%" from the device after every pixel row (Renderer.cu:105-107).  Here a
 * process-wide hook is called from the host thread that renders rank 0's tiles with the fraction (0..1) of wavefront
 * batches the device has finished: while batches are enqueued and, about 20 times a second, while the call waits for the
 * device.  NULL (the default) switches it off.  The hook must not call back into this library. */
typedef void (*ipt_progress_fn)(double fraction_done, void* user);
void ipt_set_progress(ipt_progress_fn fn, void* user);

/* -- page-locked host buffers ---------------------------------------------------------------------------
 * Optional: every entry point accepts any host pointer.  A frame buffer obtained here is copied at full PCIe
 * rate (the reference returns a pageable std::vector, RenderController.cu:62-68; at 4K the float frame is
 * 99.5 MB).  NULL on failure, text in ipt_last_error(). */
void* ipt_alloc_pinned(size_t bytes);
void ipt_free_pinned(void* p);

/* -- one-shot renders -------------------------------------------------------------------------------------
 * n_gpus = 1, 2, 4 or 8 devices of this process: tiles are interleaved statically over the devices and the
 * finished tiles are written to device 0 over NVLink peer access (no reduction), then copied to the host.
 * out_rgb32 / out_rgb64: W*H*3, row-major z*W+x, either may be NULL. */
int ipt_render(const ipt_scene* scene, const ipt_params* params, int n_gpus, float* out_rgb32, double* out_rgb64,
               ipt_stats* stats);

/* The same with the output stage fused on the device: Image.cpp:19-22's toRgb (clamp(int(x*255),0,255), no gamma) is
 * applied by a kernel and only width*height*3 BYTES cross PCIe. */
int ipt_render_rgb8(const ipt_scene* scene, const ipt_params* params, int n_gpus, uint8_t* out_rgb8, ipt_stats* stats);

/* The reference's own buffers: `objects` is n records of the 144-byte ObjectData layout (ObjectData.hpp:15-31:
 * type@0 radius@8 north@16 east@40 position@64 emission@88 color@112 reflection@136), `camera` the 72-byte
 * Camera (origin, direction, orientation as 9 doubles, Camera.hpp:8-16). */
int ipt_render_objects(const void* objects, uint32_t n_objects, uint32_t width, uint32_t height,
                       const double* camera, uint32_t samples, uint32_t max_depth, int n_gpus, double* out_image);

/* -- resident contexts (one per GPU) ----------------------------------------------------------------------- */
typedef struct ipt_ctx ipt_ctx;
ipt_ctx* ipt_ctx_create(int device);
void ipt_ctx_destroy(ipt_ctx* ctx);
int ipt_ctx_set_scene(ipt_ctx* ctx, const ipt_scene* scene);                 /* host -> device copy of the scene     */
int ipt_ctx_render(ipt_ctx* ctx, const ipt_params* params, ipt_stats* stats); /* kernels only; frame stays in HBM    */
int ipt_ctx_download(ipt_ctx* ctx, float* out_rgb32, double* out_rgb64);     /* device -> host copy of the frame     */
int ipt_ctx_download_rgb8(ipt_ctx* ctx, uint8_t* out_rgb8);                  /* toRgb on the device, bytes to host   */
/* Tiles of other ranks: a context can write its finished tiles into another context's frame (same process: pass
 * the context; other process: pass the 64-byte CUDA IPC handle exported by the owner). */
int ipt_ctx_export_frame(ipt_ctx* ctx, void* handle64);
int ipt_ctx_set_gather_target_ipc(ipt_ctx* ctx, const void* handle64);    /* after ipt_ctx_set_scene (same frame size as the owner) */
int ipt_ctx_set_gather_target(ipt_ctx* ctx, ipt_ctx* owner);
/* Owner of a tile under the static interleaved schedule. */
uint32_t ipt_tile_owner(uint32_t tile_x, uint32_t tile_y, uint32_t tiles_x, uint32_t world);

/* -- function-level access for parity tests: nearest hit of n rays (origin xyz, direction xyz, fp64) through the
 * same device code the renderer uses; self_object = -1.  out_object = object index or -1, out_t = distance. */
int ipt_ctx_trace(ipt_ctx* ctx, const double* rays, uint32_t n_rays, uint32_t flags, int32_t* out_object, double* out_t);

#ifdef __cplusplus
}
#endif
#endif
