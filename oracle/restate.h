/* TEST INFRASTRUCTURE ONLY — the CPU oracle. Nothing in the product path may include, link or call this.
 *
 * C interface of oracle/restate.cpp, a CPU restatement (FP64, scalar) of the per-pixel radiance routine of
 * AdamStudies-PWR/Improved-Path-Tracer (src/renderer/Renderer.cu:88-243 and the scene/cuda objects).
 * Pinned against the reference's own code compiled for the host (oracle/_ref/libref_host.so): see
 * tests/test_oracle_pin.py.  Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 */
#ifndef IPT_ORACLE_RESTATE_H
#define IPT_ORACLE_RESTATE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same field order and offsets as the reference's ObjectData (ObjectData.hpp:15-31), 144 bytes:
 * type@0 radius@8 north@16 east@40 position@64 emission@88 color@112 reflection@136. */
typedef struct or_object {
    int32_t type;        /* 0 = sphere, 1 = plane (finite rectangle)   ObjectData.hpp:9-13 */
    int32_t pad0;
    double radius;
    double north[3];
    double east[3];
    double position[3];
    double emission[3];
    double color[3];
    int32_t reflection;  /* 0 diffuse, 1 specular, 2 refractive        EReflectionType.hpp:6-11 */
    int32_t pad1;
} or_object;

/* camera = origin(3), direction(3), orientation(3); direction/orientation already normalised as
 * SceneData.cpp:143-145 does at load time. */
typedef struct or_scene {
    uint32_t width, height;
    double camera[9];
    uint32_t n_objects;
    uint32_t pad;
    const or_object* objects;
} or_scene;

enum {
    OR_RNG_REFERENCE = 0,  /* curand XORWOW, seed 123456, one stream per reference "CUDA thread" (22x22 cells),
                              draws in the reference's order: reproduces oracle/_ref bit for bit            */
    OR_RNG_COUNTER = 1     /* Philox4x32-7 keyed by (seed), counter = (pixel, sample, lane<<8|depth, tag):
                              the stream the B200 kernels use, so images can be compared pixel by pixel     */
};

typedef struct or_counts {
    uint64_t samples;
    uint64_t casts_reference;  /* nearest-hit queries the reference performs                              */
    uint64_t casts_needed;     /* ... minus those that provably contribute 0        */
} or_counts;

/* Renders rows [row_begin,row_end) (OR_RNG_COUNTER) or cells [cell_begin,cell_end) of the reference's 22x22
 * partition (OR_RNG_REFERENCE; use -1 for "all") into out_rgb (W*H*3 float64, row-major z*W+x, untouched pixels
 * keep their value).  Returns 0, or -1 on bad arguments. */
int or_render(const or_scene* scene, uint32_t samples, uint32_t max_depth, int rng_mode, uint64_t seed,
              int begin, int end, int nthreads, double* out_rgb, or_counts* counts);

/* Renderer::getHitObjectAndDistance (Renderer.cu:227-243) for n_rays rays (o,d as 6 doubles each). */
void or_nearest_hit(const or_scene* scene, const double* rays, uint32_t n_rays, int32_t* out_index, double* out_t);

/* SURVEY.md §8(c) Route 3 - the same two calls through the oracle's own median-split box tree, for scenes the scan cannot
 * finish (config 5's million objects).  Not in the reference; same object and bit-identical t as the scan (checked in
 * tests/test_oracle_pin.py), so frames are bit-identical to or_render's. */
int or_render_accel(const or_scene* scene, uint32_t samples, uint32_t max_depth, int rng_mode, uint64_t seed,
                    int begin, int end, int nthreads, double* out_rgb, or_counts* counts);
void or_nearest_hit_accel(const or_scene* scene, const double* rays, uint32_t n_rays, int32_t* out_index, double* out_t);

/* Function-level entry points mirroring oracle/ref_shim.cpp's ref_* ones. */
double or_sphere_intersect(double radius, const double* c, const double* o, const double* d);
double or_plane_intersect(const double* north, const double* east, const double* c, const double* o, const double* d);
void or_scatter(int kind, const double* geom, int reflection, const double* P, const double* incoming, int depth,
                unsigned long long subsequence, double* out16);

/* Philox4x32-R block function (counter[4], key[2]) -> out[4], rounds = 0: the 7 rounds the counter stream uses; and the two
 * 32-bit -> real maps of the counter stream. */
void or_philox4x32(const uint32_t* counter, const uint32_t* key, int rounds, uint32_t* out);
double or_sym24(uint32_t x);     /* (-1,1): odd multiple of 2^-24 */
double or_uniform23(uint32_t x); /* (0,1)  : odd multiple of 2^-24 */

/* Image.cpp:19-22 toRgb: clamp(int(x*255), 0, 255). */
int or_to_rgb(double x);

#ifdef __cplusplus
}
#endif
#endif
