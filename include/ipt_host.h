/* ipt_host.h — C interface of the C++ host layer that sits above the device ABI (include/ipt_abi.h):
 * scene ingest (the reference's scenes/<name>.json schema -> flattened structure-of-arrays primitive, material and BVH
 * buffers), output mapping (toRgb + PNG), the benchmark.txt record and the command-line grammar.
 * Each function names the reference code whose observable behaviour it keeps (paths relative to the
 * AdamStudies-PWR/Improved-Path-Tracer tree).  These are host one-offs, not the accelerated path. */
#ifndef IPT_HOST_H
#define IPT_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "ipt_abi.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipt_host_scene ipt_host_scene;   /* owns the arrays an ipt_scene points to */

/* SceneData::initScene (SceneData.cpp:61-96) without the printing: parses `path` and flattens it.  On failure
 * returns NULL and writes the message the reference would print ("Could not load provided json file!",
 * "Missing height or witdh data!", "No camera data!", ... — SURVEY.md App. C) to `message`. */
ipt_host_scene* ipt_host_load_scene(const char* path, char* message, size_t message_len);

/* The same from the reference's in-memory form: n ObjectData records (144-byte layout, ObjectData.hpp:15-31) and the
 * Camera (origin, direction, orientation: 9 doubles; direction/orientation are used as given). */
ipt_host_scene* ipt_host_from_objects(const void* objects, uint32_t n_objects, uint32_t width, uint32_t height,
                                      const double* camera9);

void ipt_host_free_scene(ipt_host_scene* scene);
const ipt_scene* ipt_host_scene_view(const ipt_host_scene* scene);
void ipt_host_set_size(ipt_host_scene* scene, uint32_t width, uint32_t height);   /* "spheres.json at 3840x2160" */

/* Builds (or drops) the BVH.  leaf_size 1..16 primitives per leaf; brute_max: scenes with at most this many
 * primitives get no BVH (every ray tests every primitive from shared memory).  Returns node count or <0.
 * Scenes that get a BVH and consist of small, evenly spread primitives also get the uniform grid of ipt_scene::grid_*
 * (host/grid.cpp: which scenes qualify; IPT_NO_GRID=1 in the environment switches it off). */
int ipt_host_build_bvh(ipt_host_scene* scene, uint32_t leaf_size, uint32_t brute_max);
#define IPT_DEFAULT_LEAF_SIZE 4
#define IPT_DEFAULT_BRUTE_MAX 192   /* measured crossover brute force vs BVH pipeline on B200: ~200 primitives */

/* Image.cpp:19-22: clamp(int(x*255), 0, 255) — truncation, no gamma. */
int ipt_host_to_rgb(double x);
/* Image.cpp:39-56: 8-bit RGB PNG, row 0 on top, written to `path` (zlib deflate; Magick++ is not needed). */
int ipt_host_write_png(const char* path, const float* rgb, uint32_t width, uint32_t height);
/* The same from bytes already mapped by toRgb (ipt_render_rgb8 / ipt_ctx_download_rgb8). */
int ipt_host_write_png_rgb8(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height);

/* Measurements.cpp:26-41: "HH:MM:SS.ms" with the milliseconds NOT zero-padded. */
void ipt_host_time_string(uint64_t milliseconds, char* out, size_t out_len);
/* Measurements.cpp:43-55: appends "<id>;<time>;" (no newline) to `file` (benchmark.txt). */
int ipt_host_append_benchmark(const char* file, const char* id, const char* time_string);

/* InputParser (InputParser.cpp:72-258): argv grammar, ranges, messages printed to stdout.
 * argc/argv as main() receives them.  Returns 1 if the input is valid, 0 otherwise (the reference then exits 0). */
typedef struct ipt_cli {
    char scene_path[4096];
    char scene_name[1024];     /* basename without the last extension (InputParser.cpp:41-55) */
    uint16_t samples;          /* default 40, 4..65535 */
    uint8_t max_depth;         /* default 10, 3..255   */
} ipt_cli;
int ipt_host_parse_cli(int argc, char** argv, ipt_cli* out);

#ifdef __cplusplus
}
#endif
#endif
