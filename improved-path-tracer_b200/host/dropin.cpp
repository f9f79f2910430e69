// dropin.cpp — ipt_render_objects(): the call a maintainer of the reference would make from
// RenderContoller::start() (RenderController.cu:36-70) instead of cudaMalloc/cudaMemcpy/cudaMain<<<22,22>>>:
// the reference's own ObjectData[] and Camera in, the W*H Vec3 (3 x fp64) frame out.  See INTEGRATION.md.
#include <vector>

#include "host_scene.hpp"

extern "C" int ipt_render_objects(const void* objects, uint32_t n_objects, uint32_t width, uint32_t height,
                                  const double* camera, uint32_t samples, uint32_t max_depth, int n_gpus, double* out_image)
{
    if (!objects || !camera || !out_image || n_objects == 0) return IPT_ERR_BAD_ARGUMENT;
    ipt_host_scene* s = ipt_host_from_objects(objects, n_objects, width, height, camera);
    if (!s) return IPT_ERR_BAD_ARGUMENT;
    ipt_host_build_bvh(s, IPT_DEFAULT_LEAF_SIZE, IPT_DEFAULT_BRUTE_MAX);
    ipt_params p = {};
    p.samples = samples; p.max_depth = max_depth;
    p.seed = 123456;   // the reference's curand seed (Renderer.cu:97)
    const int rc = ipt_render(ipt_host_scene_view(s), &p, n_gpus < 1 ? 1 : n_gpus, nullptr, out_image, nullptr);
    ipt_host_free_scene(s);
    return rc;
}
