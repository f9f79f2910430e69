// TEST INFRASTRUCTURE ONLY - the reference's RenderController.cu with start() replaced by the body INTEGRATION.md
// section 1 shows, compiled against the reference's own headers (oracle/Makefile target ref_dropin): the reference's
// main.cu, InputParser, SceneData and Measurements, unmodified, drive libipt_b200.so through ipt_render_objects().
// tests/test_gpu_parity.py::test_reference_program_with_dropin_controller runs the program and compares its frame
// with the one `tracer` writes for the same seed.
#include "renderer/RenderController.hpp"

#include <iostream>

#include "ipt_abi.h"

namespace tracer::renderer
{

RenderContoller::RenderContoller(scene::SceneData& sceneData, const uint32_t samples, const uint8_t maxDepth)
    : sceneData_(sceneData)
    , maxDepth_(maxDepth)
    , samples_(samples)
{}

// ---- INTEGRATION.md section 1, verbatim -------------------------------------------------------------------------
std::vector<containers::Vec3> RenderContoller::start()
{
    static_assert(sizeof(scene::objects::ObjectData) == 144 && sizeof(scene::objects::Camera) == 72);
    const auto objects = sceneData_.getObjectsData();          // std::vector<ObjectData>, JSON order
    const auto camera  = sceneData_.getCamera();               // direction/orientation already normalised (SceneData.cpp:143-145)
    const uint32_t W = sceneData_.getWidth(), H = sceneData_.getHeight();
    std::vector<containers::Vec3> image(size_t(W) * H);        // Vec3 = 3 doubles, contiguous
    const int rc = ipt_render_objects(objects.data(), uint32_t(objects.size()), W, H,
                                      reinterpret_cast<const double*>(&camera), samples_, maxDepth_,
                                      /*n_gpus=*/1, reinterpret_cast<double*>(image.data()));
    if (rc != IPT_OK) std::cout << "cudaMain kernel error: " << ipt_last_error() << std::endl;   // as cudaErrorCheck() does
    return image;
}
// -----------------------------------------------------------------------------------------------------------------

std::vector<containers::Vec3> RenderContoller::convertToVector(containers::Vec3*) { return {}; }   // unused by the new start()

}  // namespace tracer::renderer
