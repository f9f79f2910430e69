// tracer — same command line, same stdout lines, same outputs (sceneNameDXSY.png, benchmark.txt record) and the same
// exit codes as the reference program (src/main.cu:22-61); the render itself goes through the C ABI of libipt_b200.
//   tracer [-d=N|--depth=N] [-s=N|--samples=N] scene.json
// Environment (additions, all optional): IPT_GPUS=1|2|4|8 devices to tile the image over (default 1, as the reference
// uses device 0 only); IPT_SEED=n; IPT_FP64=1 parity mode; IPT_WIDTH / IPT_HEIGHT override the scene's frame size.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/ipt_host.h"

int main(int argc, char* argv[])
{
    if (ipt_device_count() <= 0) {                                        // CudaUtils.cu:13-17
        std::cout << "CUDA capable device not found! Cannot continue";
        return 0;
    }
    std::cout << "Using GPU device: " << ipt_device_name(0) << std::endl;   // CudaUtils.cu:19-21

    ipt_cli cli;
    if (!ipt_host_parse_cli(argc, argv, &cli)) return 0;                    // main.cu:29-33

    std::cout << "Loading Scene Data..." << std::endl;                      // SceneData.cpp:63
    char message[256];
    ipt_host_scene* scene = ipt_host_load_scene(cli.scene_path, message, sizeof(message));
    if (!scene) {
        std::cout << message << std::endl;
        return 0;                                                           // main.cu:35-39
    }
    if (const char* w = std::getenv("IPT_WIDTH")) if (const char* h = std::getenv("IPT_HEIGHT")) ipt_host_set_size(scene, (uint32_t)std::atoi(w), (uint32_t)std::atoi(h));
    ipt_host_build_bvh(scene, IPT_DEFAULT_LEAF_SIZE, IPT_DEFAULT_BRUTE_MAX);
    std::cout << "Data loaded successfully" << std::endl;                   // SceneData.cpp:93
    const ipt_scene* view = ipt_host_scene_view(scene);

    const std::string id = std::string(cli.scene_name) + "D" + std::to_string(+cli.max_depth) + "S" + std::to_string(+cli.samples);   // main.cu:41-43

    ipt_params params = {};
    params.samples = cli.samples;
    params.max_depth = cli.max_depth;
    params.seed = std::getenv("IPT_SEED") ? std::strtoull(std::getenv("IPT_SEED"), nullptr, 10) : 123456ull;
    if (std::getenv("IPT_FP64") && std::atoi(std::getenv("IPT_FP64"))) params.flags |= IPT_FLAG_FP64;
    // opt-in variance reduction (none of it is in the reference; the CLI grammar stays the reference's)
    if (std::getenv("IPT_ROULETTE") && std::atoi(std::getenv("IPT_ROULETTE"))) params.flags |= IPT_FLAG_RUSSIAN_ROULETTE;
    if (std::getenv("IPT_STRATIFIED") && std::atoi(std::getenv("IPT_STRATIFIED"))) params.flags |= IPT_FLAG_STRATIFIED;
    if (std::getenv("IPT_NEXT_EVENT") && std::atoi(std::getenv("IPT_NEXT_EVENT"))) params.flags |= IPT_FLAG_NEXT_EVENT;
    int gpus = std::getenv("IPT_GPUS") ? std::atoi(std::getenv("IPT_GPUS")) : 1;
    if (gpus < 1) gpus = 1;

    // Like the reference, whose first context-creating call (cudaMalloc, RenderController.cu:43) sits inside measure(): the
    // CUDA context is created inside the timed call below (IPT_VERBOSE=1 prints how long that took, "[ipt] contexts").
    std::vector<uint8_t> image((size_t)view->width * view->height * 3);   // toRgb runs on the device: bytes come back
    ipt_stats stats = {};
    // Measurements.cpp:58-70: the timed region is the whole render call (allocation, upload, kernels, copy back)
    std::cout << "Begining render..." << std::endl;
    std::printf("\rRendering %.2f%%", 0.0f);
    std::fflush(stdout);
    // Renderer.cu:105-107: progress while the frame renders (here: share of the wavefront batches the device has finished)
    ipt_set_progress([](double done, void*) { std::printf("\rRendering %.2f%%", (float)(done * 100.0)); std::fflush(stdout); }, nullptr);
    const auto t0 = std::chrono::high_resolution_clock::now();
    const int rc = ipt_render_rgb8(view, &params, gpus, image.data(), &stats);
    const auto t1 = std::chrono::high_resolution_clock::now();
    ipt_set_progress(nullptr, nullptr);
    if (rc == IPT_OK) std::printf("\rRendering %.2f%%", 100.0f);
    else std::cout << "render error: " << ipt_last_error() << std::endl;    // RenderController.cu:20-27 prints and carries on
    std::cout << " - Done" << std::endl;
    char time[64];
    ipt_host_time_string((uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count(), time, sizeof(time));
    std::cout << "Render took: " << time << std::endl;
    ipt_host_append_benchmark("benchmark.txt", id.c_str(), time);
    if (rc != IPT_OK) { ipt_host_free_scene(scene); return 1; }             // main.cu:52-56

    std::cout << "Saving Image..." << std::endl;                            // Image.cpp:41
    ipt_host_write_png_rgb8((id + ".png").c_str(), image.data(), view->width, view->height);
    if (std::getenv("IPT_VERBOSE"))
        std::fprintf(stderr, "[ipt] %.3f Msamples/s  %.3f Gbounces/s  kernels %.1f ms  launches %llu\n",
                     stats.samples / stats.render_ms * 1e-3, stats.traced_bounces / stats.render_ms * 1e-6, stats.render_ms,
                     (unsigned long long)stats.kernel_launches);
    ipt_host_free_scene(scene);
    return 0;
}
