// host_sanitize.cpp - the host layer (scene_loader, bvh, grid, output, cli) under AddressSanitizer + UBSan, CPU only:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -Iinclude tools/host_sanitize.cpp improved-path-tracer_b200/host/{scene_loader,bvh,grid,output,cli}.cpp -o /tmp/host_sanitize -lz -lpthread
//   /tmp/host_sanitize scene.json ...   (loads each file, builds the BVH + grid at leaf sizes 4, 1, 16, frees it)
// Round 2: clean (and clean under -fsanitize=thread) on the three shipped scenes, the 1M-primitive scene of config 5 and 3000 mutated copies of spheres.json / maze.json.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ipt_host.h"
int main(int argc, char** argv) {
    char msg[256];
    for (int i = 1; i < argc; i++) {
        ipt_host_scene* s = ipt_host_load_scene(argv[i], msg, sizeof msg);
        if (!s) { std::printf("%s: %s\n", argv[i], msg); continue; }
        int n = ipt_host_build_bvh(s, 4, 0);
        const ipt_scene* v = ipt_host_scene_view(s);
        std::printf("%s: nodes %d view %p\n", argv[i], n, (const void*)v);
        n = ipt_host_build_bvh(s, 1, 0);
        n = ipt_host_build_bvh(s, 16, 0);
        ipt_host_set_size(s, 33, 17);
        ipt_host_free_scene(s);
    }
    // the banded PNG encoder (several deflate streams on threads, one zlib stream out)
    const uint32_t W = 1280, H = 720;
    std::vector<float> img((size_t)W * H * 3);
    for (size_t i = 0; i < img.size(); i++) img[i] = (float)((i * 2654435761u >> 8) & 0xffff) / 50000.f - 0.1f;
    std::printf("png: %d\n", ipt_host_write_png("/tmp/host_sanitize.png", img.data(), W, H));
    std::printf("png 1x1: %d\n", ipt_host_write_png("/tmp/host_sanitize_1.png", img.data(), 1, 1));
    return 0;
}
