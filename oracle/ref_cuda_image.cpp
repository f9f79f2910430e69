// TEST / BASELINE INFRASTRUCTURE ONLY.
// Stand-in for the reference's src/utils/Image.cpp when its CUDA program is built on a box without Magick++
// (oracle/Makefile target ref_cuda).  Same entry point (Image.hpp:10-12); instead of a PNG it writes the raw frame
// "<filename>.f64": uint32 width, uint32 height, then width*height*3 float64 (row-major, as produced by the kernel),
// so that the reference GPU build's image can be compared with the host oracle channel by channel.
// The reference keeps the 8-bit frame in a stack VLA (Image.cpp:49); nothing of that kind here, so 4K frames work.
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "utils/Image.hpp"

namespace tracer::utils
{
void saveImage(const std::vector<containers::Vec3>& image, const uint32_t height, const uint32_t width, const std::string filename)
{
    std::cout << "Saving Image..." << std::endl;
    if (width * height != image.size())
    {
        std::cout << "Error saving image! Size missmatch!" << std::endl;
        return;
    }
    std::FILE* f = std::fopen((filename + ".f64").c_str(), "wb");
    if (!f) return;
    std::fwrite(&width, 4, 1, f);
    std::fwrite(&height, 4, 1, f);
    for (const auto& p : image)
    {
        const double v[3] = {p.xx_, p.yy_, p.zz_};
        std::fwrite(v, 8, 3, f);
    }
    std::fclose(f);
}
}  // namespace tracer::utils
