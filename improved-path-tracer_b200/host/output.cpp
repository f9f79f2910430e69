// output.cpp — what happens right after the hot path in the reference: Image.cpp (toRgb + PNG) and
// Measurements.cpp (time string + benchmark.txt).  Observable behaviour kept; implementation is new
// (Magick++ is replaced by a direct zlib PNG encoder).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

#include "../../include/ipt_host.h"

// Image.cpp:19-22: std::clamp(int(x * 255), 0, 255) — truncation toward zero, no gamma.  int(NaN) is undefined
// upstream; here NaN maps to 0 and values beyond the int range saturate.
extern "C" int ipt_host_to_rgb(double x)
{
    const double v = x * 255;
    if (!(v == v)) return 0;
    if (v >= 2147483647.0) return 255;
    if (v <= -2147483648.0) return 0;
    return std::clamp((int)v, 0, 255);
}

namespace {
void put32(std::vector<unsigned char>& o, uint32_t v) { o.push_back(v >> 24); o.push_back(v >> 16); o.push_back(v >> 8); o.push_back(v); }
void chunk(std::vector<unsigned char>& out, const char* type, const unsigned char* data, size_t n)
{
    put32(out, (uint32_t)n);
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4)));
}
}  // namespace

// Image.cpp:39-56: W x H, 8-bit RGB, row 0 = top.  (The reference keeps the bytes in a stack VLA, which overflows
// at 3840x2160; a heap buffer is used here.)
extern "C" int ipt_host_write_png_rgb8(const char* path, const uint8_t* rgb8, uint32_t W, uint32_t H)
{
    if (!path || !rgb8 || !W || !H) return -1;
    std::vector<unsigned char> raw((size_t)H * (1 + (size_t)W * 3));
    for (uint32_t z = 0; z < H; z++) {
        unsigned char* row = &raw[(size_t)z * (1 + (size_t)W * 3)];
        row[0] = 0;   // filter: none
        std::memcpy(row + 1, rgb8 + (size_t)z * W * 3, (size_t)W * 3);
    }
    uLongf zn = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(zn);
    if (compress2(z.data(), &zn, raw.data(), (uLong)raw.size(), 6) != Z_OK) return -1;
    std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<unsigned char> ihdr;
    put32(ihdr, W); put32(ihdr, H);
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});   // 8 bits, colour type 2 (RGB), deflate, no filter method, no interlace
    chunk(out, "IHDR", ihdr.data(), ihdr.size());
    chunk(out, "IDAT", z.data(), zn);
    chunk(out, "IEND", nullptr, 0);
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return -1;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok ? 0 : -1;
}

extern "C" int ipt_host_write_png(const char* path, const float* rgb, uint32_t W, uint32_t H)
{
    if (!path || !rgb || !W || !H) return -1;
    std::vector<uint8_t> bytes((size_t)W * H * 3);
    for (size_t i = 0; i < bytes.size(); i++) bytes[i] = (uint8_t)ipt_host_to_rgb((double)rgb[i]);
    return ipt_host_write_png_rgb8(path, bytes.data(), W, H);
}

// Measurements.cpp:21-41: every unit is "00" when zero, zero-padded to two digits below 10; the milliseconds are
// printed as a plain integer (5007 ms -> "00:00:05.7").
extern "C" void ipt_host_time_string(uint64_t ms, char* out, size_t n)
{
    auto unit = [](uint64_t v) { return v == 0 ? std::string("00") : v < 10 ? "0" + std::to_string(v) : std::to_string(v); };
    const uint64_t h = ms / 3600000; ms -= 3600000 * h;
    const uint64_t m = ms / 60000; ms -= 60000 * m;
    const uint64_t s = ms / 1000; ms -= 1000 * s;
    const std::string t = unit(h) + ":" + unit(m) + ":" + unit(s) + "." + std::to_string(ms);
    if (out && n) std::snprintf(out, n, "%s", t.c_str());
}

// Measurements.cpp:43-55: append "<id>;<time>;" — no newline (test_automation.py adds "<cpuMiB>;<gpuMiB>\n").
extern "C" int ipt_host_append_benchmark(const char* file, const char* id, const char* time_string)
{
    std::FILE* f = std::fopen(file ? file : "benchmark.txt", "ab");
    if (!f) return -1;
    std::fprintf(f, "%s;%s;", id, time_string);
    std::fclose(f);
    return 0;
}
