// output.cpp — what happens right after the hot path in the reference: Image.cpp (toRgb + PNG) and
// Measurements.cpp (time string + benchmark.txt).  Observable behaviour kept; implementation is new
// (Magick++ is replaced by a direct zlib PNG encoder).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
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
unsigned host_threads()   // IPT_HOST_THREADS as in grid.cpp
{
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if (const char* e = std::getenv("IPT_HOST_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    return std::min(nt, 64u);
}
template <class F> void parallel_for(size_t n, F f)   // f(begin, end) on contiguous parts of [0, n)
{
    const size_t nt = std::min<size_t>(host_threads(), n);
    if (nt <= 1) { f((size_t)0, n); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) th.emplace_back([=] { f(n * t / nt, n * (t + 1) / nt); });
    for (auto& t : th) t.join();
}
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
// The zlib stream is deflated in bands of rows on several threads (a 4K frame took 0.67 s on one thread, longer than its
// render): every band is a raw deflate stream of its own that ends on a byte boundary (Z_SYNC_FLUSH; Z_FINISH for the
// last), the bands are concatenated behind one zlib header and the Adler-32 of the whole is combined from the bands'.
static int write_png_rgb8(const char* path, const uint8_t* rgb8, uint32_t W, uint32_t H);
extern "C" int ipt_host_write_png_rgb8(const char* path, const uint8_t* rgb8, uint32_t W, uint32_t H)
{
    if (!path || !rgb8 || !W || !H || (uint64_t)W * H > (1ull << 29)) return -1;   // one IDAT chunk: below 2 GB of deflate output
    try { return write_png_rgb8(path, rgb8, W, H); }
    catch (...) { return -1; }                       // out of memory: nothing is thrown across the C ABI
}

static int write_png_rgb8(const char* path, const uint8_t* rgb8, uint32_t W, uint32_t H)
{
    const size_t stride = 1 + (size_t)W * 3;
    const unsigned nt = host_threads();
    // bands of at least ~256 KB so that the lost history between bands costs nothing measurable
    const uint32_t min_rows = (uint32_t)std::max<size_t>(1, (256u << 10) / stride);
    const uint32_t n_bands = std::max(1u, std::min<uint32_t>(nt * 4, H / min_rows));
    struct Band { std::vector<unsigned char> z; uLong adler = 1; size_t raw_len = 0; bool ok = false; };
    std::vector<Band> bands(n_bands);
    parallel_for(n_bands, [&](size_t b0, size_t b1) {
        for (size_t b = b0; b < b1; b++) {
            const uint32_t r0 = (uint32_t)((uint64_t)H * b / n_bands), r1 = (uint32_t)((uint64_t)H * (b + 1) / n_bands);
            Band& B = bands[b];
            std::vector<unsigned char> raw((size_t)(r1 - r0) * stride);
            for (uint32_t z = r0; z < r1; z++) {
                unsigned char* row = &raw[(size_t)(z - r0) * stride];
                row[0] = 0;   // filter: none
                std::memcpy(row + 1, rgb8 + (size_t)z * W * 3, (size_t)W * 3);
            }
            B.raw_len = raw.size();
            B.adler = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size());
            z_stream zs;
            std::memset(&zs, 0, sizeof(zs));
            if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) continue;
            B.z.resize(deflateBound(&zs, (uLong)raw.size()) + 16);
            zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size();
            zs.next_out = B.z.data(); zs.avail_out = (uInt)B.z.size();
            const bool last = b + 1 == n_bands;
            const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
            B.ok = last ? rc == Z_STREAM_END : (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0);
            B.z.resize(B.z.size() - zs.avail_out);
            deflateEnd(&zs);
        }
    });
    std::vector<unsigned char> z = {0x78, 0x9C};   // deflate, 32 KB window, default level, no dictionary
    uLong adler = 1;
    for (uint32_t b = 0; b < n_bands; b++) {
        if (!bands[b].ok) return -1;
        z.insert(z.end(), bands[b].z.begin(), bands[b].z.end());
        adler = b == 0 ? bands[b].adler : adler32_combine(adler, bands[b].adler, (z_off_t)bands[b].raw_len);
    }
    put32(z, (uint32_t)adler);
    if (z.size() > 0x7fffffffu) return -1;   // one IDAT chunk
    std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<unsigned char> ihdr;
    put32(ihdr, W); put32(ihdr, H);
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});   // 8 bits, colour type 2 (RGB), deflate, no filter method, no interlace
    chunk(out, "IHDR", ihdr.data(), ihdr.size());
    chunk(out, "IDAT", z.data(), z.size());
    chunk(out, "IEND", nullptr, 0);
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return -1;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok ? 0 : -1;
}

extern "C" int ipt_host_write_png(const char* path, const float* rgb, uint32_t W, uint32_t H)
{
    if (!path || !rgb || !W || !H || (uint64_t)W * H > (1ull << 29)) return -1;
    try {
        std::vector<uint8_t> bytes((size_t)W * H * 3);
        parallel_for(bytes.size(), [&](size_t a, size_t b) { for (size_t i = a; i < b; i++) bytes[i] = (uint8_t)ipt_host_to_rgb((double)rgb[i]); });
        return ipt_host_write_png_rgb8(path, bytes.data(), W, H);
    } catch (...) { return -1; }
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
