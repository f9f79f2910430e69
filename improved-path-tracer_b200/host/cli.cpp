// cli.cpp — the reference's command-line grammar (InputParser.cpp:72-258), message for message.
#include <algorithm>
#include <cstdio>
#include <filesystem>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ipt_host.h"

namespace {
const int MIN_DEPTH = 3, MAX_DEPTH = 255, MIN_SAMPLES = 4, MAX_SAMPLES = 65535;   // InputParser.cpp:14-21

void help()   // InputParser.cpp:249-258
{
    std::cout << "tracer [arguments] [path_to_scene]" << std::endl;
    std::cout << "[arguments] are [-s/--samples] or [-d/--depth]" << std::endl;
    std::cout << "\t [OPTIONAL] -s=number or --samples=number - Specifies number of samples per pixel. "
              << "It must be between " << MIN_SAMPLES << " and " << MAX_SAMPLES << std::endl;
    std::cout << "\t [OPTIONAL] -d=number or --depth=number - Specifies max number of reflections per ray. "
              << "It must be between " << MIN_DEPTH << " and " << MAX_DEPTH << std::endl;
    std::cout << "[path_to_scene] - Specifies path to json file with scene data. It is mandatory." << std::endl;
}

bool error(const std::string& cause)   // InputParser.cpp:241-247
{
    std::cout << "Error parsing input!" << std::endl;
    std::cout << "Cause: " << cause << std::endl;
    std::cout << "Usage:" << std::endl;
    help();
    return false;
}

std::vector<std::string> split(std::string in, const std::string& sep)   // InputParser.cpp:26-39
{
    std::vector<std::string> out;
    size_t pos;
    while ((pos = in.find(sep)) != std::string::npos) { out.push_back(in.substr(0, pos)); in.erase(0, pos + sep.size()); }
    out.push_back(in);
    return out;
}

std::string scene_name(std::string s)   // InputParser.cpp:41-55
{
    const auto slash = s.rfind('/');
    if (slash != std::string::npos) s.erase(0, slash + 1);
    const auto dot = s.rfind('.');
    if (dot != std::string::npos) s.erase(dot);
    return s;
}

bool number(const std::string& text, int lo, int hi, const char* range_msg, const char* convert_msg, int& out)   // :185-239
{
    int v;
    try { v = std::stoi(text); }
    catch (const std::out_of_range&) { return error(range_msg); }
    catch (const std::invalid_argument&) { return error(convert_msg); }
    if (v < lo || v > hi) return error(range_msg);
    out = v;
    return true;
}

bool path_ok(const std::string& path, ipt_cli* out)   // InputParser.cpp:113-129
{
    std::error_code ec;
    if (!std::filesystem::exists(path, ec)) return error("Path does not exist");
    if (!std::filesystem::is_regular_file(path, ec)) return error("Not a file");
    std::snprintf(out->scene_path, sizeof(out->scene_path), "%s", path.c_str());
    std::snprintf(out->scene_name, sizeof(out->scene_name), "%s", scene_name(path).c_str());
    return true;
}
}  // namespace

extern "C" int ipt_host_parse_cli(int argc, char** argv, ipt_cli* out)
{
    if (!out) return 0;
    out->scene_path[0] = out->scene_name[0] = '\0';
    out->samples = 40; out->max_depth = 10;                                   // InputParser.cpp:16,19
    const int n = argc - 1;                                                   // main.cu:29
    if (n < 1 || n > 3) {                                                     // InputParser.cpp:74-81
        std::stringstream m;
        m << "Got " << n << " arguments! Expected between " << 1 << " and " << 3 << " arguments";
        return error(m.str());
    }
    if (n == 1) {                                                             // :85-89
        if (std::string(argv[1]) == "--help") { help(); return 0; }
        return path_ok(argv[1], out) ? 1 : 0;
    }
    if (!path_ok(argv[n], out)) return 0;                                     // :93 — the path is the LAST argument
    for (int i = 1; i < n; i++) {                                             // :131-183
        std::string arg = argv[i];
        const auto dashes = std::count(arg.begin(), arg.end(), '-');
        if (dashes != 1 && dashes != 2) return error("Arguments can have 1 or 2 (-)! Please check your input");
        arg.erase(std::remove(arg.begin(), arg.end(), '-'), arg.end());
        const auto kv = split(arg, "=");
        if (kv.size() != 2) return error("Cannot parse argument: " + arg);
        if (dashes == 1) { if (kv[0] != "s" && kv[0] != "d") return error("Unknown short argument: " + arg); }
        else if (kv[0] != "samples" && kv[0] != "depth") return error("Unknown long argument: " + arg);
        int v;
        if (kv[0] == "s" || kv[0] == "samples") {
            if (!number(kv[1], MIN_SAMPLES, MAX_SAMPLES, "Number of samples out of range!", "Could not convert samples to number!", v)) return 0;
            out->samples = (uint16_t)v;
        }
        if (kv[0] == "d" || kv[0] == "depth") {
            if (!number(kv[1], MIN_DEPTH, MAX_DEPTH, "Depth out of range!", "Could not convert depth to number!", v)) return 0;
            out->max_depth = (uint8_t)v;
        }
    }
    return 1;
}
