// TEST INFRASTRUCTURE ONLY - not part of the product path.
//
// The reference's own command-line parser (/root/reference/src/utils/InputParser.cpp, UNMODIFIED, compiled where it lies
// by oracle/Makefile into oracle/_ref/libref_cli.so) behind a C entry point, so that the host layer's ipt_host_parse_cli
// can be compared with it argument list by argument list: validity, parsed values and every line printed
// (tests/test_host.py::test_cli_matches_the_reference_parser).  main.cu:29 passes argc - 1; so does this.
// Likewise Measurements.cpp: getTimeString and saveBenchmark (Measurements.cpp:26-55) for ipt_host_time_string and
// ipt_host_append_benchmark (test_time_string_and_benchmark_record_match_the_reference).
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "utils/InputParser.hpp"
#include "Measurements.cpp"   // the reference file, unmodified (-I$(REF)/src/utils): its time format and benchmark.txt record live in
                              // an anonymous namespace, reachable only from the same translation unit

extern "C" int ref_parse_cli(int argc, char** argv, char* scene_path, int path_len, char* scene_name, int name_len,
                             int* samples, int* max_depth)
{
    tracer::utils::InputParser parser(argc - 1, argv);
    std::cout.flush();
    std::snprintf(scene_path, (size_t)path_len, "%s", parser.getScenePath().c_str());
    std::snprintf(scene_name, (size_t)name_len, "%s", parser.getSceneName().c_str());
    *samples = parser.getSamplingRate();
    *max_depth = parser.getMaxDepth();
    return parser.isInputValid() ? 1 : 0;
}

extern "C" void ref_time_string(unsigned long long milliseconds, char* out, int out_len)
{
    std::snprintf(out, (size_t)out_len, "%s", tracer::utils::getTimeString(milliseconds).c_str());
}

// Appends to benchmark.txt in the current directory, as the reference does.
extern "C" void ref_save_benchmark(const char* id, const char* time) { tracer::utils::saveBenchmark(id, time); }
