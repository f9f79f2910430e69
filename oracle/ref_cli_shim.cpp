// TEST INFRASTRUCTURE ONLY - not part of the product path.
//
// The reference's own command-line parser (/root/reference/src/utils/InputParser.cpp, UNMODIFIED, compiled where it lies
// by oracle/Makefile into oracle/_ref/libref_cli.so) behind a C entry point, so that the host layer's ipt_host_parse_cli
// can be compared with it argument list by argument list: validity, parsed values and every line printed
// (tests/test_host.py::test_cli_matches_the_reference_parser).  main.cu:29 passes argc - 1; so does this.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "utils/InputParser.hpp"

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
