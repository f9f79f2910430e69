// host_scene.hpp — the owning form of a flattened scene (see include/ipt_abi.h: ipt_scene for the layout).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/ipt_host.h"

struct ipt_host_scene {
    ipt_scene view = {};
    std::vector<double> sphere_cxyzr, rect_plane, rect_u, rect_v, rect_bounds, mat_color, mat_emission;
    std::vector<uint32_t> sphere_object, rect_object;
    std::vector<int32_t> mat_reflection;
    // kept for the BVH builder (bounding boxes need the corners: centre +- north +- east, Plane.cu:38-41)
    std::vector<double> rect_center, rect_north, rect_east;
    std::vector<ipt_bvh_node> bvh_nodes;
    std::vector<uint32_t> bvh_slot_prim;
    // uniform grid over the BVH's slots (host/grid.cpp), empty when the scene does not qualify
    std::vector<uint32_t> grid_cell_start, grid_refs, grid_big;
    uint32_t grid_res[3] = {0, 0, 0};
    float grid_lo[3] = {0, 0, 0}, grid_cell[3] = {0, 0, 0};
    void build_grid();   // host/grid.cpp; called by ipt_host_build_bvh

    void add(int type, double radius, const double* north, const double* east, const double* position,
             const double* emission, const double* color, int reflection);
    void refresh_view();
};
