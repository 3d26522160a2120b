// scene_pack.h -- host-side re-layout of an uploaded scene (internal).
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

#include "clpt_types.h"

struct ClptNode8 {
    uint32_t x, y;
};
struct ClptFloat4 {
    float x, y, z, w;
};

struct ClptPackedScene {
    std::vector<ClptNode8> nodes;
    std::vector<ClptFloat4> leaves; // 4 per leaf
    std::vector<ClptFloat4> tri;    // 3 per leaf triangle slot
    float root_min[3], root_max[3];
    int n_nodes = 0, n_leaves = 0, n_refs = 0, n_prims = 0;
    std::vector<int> lut; // start-node table, lut_dim[0] fastest
    int lut_dim[3] = { 1, 1, 1 };
    float lut_scale[3] = { 0, 0, 0 };
};

// Returns false and fills `err` when the input is inconsistent (index out of
// range, malformed node); the caller treats that as a fatal error.
bool clpt_pack_scene(const kdnode *nodes, size_t n_nodes, const int *tri_indices, size_t n_refs,
                     const cl_int3 *corners, size_t n_corners, const Vector4 *verts, size_t n_verts,
                     size_t n_norms, ClptPackedScene &out, std::string &err);
