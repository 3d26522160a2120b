// kd_build_gpu.h -- device-side kd-tree construction and re-layout (internal).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

struct ClptGpuBuildParams {
    int max_depth = 0;  // <= 0: 8 + 1.3 log2(triangles)
    int min_split = 2;  // nodes with fewer references become leaves without looking for a plane
    float ct = 1.0f, ci = 1.0f, empty_bonus = 0.9f; // the surface-area heuristic's constants (as build_kd_sah)
};

// The tree in the reference's wire format, in device memory (owned by the caller's
// ClptGpuTree, reused and grown across builds).
struct ClptGpuTree {
    int *wire = nullptr;        // 17 words per node (include/kd_tree.h:31-50)
    size_t wire_cap = 0;        // in words
    int *tri_indices = nullptr; // leaf triangle lists
    size_t tri_indices_cap = 0;
    int n_nodes = 0, n_refs = 0, levels = 0;
};

// Builds the tree of `n_tris` triangles (corners: three int4 {v, vn, vt, 0} per triangle,
// verts: float4) on stream `s`.  Meshes of up to 2^18 triangles: the whole build is one recorded
// CUDA graph (level counts stay on the device, launches sized by fixed capacities), replayed
// while mesh size and buffers stay the same, one synchronisation at the end.  Larger meshes,
// or a mesh that outgrows the recorded capacities: level by level, one synchronisation per
// level (two counters come back to size the next level exactly).  Both give the same tree.
// False + message on malformed input.
bool clpt_gpu_build(const float4 *verts, int n_verts, const int4 *corners, int n_tris, const ClptGpuBuildParams &P,
                    ClptGpuTree &out, cudaStream_t s, char *err, size_t errlen);
bool clpt_gpu_build_was_recorded(void); // the last build ran as the recorded graph
void clpt_gpu_build_release(void);

// Device twin of clpt_pack_scene (scene_pack.cpp): wire format -> traversal layout
// (clpt_device.cuh), same numbering, same bytes.  Buffers are (re)allocated by the callee
// through `alloc` when too small.
struct ClptGpuPacked {
    uint2 *nodes = nullptr;
    float4 *leaves = nullptr, *tri = nullptr, *flat_n = nullptr;
    int *lut = nullptr;
    size_t nodes_cap = 0, leaves_cap = 0, tri_cap = 0, lut_cap = 0, flat_n_cap = 0;
    int n_nodes = 0, n_leaves = 0, n_refs = 0;
    float root_min[3], root_max[3];
    int lut_dim[3];
    float lut_scale[3];
    size_t fat_refs = 0; // triangle slots living in leaves of >= CLPT_COOP_LEAF_MIN triangles
};
bool clpt_gpu_pack(const ClptGpuTree &tree, const float4 *verts, const int4 *corners, int n_prims, ClptGpuPacked &out,
                   cudaStream_t s, char *err, size_t errlen);
