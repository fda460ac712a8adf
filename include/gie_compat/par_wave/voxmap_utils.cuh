// Stand-in for include/par_wave/voxmap_utils.cuh: the voxel-block layout and key helpers that the reference documents for
// planners reading the streamed global map (README.md:163-170).  get_VB_key / get_voxID_in_VB are __host__ __device__ when
// compiled by nvcc, as in the reference; the device-side access to the engine's pools is in gie_device_view.cuh.
#pragma once
#ifdef __CUDACC__
#define GIE_HD __host__ __device__ __forceinline__
#else
#define GIE_HD inline
#endif
#include <cstddef>
#include <cstdint>
#include "map_structure/local_batch.h"

#define EMPTY_VALUE GIE_EMPTY_VALUE
#define EMPTY_KEY (int3{EMPTY_VALUE, EMPTY_VALUE, EMPTY_VALUE})
#define VB_WIDTH 8
#define VB_SIZE 512

typedef union {
    int sq_dist[2];         // [0] = squared distance
    int parent_loc_id[2];   // [1] = wave-range coc id (11/11/10 bit)
    unsigned long long int ulong;
} Dist_id;

struct GlbVoxel {           // 40 bytes, same offsets as the reference's GlbVoxel and as gie_glbvoxel
    unsigned char occ_val = 0;
    char vox_type = VOXTYPE_UNKNOWN;
    int update_ct = 0;
    int3 coc_glb = EMPTY_KEY;
    int dist_sq = EMPTY_VALUE;
    int wave_layer = -1;
    Dist_id dist_id_pair;
};
static_assert(sizeof(GlbVoxel) == sizeof(gie_glbvoxel) && offsetof(GlbVoxel, dist_id_pair) == offsetof(gie_glbvoxel, dist_id_pair) &&
              offsetof(GlbVoxel, coc_glb) == offsetof(gie_glbvoxel, coc_glb), "GlbVoxel must match the C ABI layout");
struct VoxelBlock { GlbVoxel voxels[VB_SIZE]; };

struct CrdEqualTo { bool operator()(int3 a, int3 b) const { return a.x == b.x && a.y == b.y && a.z == b.z; } };
struct CrdLessThan {
    bool operator()(int3 a, int3 b) const { return a.x != b.x ? a.x < b.x : (a.y != b.y ? a.y < b.y : a.z < b.z); }
};
struct BlockHasher {
    size_t operator()(int3 k) const { return ((size_t)k.x * 73856093u) ^ ((size_t)k.y * 19349669u) ^ ((size_t)k.z * 83492791u); }
};

// block key of a global voxel coordinate: floor(c / 8) per axis
GIE_HD int3 get_VB_key(const int3 &c) { return make_int3(c.x >> 3, c.y >> 3, c.z >> 3); }
// voxel index inside its block, reference order (x slowest, z fastest)
GIE_HD int get_voxID_in_VB(const int3 &c) { return (c.x & 7) * 64 + (c.y & 7) * 8 + (c.z & 7); }
GIE_HD int3 reconstruct_vox_crd(const int3 &blk_offset, const int &idx)
{
    return make_int3(blk_offset.x + ((idx >> 6) & 7), blk_offset.y + ((idx >> 3) & 7), blk_offset.z + (idx & 7));
}
// voxmap_utils.cuh:161-172
GIE_HD bool invalid_dist_glb(const int dist) { return dist < 0 || dist >= 900000; }
GIE_HD bool invalid_coc_glb(const int3 coc) { return coc.x > 900000 || coc.y > 900000 || coc.z > 900000; }
GIE_HD bool invalid_blk_key(const int3 &k) { return k.x >= EMPTY_VALUE || k.y >= EMPTY_VALUE || k.z >= EMPTY_VALUE; }
