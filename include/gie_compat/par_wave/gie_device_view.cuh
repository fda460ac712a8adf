// Device-side access to the global map for GPU planners built together with the mapper (README.md:163-170).  In the
// reference a planner calls hash_table_D->get_alloc_blk_id(get_VB_key(c)) and retrive_vox_D(c, &alloc[id])
// (include/vox_hash/vhashing.h:124-134, include/par_wave/voxmap_utils.cuh:126-132) on the AoS GlbVoxel blocks.  Here blocks
// are field-major pools behind an open-addressing table; these accessors give the same answers:
//   gie_dv_find_block(view, get_VB_key(c))  -> block index or -1
//   gie_dv_voxel(view, c, &vox)             -> fills a GlbVoxel in the reference's layout, false when the block is not allocated
//   gie_dv_dist_sq / gie_dv_type            -> single fields without assembling the record
// The view comes from gie_hashmap_device_view (include/gie_b200.h).  nvcc only.
#pragma once
#include "gie_b200.h"
#include "par_wave/voxmap_utils.cuh"

__device__ __forceinline__ unsigned long long gie_dv_pack_key(int3 k)
{
    return ((unsigned long long)(unsigned)(k.x & 0x1fffff)) | ((unsigned long long)(unsigned)(k.y & 0x1fffff) << 21) |
           ((unsigned long long)(unsigned)(k.z & 0x1fffff) << 42);
}
__device__ __forceinline__ int gie_dv_find_block(const gie_device_view &v, int3 key)
{
    const unsigned long long k = gie_dv_pack_key(key);
    unsigned long long x = k;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    unsigned s = (unsigned)x & v.cap_mask;
    for (unsigned probes = 0; probes <= v.cap_mask; probes++) {
        const unsigned long long cur = v.keys[s];
        if (cur == k) { const int b = v.vals[s]; return (b < 0 || b == 0x7fffffff) ? -1 : b; }
        if (cur == ~0ULL) return -1;
        s = (s + 1) & v.cap_mask;
    }
    return -1;
}
// index of a global voxel inside the pools, or -1
__device__ __forceinline__ long long gie_dv_index(const gie_device_view &v, int3 c)
{
    const int b = gie_dv_find_block(v, get_VB_key(c));
    if (b < 0) return -1;
    return (long long)b * 512 + (c.z & 7) * 64 + (c.y & 7) * 8 + (c.x & 7);
}
__device__ __forceinline__ bool gie_dv_voxel(const gie_device_view &v, int3 c, GlbVoxel *out)
{
    const long long i = gie_dv_index(v, c);
    if (i < 0) return false;
    out->occ_val = v.occ_val[i]; out->vox_type = v.vox_type[i]; out->update_ct = v.update_ct[i];
    const unsigned long long p = v.coc_glb[i];
    out->coc_glb = make_int3((int)(p & 0x1fffff) - (1 << 20), (int)((p >> 21) & 0x1fffff) - (1 << 20), (int)((p >> 42) & 0x1fffff) - (1 << 20));
    out->dist_sq = v.dist_sq[i]; out->wave_layer = v.wave_layer[i];
    const unsigned long long pr = v.pair[i];
    out->dist_id_pair.ulong = (pr >> 32) | (pr << 32);   // reference word order: sq_dist[0] = dist, parent_loc_id[1] = id
    return true;
}
__device__ __forceinline__ int gie_dv_dist_sq(const gie_device_view &v, int3 c) { const long long i = gie_dv_index(v, c); return i < 0 ? EMPTY_VALUE : v.dist_sq[i]; }
__device__ __forceinline__ int gie_dv_type(const gie_device_view &v, int3 c) { const long long i = gie_dv_index(v, c); return i < 0 ? VOXTYPE_UNKNOWN : v.vox_type[i]; }
