// common.cuh — device-side data layout shared by all kernels of the engine.
//
// Layout in HBM:
//   local volume : x-fastest dense arrays, one per field (reference: LocMap, local_batch.h:540-562)
//   global map   : open-addressing hash (packed block key -> block index) + FIELD-MAJOR block pools;
//                  inside a block voxels are x-fastest ((z&7)*64 + (y&7)*8 + (x&7)) so that a warp walking x in the
//                  local volume reads whole 32 B sectors of every pool.  The reference's AoS/z-fastest GlbVoxel
//                  (voxmap_utils.cuh:29-48,103-109) is produced on export only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// first statement of every kernel launched through gie_launch (engine.h): let the next kernel of the stream be placed, then
// wait for the previous one to complete
#if defined(__CUDACC__)
__device__ __forceinline__ void gie_pdl_sync()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

#define GIE_VOX_UNKNOWN 0
#define GIE_VOX_FREE 1
#define GIE_VOX_OCCUPIED 2
#define GIE_VOX_FNT 3
#define GIE_EMPTY_VALUE 999999
#define GIE_WL_BLACK 16677223
#define GIE_WL_GRAY0 16677219
#define GIE_WL_GRAY1 16677220
#define GIE_RAISE_TAG (1ULL << 62)
#define GIE_INVALID_ID_STALE 0xfffffffeu
#define GIE_EMPTY_COC_PACKED (gie_pack_coc(make_int3(GIE_EMPTY_VALUE, GIE_EMPTY_VALUE, GIE_EMPTY_VALUE)))

// device error bits (sticky, in HashDev::status)
#define GIE_DEV_ERR_OUT_OF_BLOCKS 1
#define GIE_DEV_ERR_QUEUE_OVERFLOW 2
#define GIE_DEV_ERR_HASH_FULL 4
#define GIE_HASH_POISON 0x7fffffff   // value of a key slot whose insert found the block pool exhausted

struct LocDev {
    int X, Y, Z, N;
    float w;
    int thresh;
    float min_h, max_h;
    int cutoff_sq, fast;
    int max_width, max_loc_dist_sq;
    int3 pvt, upvt, half;
    float L2G[12], G2L[12];
    float3 origin;
    int32_t *ray_count;
    int8_t *inst_type;
    int8_t *glb_type;
    float *edt;
    int32_t *aux;       // batch dist_sq (later patched by mark)       == reference _aux
    int32_t *coc_aux;   // batch coc, local coords packed 11/11/10     == reference _coc_idx_aux
    int32_t *wave_layer;
    unsigned long long *pair;  // (dist_sq << 32) | wave-range coc id  == reference _dist_id_pair
    uint8_t *nbr_flag;         // 1 = the closest obstacle of this known voxel lies outside the local volume but inside the wave range
                               // (written with the pair by k_mark_blocks; k_frontiers tests it for six neighbours instead of decoding six pairs)
    // --- volumes sharded over GPUs ---
    // A slab map holds rows [ys0, ys0 + ysn) of the batch-EDT arrays (g2, cxy, aux, coc_aux are [Z][ysn][X]); the sweeps
    // take their work items from that range.  A whole map has ys0 = 0, ysn = Y.
    int ys0, ysn;
    // The map that runs the sparse stages of a sharded volume reads the batch-EDT result (aux, coc_aux) out of the slabs,
    // its own or a peer GPU's (mapped through CUDA IPC): n_slabs > 1, slab g holds rows [g * slab_rows, (g + 1) * slab_rows).
    int n_slabs, slab_rows;
    int32_t *aux_s[8], *coc_s[8];
};

struct HashDev {
    unsigned long long *keys;  // packed block key or ~0
    int32_t *vals;
    uint32_t cap_mask;
    int block_max;
    int *block_count;
    int *status;
    int3 *block_keys;          // per allocated block
    // field-major pools, index = block * 512 + voxel
    uint8_t *occ_val;
    int8_t *vox_type;
    int32_t *update_ct;
    unsigned long long *coc_glb;  // 21 bits per axis, biased by 2^20
    int32_t *dist_sq;
    int32_t *wave_layer;
    unsigned long long *pair;
    // per-frame dense block table around the local volume (one hash probe per block, not per voxel)
    int32_t *btab;
    uint8_t *touched;          // per table entry: a sensor kernel wrote a counter / a type inside this block this frame
    uint8_t *dirty;            // per block: changed since the last gie_hashmap_stream_changed (reference: stream_VB_keys_D)
    int3 tab_org;   // block coords of table entry (0,0,0)
    int3 tab_dim;
};

__host__ __device__ __forceinline__ unsigned long long gie_pack_key(int3 k)
{
    return ((unsigned long long)(uint32_t)(k.x & 0x1fffff)) | ((unsigned long long)(uint32_t)(k.y & 0x1fffff) << 21) |
           ((unsigned long long)(uint32_t)(k.z & 0x1fffff) << 42);
}
__host__ __device__ __forceinline__ unsigned long long gie_pack_coc(int3 c)
{
    return ((unsigned long long)(uint32_t)((c.x + (1 << 20)) & 0x1fffff)) |
           ((unsigned long long)(uint32_t)((c.y + (1 << 20)) & 0x1fffff) << 21) |
           ((unsigned long long)(uint32_t)((c.z + (1 << 20)) & 0x1fffff) << 42);
}
__host__ __device__ __forceinline__ int3 gie_unpack_coc(unsigned long long p)
{
    return make_int3((int)(p & 0x1fffff) - (1 << 20), (int)((p >> 21) & 0x1fffff) - (1 << 20),
                     (int)((p >> 42) & 0x1fffff) - (1 << 20));
}
__host__ __device__ __forceinline__ unsigned long long gie_mix64(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
__host__ __device__ __forceinline__ int3 operator+(int3 a, int3 b) { return make_int3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ int3 operator-(int3 a, int3 b) { return make_int3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ bool eq3(int3 a, int3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
__host__ __device__ __forceinline__ int sqd3(int3 a, int3 b)
{
    int dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return dx * dx + dy * dy + dz * dz;
}
// reference: get_VB_key (voxmap_utils.cuh:93-101)
__host__ __device__ __forceinline__ int3 gie_vb_key(int3 c) { return make_int3(c.x >> 3, c.y >> 3, c.z >> 3); }
// engine-internal voxel order inside a block (x fastest)
__host__ __device__ __forceinline__ int gie_vox_in_block(int3 c) { return (c.z & 7) * 64 + (c.y & 7) * 8 + (c.x & 7); }
// reference voxel order (voxmap_utils.cuh:103-109), used by export
__host__ __device__ __forceinline__ int gie_ref_vox_in_block(int3 c) { return (c.x & 7) * 64 + (c.y & 7) * 8 + (c.z & 7); }

__host__ __device__ __forceinline__ unsigned long long gie_mk_pair(int dist, uint32_t id)
{
    return ((unsigned long long)(uint32_t)dist << 32) | id;
}
__host__ __device__ __forceinline__ int gie_pair_dist(unsigned long long p) { return (int)(uint32_t)(p >> 32); }
__host__ __device__ __forceinline__ uint32_t gie_pair_id(unsigned long long p) { return (uint32_t)p; }
// wave-range coc codec, local_batch.h:12-17,173-208
__host__ __device__ __forceinline__ int3 gie_id2wr(uint32_t id)
{
    return make_int3((int)(id & 0x7ff), (int)((id >> 11) & 0x7ff), (int)((id >> 22) & 0x3ff));
}
__host__ __device__ __forceinline__ uint32_t gie_wr2id(int3 c)
{
    return (uint32_t)c.x | ((uint32_t)c.y << 11) | ((uint32_t)c.z << 22);
}
#define GIE_WR_X 2046
#define GIE_WR_Y 2046
#define GIE_WR_Z 1022
__host__ __device__ __forceinline__ bool gie_inside_wr(int3 c)
{
    return !(c.x < 0 || c.x >= GIE_WR_X || c.y < 0 || c.y >= GIE_WR_Y || c.z < 0 || c.z >= GIE_WR_Z);
}
__host__ __device__ __forceinline__ bool gie_inside_loc(const LocDev &m, int3 c)
{
    return !(c.x < 0 || c.x >= m.X || c.y < 0 || c.y >= m.Y || c.z < 0 || c.z >= m.Z);
}
__host__ __device__ __forceinline__ int gie_lidx(const LocDev &m, int3 c) { return c.x + c.y * m.X + c.z * m.X * m.Y; }
// batch-EDT result of local voxel c: the map's own array, or the slab (possibly on a peer GPU) that holds row c.y
__host__ __device__ __forceinline__ int32_t *gie_aux_ptr(const LocDev &m, int3 c)
{
    if (m.n_slabs <= 1) return m.aux + gie_lidx(m, c);
    const int g = c.y / m.slab_rows;
    return m.aux_s[g] + ((size_t)c.z * m.slab_rows + (c.y - g * m.slab_rows)) * m.X + c.x;
}
__host__ __device__ __forceinline__ int32_t *gie_coc_aux_ptr(const LocDev &m, int3 c)
{
    if (m.n_slabs <= 1) return m.coc_aux + gie_lidx(m, c);
    const int g = c.y / m.slab_rows;
    return m.coc_s[g] + ((size_t)c.z * m.slab_rows + (c.y - g * m.slab_rows)) * m.X + c.x;
}
// voxmap_utils.cuh:161-172
__host__ __device__ __forceinline__ bool gie_invalid_dist_glb(int d) { return d < 0 || d >= 900000; }
__host__ __device__ __forceinline__ bool gie_invalid_coc_glb(int3 c) { return c.x > 900000 || c.y > 900000 || c.z > 900000; }

#ifdef __CUDACC__
// hash probe (read-only)
__device__ __forceinline__ int gie_hash_find(const HashDev &h, int3 key)
{
    unsigned long long k = gie_pack_key(key);
    uint32_t s = (uint32_t)gie_mix64(k) & h.cap_mask;
    for (uint32_t probes = 0; probes <= h.cap_mask; probes++) {   // bounded: a full table ends the walk
        unsigned long long cur = __ldcg(&h.keys[s]);
        if (cur == k) {
            int v;
            // the value is published after the key; spin until visible (insert is two stores)
            while ((v = *(volatile int32_t *)&h.vals[s]) < 0) { }
            return v == GIE_HASH_POISON ? -1 : v;   // key claimed while the block pool was exhausted: no block behind it
        }
        if (cur == ~0ULL) return -1;
        s = (s + 1) & h.cap_mask;
    }
    return -1;
}
// block index for a global voxel coordinate: dense table first, hash probe outside it
__device__ __forceinline__ int gie_block_of(const HashDev &h, int3 glb)
{
    int3 key = gie_vb_key(glb);
    int3 t = key - h.tab_org;
    if (t.x >= 0 && t.x < h.tab_dim.x && t.y >= 0 && t.y < h.tab_dim.y && t.z >= 0 && t.z < h.tab_dim.z)
        return __ldg(&h.btab[(t.z * h.tab_dim.y + t.y) * h.tab_dim.x + t.x]);
    return gie_hash_find(h, key);
}
__device__ __forceinline__ int gie_tab_index(const HashDev &h, int3 glb)
{
    int3 t = gie_vb_key(glb) - h.tab_org;
    return (t.z * h.tab_dim.y + t.y) * h.tab_dim.x + t.x;
}
// the sensor kernels record which blocks they wrote into (plain byte stores of the same value: no atomics needed); the merge
// then visits only touched or allocated blocks instead of streaming the whole volume
__device__ __forceinline__ void gie_touch_block(const HashDev &h, int3 glb) { h.touched[gie_tab_index(h, glb)] = 1; }
#endif
