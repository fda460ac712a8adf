// hashmap.cu — voxel-block hash access path, block allocation and the OGM -> global-map merge.
//
// Replaces (reference repo paths):
//   include/vox_hash/vhashing.h:124-191,387-455 (lookup / insert under bucket locks), blockalloc.h:49-67
//   src/kernel/par_wave/glb_hash_map.cu:58-113  allocHashTB (thrust::sort of X*Y*Z keys + unique + copy_if + retry loop)
//   src/kernel/par_wave/alloc_helper.cuh:13-73  RequiresAllocation / TryAllocateKernel / ReturnAllocations
//   src/kernel/par_wave/unify_helper.cuh:35-197 updateHashOGMWithPntCld / updateHashOGMWithSensor
//   include/par_wave/voxmap_utils.cuh:181-200   set_hashvoxel_occ_val
//
// Mechanism: a lock-free open-addressing table (one 64-bit CAS per insert) and a per-frame dense block table built with
// ONE probe per block; the merge kernel allocates a block the first time it meets an observed voxel whose block is
// missing (a fresh block is all-UNKNOWN, so voxels of that block merged earlier with "no block" are already right).
#include "engine.h"

namespace {

__global__ void k_build_btab(HashDev h, int entries)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= entries) return;
    int3 t = make_int3(i % h.tab_dim.x, (i / h.tab_dim.x) % h.tab_dim.y, i / (h.tab_dim.x * h.tab_dim.y));
    h.btab[i] = gie_hash_find(h, t + h.tab_org);
}

// insert-or-find; returns block index or -1 when the pool / table is exhausted (status bit set)
__device__ int hash_insert(const HashDev &h, int3 key)
{
    unsigned long long k = gie_pack_key(key);
    uint32_t s = (uint32_t)gie_mix64(k) & h.cap_mask;
    for (uint32_t probes = 0; probes <= h.cap_mask; probes++) {
        unsigned long long cur = __ldcg(&h.keys[s]);
        if (cur == ~0ULL) {
            unsigned long long prev = atomicCAS(&h.keys[s], ~0ULL, k);
            if (prev == ~0ULL) {
                int b = atomicAdd(h.block_count, 1);
                if (b >= h.block_max) {
                    atomicOr(h.status, GIE_DEV_ERR_OUT_OF_BLOCKS);
                    atomicExch(&h.vals[s], 0x7fffffff);   // poison: readers stop spinning
                    return -1;
                }
                h.block_keys[b] = key;
                __threadfence();
                atomicExch(&h.vals[s], b);
                return b;
            }
            cur = prev;
        }
        if (cur == k) {
            int v;
            while ((v = *(volatile int32_t *)&h.vals[s]) < 0) { }   // volatile: the publish comes from another thread
            return v == 0x7fffffff ? -1 : v;
        }
        s = (s + 1) & h.cap_mask;
    }
    atomicOr(h.status, GIE_DEV_ERR_HASH_FULL);
    return -1;
}

// set_hashvoxel_occ_val (voxmap_utils.cuh:181-200)
__device__ __forceinline__ void set_occ_val(uint8_t &occ, int8_t &type, float val, float a, int thresh)
{
    if (type != GIE_VOX_UNKNOWN) val = __fadd_rn(__fmul_rn(a, val), __fmul_rn(__fsub_rn(1.0f, a), (float)occ));
    else val = __fadd_rn(__fmul_rn(a, val), __fmul_rn(__fsub_rn(1.0f, a), 0.0f));
    if (val > 254.f) val = 254.f;
    if (val < 1.f) val = 1.f;
    occ = (uint8_t)val;
    type = (occ > thresh) ? GIE_VOX_OCCUPIED : GIE_VOX_FREE;
}

// updateHashOGMWithPntCld / updateHashOGMWithSensor, one thread per voxel, x fastest
template <bool PNTCLD>
__global__ void __launch_bounds__(256) k_merge_ogm(LocDev m, HashDev h, int map_ct)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y, z = blockIdx.z;
    const bool valid = x < m.X;
    int3 c = make_int3(valid ? x : 0, y, z);
    int id = gie_lidx(m, c);
    int count = 0;
    int8_t inst = GIE_VOX_UNKNOWN;
    if (valid) {
        inst = m.inst_type[id];
        if (PNTCLD) {
            count = m.ray_count[id];
            if (count != 0) m.ray_count[id] = 0;
        }
        if (inst != GIE_VOX_UNKNOWN) m.inst_type[id] = GIE_VOX_UNKNOWN;
    }
    bool observed = valid && (PNTCLD ? (count != 0) : (inst == GIE_VOX_OCCUPIED || inst == GIE_VOX_FREE));
    int3 glb = c + m.pvt;
    int ti = gie_tab_index(h, glb);
    int blk = __ldcg(&h.btab[ti]);
    // warp-aggregated allocation: one lane per distinct missing block inserts, the others take its result
    {
        const bool need = blk < 0 && observed;
        const int lane = threadIdx.x & 31;
        unsigned grp = __match_any_sync(0xffffffffu, need ? ti : -1);
        int leader = __ffs(grp) - 1;
        int res = -1;
        if (need && lane == leader) {
            res = hash_insert(h, gie_vb_key(glb));
            if (res >= 0) h.btab[ti] = res;
        }
        res = __shfl_sync(0xffffffffu, res, leader);
        if (need) blk = res;
    }
    if (!valid) return;
    if (blk < 0) { m.glb_type[id] = GIE_VOX_UNKNOWN; return; }
    size_t vi = (size_t)blk * 512 + gie_vox_in_block(glb);
    int8_t type = h.vox_type[vi];
    if (observed) {
        uint8_t occ = h.occ_val[vi];
        if (PNTCLD) {
            if (count > 0) set_occ_val(occ, type, 250.f, 1.f, m.thresh);
            else {
                float p = fminf(1.f, __fdiv_rn((float)(-count), 10.f));
                set_occ_val(occ, type, 0.f, p, m.thresh);
            }
        } else {
            if (inst == GIE_VOX_OCCUPIED) set_occ_val(occ, type, 250.f, 0.8f, m.thresh);
            else set_occ_val(occ, type, 0.f, 0.5f, m.thresh);
        }
        h.occ_val[vi] = occ;
        h.vox_type[vi] = type;
    }
    m.glb_type[id] = type;
}

__global__ void k_export(HashDev h, int nblocks, gie_glbvoxel *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks * 512) return;
    int b = i >> 9, v = i & 511;                 // v: engine order (z&7)*64 + (y&7)*8 + (x&7)
    int3 l = make_int3(v & 7, (v >> 3) & 7, v >> 6);
    gie_glbvoxel o;
    o.occ_val = h.occ_val[i]; o.vox_type = h.vox_type[i]; o.update_ct = h.update_ct[i];
    int3 coc = gie_unpack_coc(h.coc_glb[i]);
    o.coc_glb[0] = coc.x; o.coc_glb[1] = coc.y; o.coc_glb[2] = coc.z;
    o.dist_sq = h.dist_sq[i]; o.wave_layer = h.wave_layer[i]; o.dist_id_pair = h.pair[i];
    out[(size_t)b * 512 + gie_ref_vox_in_block(l)] = o;
}

}  // namespace

int gie_hash_begin_frame(gie_hashmap *hm)
{
    gie_locmap *lm = hm->lm;
    hm->d.tab_org = gie_vb_key(lm->d.pvt) - make_int3(hm->halo_blocks, hm->halo_blocks, hm->halo_blocks);
    int entries = (int)hm->tab_entries;
    k_build_btab<<<(entries + 255) / 256, 256, 0, lm->stream>>>(hm->d, entries);
    lm->launches++;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_update_ogm(gie_hashmap *hm, int input_pntcld, int map_ct)
{
    gie_locmap *lm = hm->lm;
    StageTimer t(lm, GIE_ST_HASH_MERGE);
    dim3 block(256), grid((lm->d.X + 255) / 256, lm->d.Y, lm->d.Z);
    if (lm->d.X <= 128) { block = dim3(128); grid.x = (lm->d.X + 127) / 128; }
    if (input_pntcld) k_merge_ogm<true><<<grid, block, 0, lm->stream>>>(lm->d, hm->d, map_ct);
    else k_merge_ogm<false><<<grid, block, 0, lm->stream>>>(lm->d, hm->d, map_ct);
    lm->launches++;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_export(gie_hashmap *hm, int nblocks, gie_glbvoxel *out_dev)
{
    if (nblocks <= 0) return GIE_OK;
    long long tot = (long long)nblocks * 512;
    k_export<<<(unsigned)((tot + 255) / 256), 256, 0, hm->lm->stream>>>(hm->d, nblocks, out_dev);
    hm->lm->launches++;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}
