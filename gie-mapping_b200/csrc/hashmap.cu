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
#include <algorithm>

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

// updateHashOGMWithPntCld / updateHashOGMWithSensor.  Each thread owns VEC consecutive voxels along x (vector loads /
// stores on the dense arrays); a grid-stride loop over a grid sized to the SM count keeps CTAs resident.
// external obstacles (unify_helper.cuh:68-86,149-162; insideAABB voxmap_utils.cuh:203-207): box 0 is a fence (obstacle
// when OUTSIDE it), boxes 1.. are obstacles inside.  obs[i] = {ll.xyz, ur.xyz, activated}
__device__ __forceinline__ bool ext_obs_flag(const LocDev &m, int3 glb, int n_obs, const float *__restrict__ obs)
{
    const float px = (float)glb.x * m.w, py = (float)glb.y * m.w, pz = (float)glb.z * m.w;
    for (int i = 0; i < n_obs; i++) {
        const float *o = obs + 7 * i;
        if (__ldg(&o[6]) == 0.f) continue;
        bool in = (px >= __ldg(&o[0]) && py >= __ldg(&o[1]) && pz >= __ldg(&o[2])) && (px <= __ldg(&o[3]) && py <= __ldg(&o[4]) && pz <= __ldg(&o[5]));
        if (i == 0) { if (!in) return true; }
        else if (in) return true;
    }
    return false;
}

template <bool PNTCLD, int VEC>
__global__ void __launch_bounds__(256) k_merge_ogm(LocDev m, HashDev h, int map_ct, int stream, int n_obs, const float *__restrict__ obs,
                                                   int tpr_log2)
{
    // A CTA pass covers 256 >> tpr_log2 rows (y,z) of the volume with 1 << tpr_log2 threads per row; everything that
    // depends only on (y,z) — the block-table row, the voxel offset inside a block — is computed once per row, and all loop
    // bounds are uniform so that the rare allocation path can use warp collectives.
    const int gx = m.X / VEC;                        // X % VEC == 0: a group of VEC voxels never straddles a row
    const int tpr = 1 << tpr_log2, rpc = 256 >> tpr_log2;
    const int r_in = threadIdx.x >> tpr_log2, xg0 = threadIdx.x & (tpr - 1);
    const int nrows = m.Y * m.Z;
    const int lane = threadIdx.x & 31;
    for (int row0 = blockIdx.x * rpc; row0 < nrows; row0 += gridDim.x * rpc) {
        const int row = row0 + r_in;
        const bool row_ok = row < nrows;
        const int y = row_ok ? row % m.Y : 0, z = row_ok ? row / m.Y : 0;
        const int gy = y + m.pvt.y, gz = z + m.pvt.z;
        const int trow = (((gz >> 3) - h.tab_org.z) * h.tab_dim.y + ((gy >> 3) - h.tab_org.y)) * h.tab_dim.x - h.tab_org.x;
        const int vrow = (gz & 7) * 64 + (gy & 7) * 8;
        for (int xgb = 0; xgb < gx; xgb += tpr) {
            const int xg = xgb + xg0;
            const bool valid = row_ok && xg < gx;
            const int x0 = xg * VEC;
            const int id0 = valid ? row * m.X + x0 : 0;
            const int gx0 = x0 + m.pvt.x;
            int cnt[VEC], ti[VEC], blk[VEC];
            int8_t inst[VEC], out_type[VEC];
            if (VEC >= 4) {   // VEC consecutive voxels per thread: 16-byte loads of the counters, all issued before any use
#pragma unroll
                for (int v = 0; v < VEC; v += 4) {
                    char4 i4 = valid ? *reinterpret_cast<const char4 *>(m.inst_type + id0 + v) : make_char4(0, 0, 0, 0);
                    inst[v] = i4.x; inst[v + 1] = i4.y; inst[v + 2] = i4.z; inst[v + 3] = i4.w;
                    int4 c4 = make_int4(0, 0, 0, 0);
                    if (PNTCLD && valid) c4 = *reinterpret_cast<const int4 *>(m.ray_count + id0 + v);
                    cnt[v] = c4.x; cnt[v + 1] = c4.y; cnt[v + 2] = c4.z; cnt[v + 3] = c4.w;
                }
            } else {
                inst[0] = valid ? m.inst_type[id0] : 0;
                cnt[0] = (PNTCLD && valid) ? m.ray_count[id0] : 0;
            }
            bool any_cnt = false, any_inst = false, any_need = false, any_blk = false;
            bool observed[VEC];
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                any_cnt |= cnt[k] != 0; any_inst |= inst[k] != GIE_VOX_UNKNOWN;
                observed[k] = valid && (PNTCLD ? (cnt[k] != 0) : (inst[k] == GIE_VOX_OCCUPIED || inst[k] == GIE_VOX_FREE));
                ti[k] = trow + ((gx0 + k) >> 3);
                if (k > 0 && ti[k] == ti[k - 1]) blk[k] = blk[k - 1];
                else blk[k] = valid ? __ldcg(&h.btab[ti[k]]) : -1;
                any_need |= observed[k] && blk[k] < 0;
                any_blk |= blk[k] >= 0;
            }
            if (__any_sync(0xffffffffu, any_need)) {
                // rare after the first frames.  Warp-aggregated allocation: one lane per distinct missing block inserts,
                // the others take its result
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    bool need = observed[k] && blk[k] < 0;
                    if (need) {   // a sibling voxel of this thread (or another warp) may have created the block meanwhile
                        int b = __ldcg(&h.btab[ti[k]]);
                        if (b >= 0) { blk[k] = b; need = false; }
                    }
                    unsigned grp = __match_any_sync(0xffffffffu, need ? ti[k] : -1);
                    int leader = __ffs(grp) - 1;
                    int res = -1;
                    if (need && lane == leader) {
                        res = hash_insert(h, make_int3((gx0 + k) >> 3, gy >> 3, gz >> 3));
                        if (res >= 0) h.btab[ti[k]] = res;
                    }
                    res = __shfl_sync(0xffffffffu, res, leader);
                    if (need) blk[k] = res;
                    any_blk |= blk[k] >= 0;
                }
            }
            if (!valid) continue;
#pragma unroll
            for (int k = 0; k < VEC; k++) out_type[k] = GIE_VOX_UNKNOWN;
            if (any_blk) {
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    if (blk[k] < 0) continue;
                    const int3 glb = make_int3(gx0 + k, gy, gz);
                    size_t vi = (size_t)blk[k] * 512 + vrow + (glb.x & 7);
                    int8_t type = h.vox_type[vi];
                    const bool occ_flag = n_obs > 0 && ext_obs_flag(m, glb, n_obs, obs);
                    if (observed[k] || occ_flag) {
                        const int8_t old_type = type;
                        uint8_t occ = h.occ_val[vi];
                        if (PNTCLD) {
                            if (cnt[k] > 0 || occ_flag) set_occ_val(occ, type, 250.f, 1.f, m.thresh);
                            else {
                                float p = fminf(1.f, __fdiv_rn((float)(-cnt[k]), 10.f));
                                set_occ_val(occ, type, 0.f, p, m.thresh);
                            }
                        } else {
                            if (inst[k] == GIE_VOX_OCCUPIED || occ_flag) set_occ_val(occ, type, 250.f, 0.8f, m.thresh);
                            else set_occ_val(occ, type, 0.f, 0.5f, m.thresh);
                        }
                        h.occ_val[vi] = occ;
                        h.vox_type[vi] = type;
                        if (stream && type != old_type) h.dirty[blk[k]] = 1;
                    }
                    out_type[k] = type;
                }
            }
            if (VEC >= 4) {
#pragma unroll
                for (int v = 0; v < VEC; v += 4) {
                    if (any_cnt) *reinterpret_cast<int4 *>(m.ray_count + id0 + v) = make_int4(0, 0, 0, 0);
                    if (any_inst) *reinterpret_cast<char4 *>(m.inst_type + id0 + v) = make_char4(0, 0, 0, 0);
                    *reinterpret_cast<char4 *>(m.glb_type + id0 + v) = make_char4(out_type[v], out_type[v + 1], out_type[v + 2], out_type[v + 3]);
                }
            } else {
                if (any_cnt) m.ray_count[id0] = 0;
                if (any_inst) m.inst_type[id0] = GIE_VOX_UNKNOWN;
                m.glb_type[id0] = out_type[0];
            }
        }
    }
}

// Allocation-only pre-pass, used when external-obstacle boxes are active: the reference allocates every touched block
// before the merge (allocHashTB), so a box voxel must see a block that another voxel's observation creates this frame.
template <bool PNTCLD>
__global__ void __launch_bounds__(256) k_alloc_observed(LocDev m, HashDev h)
{
    const int n_pad = (m.N + 31) & ~31;
    const int lane = threadIdx.x & 31;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < n_pad; id += gridDim.x * blockDim.x) {
        const bool valid = id < m.N;
        bool observed = false;
        int ti = -1;
        int3 glb = make_int3(0, 0, 0);
        if (valid) {
            observed = PNTCLD ? (m.ray_count[id] != 0) : (m.inst_type[id] == GIE_VOX_OCCUPIED || m.inst_type[id] == GIE_VOX_FREE);
            glb = make_int3(id % m.X, (id / m.X) % m.Y, id / (m.X * m.Y)) + m.pvt;
            ti = gie_tab_index(h, glb);
        }
        const bool need = observed && __ldcg(&h.btab[ti]) < 0;
        unsigned grp = __match_any_sync(0xffffffffu, need ? ti : -1);
        if (need && lane == __ffs(grp) - 1) {
            int res = hash_insert(h, gie_vb_key(glb));
            if (res >= 0) h.btab[ti] = res;
        }
    }
}

// ordered list of the blocks whose dirty flag is set (one CTA-wide scan per 1024 blocks, atomics only per CTA)
__global__ void __launch_bounds__(256) k_list_changed(HashDev h, int nblocks, int *list, int *count, int clear)
{
    __shared__ int base;
    __shared__ int wcnt[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b0 = blockIdx.x * blockDim.x; b0 < nblocks; b0 += gridDim.x * blockDim.x) {
        int b = b0 + threadIdx.x;
        bool d = b < nblocks && h.dirty[b] != 0;
        unsigned bal = __ballot_sync(0xffffffffu, d);
        if (lane == 0) wcnt[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int i = 0; i < 8; i++) { int c = wcnt[i]; wcnt[i] = tot; tot += c; }
            base = tot ? atomicAdd(count, tot) : 0;
        }
        __syncthreads();
        if (d) {
            list[base + wcnt[wid] + __popc(bal & ((1u << lane) - 1))] = b;
            if (clear) h.dirty[b] = 0;
        }
        __syncthreads();
    }
}

// GlbHashMap::streamPipeline's device half (glb_hash_map.cu:232-244, getUpdatedAddr unify_helper.cuh:11-32): gather the
// listed blocks into the reference's AoS GlbVoxel layout / voxel order, one CTA per block
__global__ void __launch_bounds__(512) k_gather_changed(HashDev h, const int *__restrict__ list, int first, int32_t *keys, gie_glbvoxel *out)
{
    const int b = list[first + blockIdx.x];
    const int v = threadIdx.x;                   // engine order (z&7)*64 + (y&7)*8 + (x&7)
    const size_t i = (size_t)b * 512 + v;
    int3 l = make_int3(v & 7, (v >> 3) & 7, v >> 6);
    gie_glbvoxel o;
    o.occ_val = h.occ_val[i]; o.vox_type = h.vox_type[i]; o.update_ct = h.update_ct[i];
    int3 coc = gie_unpack_coc(h.coc_glb[i]);
    o.coc_glb[0] = coc.x; o.coc_glb[1] = coc.y; o.coc_glb[2] = coc.z;
    o.dist_sq = h.dist_sq[i]; o.wave_layer = h.wave_layer[i];
    unsigned long long p = h.pair[i];
    o.dist_id_pair = (p >> 32) | (p << 32);      // reference word order: sq_dist[0] = dist, parent_loc_id[1] = id
    out[(size_t)blockIdx.x * 512 + gie_ref_vox_in_block(l)] = o;
    if (v < 3) { int3 k = h.block_keys[b]; keys[3 * blockIdx.x + v] = v == 0 ? k.x : (v == 1 ? k.y : k.z); }
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
    o.dist_sq = h.dist_sq[i]; o.wave_layer = h.wave_layer[i];
    unsigned long long p = h.pair[i];
    o.dist_id_pair = (p >> 32) | (p << 32);      // reference word order: sq_dist[0] = dist, parent_loc_id[1] = id
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

int gie_launch_update_ogm(gie_hashmap *hm, int input_pntcld, int map_ct, int stream, int n_obs)
{
    const float *obs = hm->obs_dev;
    gie_locmap *lm = hm->lm;
    StageTimer t(lm, GIE_ST_HASH_MERGE);
    const int vec = (lm->d.X % 8 == 0) ? 8 : (lm->d.X % 4 == 0) ? 4 : 1;
    const int gx = lm->d.X / vec;
    int tpr_log2 = 5;
    while ((1 << tpr_log2) < gx && tpr_log2 < 8) tpr_log2++;
    const int rpc = 256 >> tpr_log2;
    const long long passes = ((long long)lm->d.Y * lm->d.Z + rpc - 1) / rpc;
    int grid = (int)std::min<long long>(passes, (long long)lm->num_sms * 16);
    if (n_obs > 0) {
        int g1 = (int)std::min<long long>(((long long)lm->d.N + 255) / 256, (long long)lm->num_sms * 16);
        if (input_pntcld) k_alloc_observed<true><<<g1, 256, 0, lm->stream>>>(lm->d, hm->d);
        else k_alloc_observed<false><<<g1, 256, 0, lm->stream>>>(lm->d, hm->d);
        lm->launches++;
    }
    if (vec == 8) {
        if (input_pntcld) k_merge_ogm<true, 8><<<grid, 256, 0, lm->stream>>>(lm->d, hm->d, map_ct, stream, n_obs, obs, tpr_log2);
        else k_merge_ogm<false, 8><<<grid, 256, 0, lm->stream>>>(lm->d, hm->d, map_ct, stream, n_obs, obs, tpr_log2);
    } else if (vec == 4) {
        if (input_pntcld) k_merge_ogm<true, 4><<<grid, 256, 0, lm->stream>>>(lm->d, hm->d, map_ct, stream, n_obs, obs, tpr_log2);
        else k_merge_ogm<false, 4><<<grid, 256, 0, lm->stream>>>(lm->d, hm->d, map_ct, stream, n_obs, obs, tpr_log2);
    } else {
        if (input_pntcld) k_merge_ogm<true, 1><<<grid, 256, 0, lm->stream>>>(lm->d, hm->d, map_ct, stream, n_obs, obs, tpr_log2);
        else k_merge_ogm<false, 1><<<grid, 256, 0, lm->stream>>>(lm->d, hm->d, map_ct, stream, n_obs, obs, tpr_log2);
    }
    lm->launches++;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_list_changed(gie_hashmap *hm, int nblocks, int clear)
{
    gie_locmap *lm = hm->lm;
    GIE_CUDA_CHECK(cudaMemsetAsync(hm->changed_count, 0, sizeof(int), lm->stream));
    if (nblocks > 0) {
        int grid = std::min((nblocks + 255) / 256, lm->num_sms * 8);
        k_list_changed<<<grid, 256, 0, lm->stream>>>(hm->d, nblocks, hm->changed_list, hm->changed_count, clear);
        lm->launches++;
    }
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_gather_changed(gie_hashmap *hm, int first, int n, int32_t *keys_dev, gie_glbvoxel *out_dev)
{
    if (n <= 0) return GIE_OK;
    k_gather_changed<<<n, 512, 0, hm->lm->stream>>>(hm->d, hm->changed_list, first, keys_dev, out_dev);
    hm->lm->launches++;
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
