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
#include <cuda/barrier>

namespace {

__global__ void k_build_btab(HashDev h, int entries)
{
    gie_pdl_sync();
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
            // pool already exhausted: do not burn a key slot for an insert that cannot succeed
            if (*(volatile int *)h.block_count >= h.block_max) { atomicOr(h.status, GIE_DEV_ERR_OUT_OF_BLOCKS); return -1; }
            unsigned long long prev = atomicCAS(&h.keys[s], ~0ULL, k);
            if (prev == ~0ULL) {
                int b = atomicAdd(h.block_count, 1);
                if (b >= h.block_max) {
                    atomicOr(h.status, GIE_DEV_ERR_OUT_OF_BLOCKS);
                    atomicExch(&h.vals[s], GIE_HASH_POISON);   // poison: readers stop spinning and see "no block"
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
            return v == GIE_HASH_POISON ? -1 : v;
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

// updateHashOGMWithPntCld / updateHashOGMWithSensor (unify_helper.cuh:35-197) + allocHashTB (glb_hash_map.cu:58-113).
//
// Block-centric: the sensor kernels flag the blocks they wrote into (HashDev::touched), and the merge visits the blocks that
// are touched or allocated and intersect the local volume — a few percent of the table in the headline scene — one block
// per CTA pass: counters are read and reset only inside touched blocks, a touched block with an observed voxel is
// allocated on the spot (before any of its voxels is merged, like the reference's allocate-everything-first order), hash
// pools are accessed as 512 consecutive entries per field, and glb_type (cleared to UNKNOWN for the whole volume by a
// memset) is written for the voxels of allocated blocks.
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

__global__ void __launch_bounds__(256) k_list_merge_blocks(LocDev m, HashDev h, int entries, int *__restrict__ list, int *__restrict__ count)
{
    gie_pdl_sync();
    const int lane = threadIdx.x & 31;
    const int padded = (entries + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += gridDim.x * blockDim.x) {
        bool take = false;
        if (i < entries && (__ldcg(&h.btab[i]) >= 0 || h.touched[i])) {
            int3 k = make_int3(i % h.tab_dim.x, (i / h.tab_dim.x) % h.tab_dim.y, i / (h.tab_dim.x * h.tab_dim.y)) + h.tab_org;
            int3 lo = make_int3(k.x * 8, k.y * 8, k.z * 8) - m.pvt;
            take = lo.x + 7 >= 0 && lo.x < m.X && lo.y + 7 >= 0 && lo.y < m.Y && lo.z + 7 >= 0 && lo.z < m.Z;
        }
        unsigned bal = __ballot_sync(0xffffffffu, take);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(count, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (take) list[base + __popc(bal & ((1u << lane) - 1))] = i;
    }
}

// glb_type must read UNKNOWN wherever no block exists.  The only voxels that can hold anything else are those of the blocks
// the PREVIOUS merge visited (k_merge_ogm writes the voxels of listed blocks; the frontier marks land inside them too), at
// the positions they had under the previous pivot — so those are cleared instead of the whole volume (134 MB per frame at
// 512^3, 1 GB at 1024^3, against a few MB here).  One CTA pass per listed block, 2 voxels per thread.
__global__ void __launch_bounds__(256) k_clear_prev_blocks(LocDev m, int3 prev_pvt, int3 tab_org, int3 tab_dim, const int *__restrict__ list,
                                                           const int *__restrict__ count, unsigned long long *__restrict__ ytab, int WY,
                                                           int *__restrict__ slice_has)
{
    gie_pdl_sync();
    if (ytab && blockIdx.x == 0)
        for (int z = threadIdx.x; z < m.Z; z += blockDim.x) slice_has[z] = 0;   // k_merge_ogm flags the slices it sets a bit in
    const int n = __ldcg(count);
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const int ti = __ldcg(&list[b]);
        const int3 k = make_int3(ti % tab_dim.x, (ti / tab_dim.x) % tab_dim.y, ti / (tab_dim.x * tab_dim.y)) + tab_org;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int v = threadIdx.x + 256 * u;
            const int3 c = make_int3(k.x * 8 + (v & 7), k.y * 8 + ((v >> 3) & 7), k.z * 8 + (v >> 6)) - prev_pvt;
            if (gie_inside_loc(m, c)) m.glb_type[gie_lidx(m, c)] = GIE_VOX_UNKNOWN;
            // the batch EDT's y-pass bits (edt.cu) of this block's columns go with it: k_merge_ogm sets this frame's
            if (ytab && (v & 0x38) == 0 && c.x >= 0 && c.x < m.X && c.z >= 0 && c.z < m.Z && c.y + 7 >= 0 && c.y < m.Y) {
                const int w0 = max(c.y, 0) >> 5, w1 = min(c.y + 7, m.Y - 1) >> 5;
                for (int wy = w0; wy <= w1; wy++) reinterpret_cast<uint32_t *>(ytab + ((size_t)c.z * WY + wy) * m.X + c.x)[0] = 0;
            }
        }
    }
}

template <bool PNTCLD>
__global__ void __launch_bounds__(256) k_merge_ogm(LocDev m, HashDev h, int map_ct, int stream, int n_obs, const float *__restrict__ obs,
                                                   const int *__restrict__ list, const int *__restrict__ count,
                                                   unsigned long long *__restrict__ ytab, int WY, int *__restrict__ slice_has)
{
    gie_pdl_sync();
    using barrier_t = cuda::barrier<cuda::thread_scope_block>;
    __shared__ int s_blk;
    __shared__ alignas(16) int8_t s_type[512];
    __shared__ int8_t s_new[512];   // the block's types after the merge (voxels inside the volume), for the y-pass bits
    __shared__ alignas(16) uint8_t s_occ[512];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier_t bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cuda::device::experimental::fence_proxy_async_shared_cta(); }
    __syncthreads();
    const int n = __ldcg(count);
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const int ti = __ldcg(&list[b]);
        const bool touched = h.touched[ti] != 0;
        int blk = __ldcg(&h.btab[ti]);
        const int3 k = make_int3(ti % h.tab_dim.x, (ti / h.tab_dim.x) % h.tab_dim.y, ti / (h.tab_dim.x * h.tab_dim.y)) + h.tab_org;
        // two voxels per thread (v = tid, tid + 256), engine order inside the block (x fastest)
        int id[2], cnt[2];
        int8_t inst[2];
        bool inside[2], observed[2];
        bool any_obs = false;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int v = threadIdx.x + 256 * u;
            const int3 c = make_int3(k.x * 8 + (v & 7), k.y * 8 + ((v >> 3) & 7), k.z * 8 + (v >> 6)) - m.pvt;
            inside[u] = gie_inside_loc(m, c);
            id[u] = inside[u] ? gie_lidx(m, c) : 0;
            cnt[u] = 0; inst[u] = GIE_VOX_UNKNOWN;
            if (inside[u] && touched) {   // reset observation of one scan (unify_helper.cuh:48-51,131-133)
                inst[u] = m.inst_type[id[u]];
                if (PNTCLD) cnt[u] = m.ray_count[id[u]];
                if (inst[u] != GIE_VOX_UNKNOWN) m.inst_type[id[u]] = GIE_VOX_UNKNOWN;
                if (PNTCLD && cnt[u] != 0) m.ray_count[id[u]] = 0;
            }
            observed[u] = PNTCLD ? (cnt[u] != 0) : (inst[u] == GIE_VOX_OCCUPIED || inst[u] == GIE_VOX_FREE);
            any_obs |= observed[u];
        }
        if (blk < 0) {   // uniform across the CTA
            if (!__syncthreads_or(any_obs)) continue;      // never observed, not allocated: glb_type stays UNKNOWN
            if (threadIdx.x == 0) {
                int res = hash_insert(h, k);
                if (res >= 0) h.btab[ti] = res;
                s_blk = res;
            }
            __syncthreads();
            blk = s_blk;
            __syncthreads();
            if (blk < 0) continue;                         // pool exhausted: status bit is set
        }
        // the block's type and occupancy bytes (2 x 512 B, contiguous in the field-major pools) staged by TMA bulk copies
        {
            barrier_t::arrival_token tok;
            if (threadIdx.x == 0) {
                cuda::device::memcpy_async_tx(s_type, h.vox_type + (size_t)blk * 512, cuda::aligned_size_t<16>(512), bar);
                cuda::device::memcpy_async_tx(s_occ, h.occ_val + (size_t)blk * 512, cuda::aligned_size_t<16>(512), bar);
                tok = cuda::device::barrier_arrive_tx(bar, 1, 1024);
            } else tok = bar.arrive();
            bar.wait(std::move(tok));
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (!inside[u]) continue;
            const int v = threadIdx.x + 256 * u;
            const size_t vi = (size_t)blk * 512 + v;
            int8_t type = s_type[v];
            const int3 glb = make_int3(k.x * 8 + (v & 7), k.y * 8 + ((v >> 3) & 7), k.z * 8 + (v >> 6));
            const bool occ_flag = n_obs > 0 && ext_obs_flag(m, glb, n_obs, obs);
            if (observed[u] || occ_flag) {
                const int8_t old_type = type;
                uint8_t occ = s_occ[v];
                if (PNTCLD) {
                    if (cnt[u] > 0 || occ_flag) set_occ_val(occ, type, 250.f, 1.f, m.thresh);
                    else {
                        float p = fminf(1.f, __fdiv_rn((float)(-cnt[u]), 10.f));
                        set_occ_val(occ, type, 0.f, p, m.thresh);
                    }
                } else {
                    if (inst[u] == GIE_VOX_OCCUPIED || occ_flag) set_occ_val(occ, type, 250.f, 0.8f, m.thresh);
                    else set_occ_val(occ, type, 0.f, 0.5f, m.thresh);
                }
                h.occ_val[vi] = occ;
                h.vox_type[vi] = type;
                if (stream && type != old_type) h.dirty[blk] = 1;
            }
            m.glb_type[id[u]] = type;
            s_new[v] = type;
        }
        // The batch EDT starts from OCCUPIED bits packed along y (ytab, edt.cu).  Every OCCUPIED voxel of the volume passes
        // through here, so the bits are set here, from the types the CTA holds anyway, instead of by a kernel that reads
        // glb_type back: thread = one (z, x) column of the block, 8 voxels along y, which may straddle two 32-row words.
        if (ytab) {
            __syncthreads();
            if (threadIdx.x < 64) {
                const int col = threadIdx.x;
                const int x = k.x * 8 + (col & 7) - m.pvt.x, z = k.z * 8 + (col >> 3) - m.pvt.z, y0 = k.y * 8 - m.pvt.y;
                if (x >= 0 && x < m.X && z >= 0 && z < m.Z) {
                    uint32_t bits[2] = { 0, 0 };
                    const int w0 = max(y0, 0) >> 5;
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        const int y = y0 + r;
                        if (y >= 0 && y < m.Y && s_new[(col >> 3) * 64 + r * 8 + (col & 7)] == GIE_VOX_OCCUPIED) bits[(y >> 5) - w0] |= 1u << (y & 31);
                    }
                    if (bits[0]) atomicOr(reinterpret_cast<uint32_t *>(ytab + ((size_t)z * WY + w0) * m.X + x), bits[0]);
                    if (bits[1]) atomicOr(reinterpret_cast<uint32_t *>(ytab + ((size_t)z * WY + w0 + 1) * m.X + x), bits[1]);
                    if (bits[0] | bits[1]) slice_has[z] = 1;
                }
            }
        }
        __syncthreads();   // the staging buffers are free for the next block
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
    GIE_CUDA_CHECK(cudaMemsetAsync(hm->d.touched, 0, hm->tab_entries, lm->stream));
    gie_launch(k_build_btab, dim3((entries + 255) / 256), dim3(256), 0, lm->stream, hm->d, entries);
    lm->launches++;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_update_ogm(gie_hashmap *hm, int input_pntcld, int map_ct, int stream, int n_obs)
{
    const float *obs = hm->obs_dev;
    gie_locmap *lm = hm->lm;
    StageTimer t(lm, GIE_ST_HASH_MERGE);
    const int entries = (int)hm->tab_entries;
    // UNKNOWN wherever no block exists: clear what the previous merge may have written (the whole array only the first time
    // and after a test upload), then list the blocks of this merge into the other buffer
    // The y-pass bits of the batch EDT (lm->ytab, low word of every entry) are kept in step with glb_type: cleared with the
    // previous merge's blocks, set by k_merge_ogm.  They are in step when the last thing that wrote them was the previous merge.
    gie_hashmap::BlockList &prev = hm->blists[hm->bl_cur];
    const int WY = (lm->d.Y + 31) / 32;
    unsigned long long *ytab = (lm->ytab && !getenv("GIE_YBITS_DENSE")) ? lm->ytab : nullptr;
    const bool ytab_in_step = ytab && prev.valid && !lm->glb_type_foreign && lm->ytab_serial == hm->merge_serial;
    if (ytab && !ytab_in_step) {
        GIE_CUDA_CHECK(cudaMemsetAsync(ytab, 0, (size_t)lm->d.Z * WY * lm->d.X * 8, lm->stream));
        GIE_CUDA_CHECK(cudaMemsetAsync(lm->slice_has, 0, (size_t)lm->d.Z * 4, lm->stream));
    }
    if (prev.valid && !lm->glb_type_foreign) {
        gie_launch(k_clear_prev_blocks, dim3(lm->num_sms * 8), dim3(256), 0, lm->stream, lm->d, prev.pvt, prev.tab_org, hm->d.tab_dim, prev.list, prev.count,
                   ytab_in_step ? ytab : (unsigned long long *)nullptr, WY, lm->slice_has);
        lm->launches++;
    } else GIE_CUDA_CHECK(cudaMemsetAsync(lm->d.glb_type, 0, (size_t)lm->d.N, lm->stream));
    lm->glb_type_foreign = false;
    hm->bl_cur ^= 1;
    hm->merge_serial++;
    gie_hashmap::BlockList &cur = hm->blists[hm->bl_cur];
    cur.valid = true; cur.pvt = lm->d.pvt; cur.tab_org = hm->d.tab_org;
    GIE_CUDA_CHECK(cudaMemsetAsync(cur.count, 0, sizeof(int), lm->stream));
    gie_launch(k_list_merge_blocks, dim3(std::min((entries + 255) / 256, lm->num_sms * 8)), dim3(256), 0, lm->stream, lm->d, hm->d, entries, cur.list, cur.count);
    const int grid = lm->num_sms * 16;
    if (input_pntcld) gie_launch(k_merge_ogm<true>, dim3(grid), dim3(256), 0, lm->stream, lm->d, hm->d, map_ct, stream, n_obs, obs, cur.list, cur.count, ytab, WY, lm->slice_has);
    else gie_launch(k_merge_ogm<false>, dim3(grid), dim3(256), 0, lm->stream, lm->d, hm->d, map_ct, stream, n_obs, obs, cur.list, cur.count, ytab, WY, lm->slice_has);
    lm->ytab_serial = ytab ? hm->merge_serial : -1;
    lm->launches += 2;
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
