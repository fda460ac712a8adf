// wave.cu — merge of the fresh local EDT with the global (hashed) EDT: limited-observation fix-up, frontier extraction,
// the three wavefronts and the commit.
//
// Replaces GlbHashMap::mergeNewObsv (reference src/kernel/par_wave/glb_hash_map.cu:146-207):
//   MarkLimitedObserve  src/kernel/par_wave/unify_helper.cuh:201-273
//   obtainFrontiers     unify_helper.cuh:275-446
//   parWave host loop   src/kernel/par_wave/wave_helper.h:8-93  (>= 3 blocking device<->host copies per BFS level)
//   BFS_in_block / BFS_one_layer, raise_outside / lower_outside / lower_inside, id_atomicMin
//                       src/kernel/par_wave/wave_core.cuh:9-22,103-523
//   UpdateHashBatch     unify_helper.cuh:448-523
//
// Mechanism: ONE persistent cooperative kernel runs wave A, B and C back to back with a device-side grid barrier
// between BFS levels and global ring queues; there is no host round trip inside a frame.  The relaxation is a 64-bit
// atomicMin on (dist_sq << 32 | coc id), which also makes the result independent of thread scheduling — the reference
// resolves equal-distance offers by arrival order (its distances are repeatable run to run all the same, measured on a B200);
// the deterministic rules D1-D3, D5 and what they cost in agreement (54 of 1.5 M known voxel-frames of the pinned fixtures)
// are in DESIGN.md §3.2 and are restated identically by the CPU oracle.
#include "engine.h"
#include <cooperative_groups.h>
#include <algorithm>

namespace {

constexpr int WAVE_THREADS = 512;   // one CTA per SM: the grid barrier costs grow with the number of CTAs

// counters layout (ints)
enum { C_A0 = 0, C_B0 = 3, C_C0 = 6, C_LEVA = 10, C_LEVB = 11, C_LEVC = 12, C_FA = 13, C_FB = 14, C_FC = 15,
       C_FB_AFTER_A = 16, C_FC_AFTER_B = 17, C_CUR_LEVEL = 18, C_DONE = 19, C_COUNT = 32 };

// cluster-local mode of wave C (see wave_c_local): per-CTA frontier queues in shared memory
constexpr int LQ_CAP = 2048;
struct LocalQ {
    unsigned long long q[2][LQ_CAP];
    uint32_t snap[LQ_CAP];
    int n[2];
    int total, nmax;
    int spill_base;
    int n0, gather;   // hand-off to the solo mode: CTA 0's own count as every CTA read it / entries gathered from the peers
    int solo_level, solo_cb;
};
constexpr int SOLO_ENTER = 256;   // frontier size at or below which ONE CTA runs the levels alone (6 * 256 pushes fit LQ_CAP)

// Wave-C queue entry.  low word: local voxel coords packed 10/10/10 (X, Y <= 1024, Z <= 1022), so that no kernel has to
// divide a linear index back into coordinates; high word: the coc id the PUSHER offered.  De-duplication is by that id:
// a voxel lowered by several neighbours in one level is queued once per successful atomicMin, and only the copy whose id
// equals the voxel's final pair id is processed (exactly one: a later offer only gets in if it is strictly smaller).  The
// reference (and the oracle) de-duplicate with a gray/black colour per voxel instead (wave_core.cuh:372-389); the set of
// voxels processed per level is the same, without the extra atomicExch round trip per relaxation.
#define C_ALWAYS 0xffffffffu   // seeds: already unique (colour 1 in k_frontiers / wave B), always processed
__device__ __forceinline__ unsigned long long c_entry(int3 c, uint32_t sid)
{
    return ((unsigned long long)sid << 32) | (uint32_t)(c.x | (c.y << 10) | (c.z << 20));
}
__device__ __forceinline__ int3 c_entry_coord(unsigned long long e)
{
    uint32_t p = (uint32_t)e;
    return make_int3((int)(p & 1023), (int)((p >> 10) & 1023), (int)(p >> 20));
}

struct WaveDev {
    unsigned long long *qA[3];
    unsigned long long *qB[3];
    unsigned long long *qC[3];   // inside queue: (pusher coc id << 32) | packed local coords, see c_entry
    unsigned long long *cseed_key;
    int cap;
    int *cnt;
    unsigned int *barrier;
    int32_t *dec_dist;
    unsigned long long *dec_coc;
    unsigned long long *dec_pair;
    int32_t *dec_flags;
    uint32_t *snap_id;
    int display;   // display_glb_edt: record changed blocks for streaming
    int cluster_size, local_enter, local_spill;   // cluster-local mode of wave C
    int no_solo; // GIE_WAVE_NO_SOLO: keep the tail of wave C in the cluster mode (measurement switch)
    int epoch;   // value that marks a voxel as "already a wave-C seed" in m.wave_layer for THIS merge (never reused, so the
                 // array needs no per-frame clearing; the reference rewrites _loc_wave_layer for every voxel)

    unsigned long long *trace;   // diagnostics (GIE_WAVE_TRACE=1): per wave-C level {n, t0, t_phase1, t_bar1, t_phase2, t_bar2} in ns
};
constexpr int TRACE_LEVELS = 4096;
__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");
    return t;
}

__constant__ int3 DIRS6[6] = { { -1, 0, 0 }, { 1, 0, 0 }, { 0, -1, 0 }, { 0, 1, 0 }, { 0, 0, -1 }, { 0, 0, 1 } };

__device__ __forceinline__ unsigned long long pack_glb(int3 c) { return gie_pack_coc(c); }
__device__ __forceinline__ int3 unpack_glb(unsigned long long p) { return gie_unpack_coc(p); }

__device__ __forceinline__ bool vox_ref(const HashDev &h, int3 glb, size_t &vi)
{
    int b = gie_block_of(h, glb);
    if (b < 0) return false;
    vi = (size_t)b * 512 + gie_vox_in_block(glb);
    return true;
}

template <typename T>
__device__ __forceinline__ bool q_push(T *q, int *cnt, int cap, T v, int *status)
{
    int i = atomicAdd(cnt, 1);
    if (i >= cap) { atomicOr(status, GIE_DEV_ERR_QUEUE_OVERFLOW); return false; }
    q[i] = v;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// The per-voxel kernels of the merge (frontiers, commit) walk the ALLOCATED blocks that intersect the local volume instead
// of the whole volume: a known voxel always has a block, and in the headline scene the allocated blocks cover ~5 % of the
// 512^3 volume.  k_list_blocks compacts the dense block table into that list once per merge; the kernels then take one
// block per CTA pass (hash pools are then read as 512 consecutive entries per field).
// Next to the table index the list carries the local coordinates of the block's first voxel and its pool index: decoding a
// table index takes two divisions and two remainders by run-time values, which the per-voxel kernels used to repeat for each
// of a block's 512 voxels (a quarter of k_frontiers' instructions).
__global__ void __launch_bounds__(256) k_list_blocks(LocDev m, HashDev h, int entries, int *__restrict__ list, int4 *__restrict__ org,
                                                     int *__restrict__ count)
{
    gie_pdl_sync();
    const int lane = threadIdx.x & 31;
    const int padded = (entries + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += gridDim.x * blockDim.x) {
        bool take = false;
        int blk = -1;
        int3 lo = make_int3(0, 0, 0);
        if (i < entries && (blk = __ldcg(&h.btab[i])) >= 0) {
            int3 k = make_int3(i % h.tab_dim.x, (i / h.tab_dim.x) % h.tab_dim.y, i / (h.tab_dim.x * h.tab_dim.y)) + h.tab_org;
            lo = make_int3(k.x * 8, k.y * 8, k.z * 8) - m.pvt;   // local coords of the block's first voxel
            take = lo.x + 7 >= 0 && lo.x < m.X && lo.y + 7 >= 0 && lo.y < m.Y && lo.z + 7 >= 0 && lo.z < m.Z;
        }
        unsigned bal = __ballot_sync(0xffffffffu, take);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(count, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (take) {
            const int at = base + __popc(bal & ((1u << lane) - 1));
            list[at] = i;
            org[at] = make_int4(lo.x, lo.y, lo.z, blk);
        }
    }
}
// local coordinate of voxel v (engine order, x fastest) of the block whose first voxel is at o; false when outside the volume
__device__ __forceinline__ bool block_voxel_local(const LocDev &m, int4 o, int v, int3 &c)
{
    c = make_int3(o.x + (v & 7), o.y + ((v >> 3) & 7), o.z + (v >> 6));
    return gie_inside_loc(m, c);
}

// MarkLimitedObserve (unify_helper.cuh:201-273).  The reference skips UNKNOWN voxels altogether (:217-218): their
// _dist_id_pair keeps whatever an earlier frame left at that local index, and the wavefronts later relax against those
// stale words.  That is deterministic and is reproduced here: the pair array is written for KNOWN voxels only (and by the
// waves), never cleared.  A known voxel always has a block, so the kernel walks the allocated blocks that intersect the
// volume.  Per known voxel: start from the batch result; "sees nothing" (no obstacle in the volume) -> (EMPTY, 0xffffffff);
// where the batch EDT found something farther than the distance the global map remembers to an obstacle that has left the
// volume, keep the remembered one; a coc outside the wave range invalidates the distance word only (the id word stays
// stale, :258-261).
__global__ void __launch_bounds__(256) k_mark_blocks(LocDev m, HashDev h, const int4 *__restrict__ org, const int *__restrict__ count)
{
    gie_pdl_sync();
    const int n = __ldcg(count);
    const int mw = m.max_width;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const int4 o = __ldcg(&org[b]);
        const int blk = o.w;
        for (int v = threadIdx.x; v < 512; v += blockDim.x) {
            int3 c;
            if (!block_voxel_local(m, o, v, c)) continue;
            const int id = gie_lidx(m, c);
            if (m.glb_type[id] == GIE_VOX_UNKNOWN) continue;
            const size_t vi = (size_t)blk * 512 + v;
            int32_t *const paux = gie_aux_ptr(m, c);
            const int dist_new = *paux, dist_old = h.dist_sq[vi];
            int3 coc_new = gie_id2wr((uint32_t)*gie_coc_aux_ptr(m, c));   // same 11/11/10 packing, local coords
            int aux = dist_new, pdist;
            uint32_t pid;
            bool have_id = false;
            if (coc_new.x > mw || coc_new.y > mw || coc_new.z > mw) { pid = 0xffffffffu; have_id = true; aux = GIE_EMPTY_VALUE; }   // invalid_coc_buf
            if (dist_new > dist_old) {
                const int3 coc_buf_old = gie_unpack_coc(h.coc_glb[vi]) - m.pvt;
                if (!gie_inside_loc(m, coc_buf_old)) { coc_new = coc_buf_old; aux = dist_old; }
            }
            const int3 wr = coc_new + m.pvt - m.upvt;
            if (!gie_inside_wr(wr)) {
                pdist = GIE_EMPTY_VALUE; aux = GIE_EMPTY_VALUE;
                if (!have_id) pid = gie_pair_id(m.pair[id]);     // stale id word
            } else { pdist = aux; pid = gie_wr2id(wr); }
            m.pair[id] = gie_mk_pair(pdist, pid);
            {   // what k_frontiers asks of a NEIGHBOUR's pair (unify_helper.cuh:300-312), answered once here
                const int3 fwr = gie_id2wr(pid), fcb = fwr + m.upvt - m.pvt;
                m.nbr_flag[id] = (!gie_inside_loc(m, fcb) && gie_inside_wr(fwr)) ? 1 : 0;
            }
            if (aux != dist_new) *paux = aux;
        }
    }
}

// obtainFrontiers (unify_helper.cuh:275-446) for one known voxel.  pair[] is read-only here: a lowered own pair (frontier C
// seed) is deferred into cseed_key and applied by the wave kernel, which removes the reference's _g/_coc_idx backup arrays.
__device__ __forceinline__ void frontier_voxel(const LocDev &m, const HashDev &h, const WaveDev &w, int map_ct, int3 c, int id, int8_t type)
{
    unsigned long long pr = m.pair[id];
    int3 cur_wr = gie_id2wr(gie_pair_id(pr));
    int3 cur_coc_glb = cur_wr + m.upvt;
    int3 cur_coc_buf = cur_coc_glb - m.pvt;
    int cur_dist = gie_pair_dist(pr);
    if (gie_inside_loc(m, cur_coc_buf)) {
        bool nbr_unknown = false, lowered = false;
        unsigned long long new_key = 0;
#pragma unroll
        for (int d = 0; d < 6; d++) {
            int3 nb = c + DIRS6[d];
            if (gie_inside_loc(m, nb)) {
                int nid = gie_lidx(m, nb);
                if (m.glb_type[nid] == GIE_VOX_UNKNOWN) { nbr_unknown = true; continue; }
                if (m.nbr_flag[nid]) {                   // its closest obstacle is outside the volume and inside the wave range (rare)
                    int3 nwr = gie_id2wr(gie_pair_id(m.pair[nid]));
                    int3 ncb = nwr + m.upvt - m.pvt;
                    int d2 = sqd3(ncb, c);
                    if (d2 < cur_dist) { new_key = gie_mk_pair(d2, gie_wr2id(nwr)); lowered = true; }
                }
            } else {
                int3 nglb = nb + m.pvt;
                size_t vi;
                if (!vox_ref(h, nglb, vi)) { nbr_unknown = true; continue; }
                if (h.vox_type[vi] == GIE_VOX_UNKNOWN) { nbr_unknown = true; continue; }
                int ndist = h.dist_sq[vi];
                if (gie_invalid_dist_glb(ndist)) continue;
                int3 ncoc = gie_unpack_coc(h.coc_glb[vi]);
                if (gie_invalid_coc_glb(ncoc)) continue;
                int3 nwr = ncoc - m.upvt;
                bool n_valid = gie_inside_wr(nwr);
                int3 ncb = ncoc - m.pvt;
                bool n_local = gie_inside_loc(m, ncb);
                if (!n_local && n_valid) {
                    int d2 = sqd3(ncb, c);
                    if (d2 < cur_dist) { new_key = gie_mk_pair(d2, gie_wr2id(nwr)); lowered = true; }
                }
                if (m.fast) continue;
                int c2n = sqd3(nb, cur_coc_buf);
                if (c2n < ndist) {                       // lower-out seed (frontier B)
                    h.wave_layer[vi] = 1; h.update_ct[vi] = map_ct;
                    h.pair[vi] = gie_mk_pair(c2n, gie_wr2id(cur_wr));
                    q_push(w.qB[0], &w.cnt[C_B0], w.cap, pack_glb(nglb), h.status);
                } else if (c2n > ndist && n_local) {     // raise-out seed (frontier A)
                    if (m.glb_type[gie_lidx(m, ncb)] != GIE_VOX_OCCUPIED) {
                        h.dist_sq[vi] = c2n; h.coc_glb[vi] = gie_pack_coc(cur_coc_glb); h.wave_layer[vi] = -map_ct;
                        h.pair[vi] = gie_mk_pair(c2n, gie_wr2id(cur_wr));
                        q_push(w.qA[0], &w.cnt[C_A0], w.cap, pack_glb(nglb), h.status);
                    }
                }
            }
        }
        if (lowered) {                                   // lower-in seed (frontier C)
            m.wave_layer[id] = w.epoch;
            int i = atomicAdd(&w.cnt[C_C0], 1);
            if (i < w.cap) { w.qC[0][i] = c_entry(c, C_ALWAYS); w.cseed_key[i] = new_key; }
            else atomicOr(h.status, GIE_DEV_ERR_QUEUE_OVERFLOW);
        }
        if (type == GIE_VOX_FREE && nbr_unknown) m.glb_type[id] = GIE_VOX_FNT;
    }
}
__global__ void __launch_bounds__(256) k_frontiers(LocDev m, HashDev h, WaveDev w, int map_ct, const int4 *__restrict__ org,
                                                   const int *__restrict__ count)
{
    gie_pdl_sync();
    const int n = __ldcg(count);
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const int4 o = __ldcg(&org[b]);
        for (int v = threadIdx.x; v < 512; v += blockDim.x) {
            int3 c;
            if (!block_voxel_local(m, o, v, c)) continue;
            const int id = gie_lidx(m, c);
            const int8_t type = m.glb_type[id];
            if (type != GIE_VOX_UNKNOWN) frontier_voxel(m, h, w, map_ct, c, id, type);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Grid barrier for the persistent wave kernel.  A flat atomic counter costs ~27 cycles of L2 atomic-unit time per arriving
// CTA (same-address atomics serialise), i.e. microseconds per barrier; here every CTA publishes its arrival with a plain
// release store to its OWN 128-byte line, CTA 0 polls all lines in parallel (one thread per line) and releases the grid with
// one store that the other CTAs poll.  bar[32 * cta] = arrival generation of that CTA, bar[32 * gridDim.x] = release.
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &gen)
{
    gen++;
    __syncthreads();
    unsigned *release = bar + 32 * gridDim.x;
    if (blockIdx.x == 0) {
        for (int c = 1 + threadIdx.x; c < gridDim.x; c += blockDim.x)
            while (ld_acquire_gpu(bar + 32 * c) < gen) { }
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); st_release_gpu(release, gen); }
    } else {
        if (threadIdx.x == 0) {
            __threadfence();
            st_release_gpu(bar + 32 * blockIdx.x, gen);
            while (ld_acquire_gpu(release) < gen) { }
            __threadfence();
        }
        __syncthreads();
    }
}

// Queue push aggregated per warp: every lane of a converged warp calls it with its own items (0..6); one atomicAdd per warp.
template <typename T>
__device__ __forceinline__ void q_push_warp(T *q, int *cnt, int cap, const T *items, int n_items, int *status)
{
    const int lane = threadIdx.x & 31;
    int incl = n_items;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 31) base = atomicAdd(cnt, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - n_items;
    for (int k = 0; k < n_items; k++) {
        if (base + k >= cap) { atomicOr(status, GIE_DEV_ERR_QUEUE_OVERFLOW); break; }
        q[base + k] = items[k];
    }
}

// wave A: raise_outside (wave_core.cuh:103-224), phase 1 = offers + own decision, phase 2 = apply
__device__ void waveA_phase1(const LocDev &m, const HashDev &h, const WaveDev &w, int map_ct, const unsigned long long *cur,
                             int n, unsigned long long *next, int *next_cnt, int tid, int nthreads)
{
    for (int i = tid; i < n; i += nthreads) {
        int3 cg = unpack_glb(__ldcg(&cur[i]));
        w.dec_flags[i] = 0;
        size_t vi;
        if (!vox_ref(h, cg, vi)) continue;
        int o_dist = __ldcg(&h.dist_sq[vi]);
        if (o_dist > m.cutoff_sq) continue;
        if (w.display) h.dirty[vi >> 9] = 1;   // wave_core.cuh:128-134
        int3 lcoc = gie_unpack_coc(__ldcg(&h.coc_glb[vi]));
        int3 o_coc = lcoc;
        unsigned long long o_pair = __ldcg(&h.pair[vi]);
        int3 cur_wr = lcoc - m.upvt;
        bool touched = false, in_q = false;
        for (int d = 0; d < 6; d++) {
            int3 ng = cg + DIRS6[d];
            if (gie_inside_loc(m, ng - m.pvt)) continue;
            size_t ni;
            if (!vox_ref(h, ng, ni)) continue;
            int3 ncoc = gie_unpack_coc(__ldcg(&h.coc_glb[ni]));
            int ndist = __ldcg(&h.dist_sq[ni]);
            if (h.vox_type[ni] == GIE_VOX_UNKNOWN || gie_invalid_coc_glb(ncoc) || gie_invalid_dist_glb(ndist)) continue;
            if (__ldcg(&h.wave_layer[ni]) == -map_ct || __ldcg(&h.update_ct[ni]) == -map_ct) continue;
            if (eq3(ncoc, lcoc)) continue;
            bool raised = false;
            int3 ncb = ncoc - m.pvt;
            if (gie_inside_loc(m, ncb) && *gie_aux_ptr(m, ncb) != 0) {
                unsigned long long cand = GIE_RAISE_TAG | gie_mk_pair(sqd3(lcoc, ng), gie_wr2id(cur_wr));
                unsigned long long old = __ldcg(&h.pair[ni]);
                for (;;) {
                    if (old & GIE_RAISE_TAG) { atomicMin(&h.pair[ni], cand); break; }
                    unsigned long long prev = atomicCAS(&h.pair[ni], old, cand);
                    if (prev == old) { q_push(next, next_cnt, w.cap, pack_glb(ng), h.status); break; }
                    old = prev;
                }
                raised = true;
            }
            if (!raised) {
                int d2 = sqd3(ncoc, cg);
                if (o_dist > d2) {
                    o_dist = d2; o_coc = ncoc; touched = true;
                    int3 nwr = ncoc - m.upvt;
                    if (!gie_inside_wr(nwr)) continue;
                    o_pair = gie_mk_pair(d2, gie_wr2id(nwr));
                    if (!in_q) { in_q = true; q_push(w.qB[0], &w.cnt[C_B0], w.cap, pack_glb(cg), h.status); }
                }
            }
        }
        if (touched) {
            w.dec_flags[i] = 1; w.dec_dist[i] = o_dist; w.dec_coc[i] = gie_pack_coc(o_coc); w.dec_pair[i] = o_pair;
        }
    }
}
__device__ void waveA_phase2(const LocDev &m, const HashDev &h, const WaveDev &w, int map_ct, const unsigned long long *cur,
                             int n, const unsigned long long *next, int n_next, int tid, int nthreads)
{
    for (int i = tid; i < n; i += nthreads) {
        if (!w.dec_flags[i]) continue;
        size_t vi;
        if (!vox_ref(h, unpack_glb(__ldcg(&cur[i])), vi)) continue;
        h.dist_sq[vi] = w.dec_dist[i]; h.coc_glb[vi] = w.dec_coc[i]; h.wave_layer[vi] = 1; h.update_ct[vi] = map_ct;
        h.pair[vi] = w.dec_pair[i];
    }
    for (int i = tid; i < n_next; i += nthreads) {
        size_t vi;
        if (!vox_ref(h, unpack_glb(__ldcg(&next[i])), vi)) continue;
        unsigned long long p = __ldcg(&h.pair[vi]) & ~GIE_RAISE_TAG;
        h.pair[vi] = p;
        h.dist_sq[vi] = gie_pair_dist(p);
        h.coc_glb[vi] = gie_pack_coc(gie_id2wr(gie_pair_id(p)) + m.upvt);
        h.wave_layer[vi] = -map_ct; h.update_ct[vi] = -map_ct;
    }
}

// wave B: lower_outside (wave_core.cuh:229-350)
__device__ void waveB_phase1(const LocDev &m, const HashDev &h, const WaveDev &w, const unsigned long long *cur, int n,
                             int tid, int nthreads)
{
    for (int i = tid; i < n; i += nthreads) {
        size_t vi;
        w.snap_id[i] = 0xffffffffu;
        if (!vox_ref(h, unpack_glb(__ldcg(&cur[i])), vi)) continue;
        h.wave_layer[vi] = GIE_WL_BLACK;
        if (w.display) h.dirty[vi >> 9] = 1;   // wave_core.cuh:250-256
        if (__ldcg(&h.dist_sq[vi]) > m.cutoff_sq) continue;
        unsigned long long p = __ldcg(&h.pair[vi]);
        uint32_t id = gie_pair_id(p);
        h.coc_glb[vi] = gie_pack_coc(gie_id2wr(id) + m.upvt);
        h.dist_sq[vi] = gie_pair_dist(p);
        w.snap_id[i] = id;
    }
}
__device__ void waveB_phase2(const LocDev &m, const HashDev &h, const WaveDev &w, int map_ct, const unsigned long long *cur,
                             int n, unsigned long long *next, int *next_cnt, int gray, int tid, int nthreads)
{
    const int lane = threadIdx.x & 31;
    for (int i0 = tid - lane; i0 < n; i0 += nthreads) {   // warp-uniform trip count: the pushes below are warp collectives
        const int i = i0 + lane;
        unsigned long long out_items[6];
        unsigned long long in_items[6];
        int n_out = 0, n_in = 0;
        uint32_t sid = i < n ? w.snap_id[i] : 0xffffffffu;
        if (sid != 0xffffffffu) {
            int3 cg = unpack_glb(__ldcg(&cur[i]));
            int3 coc = gie_id2wr(sid) + m.upvt;
            // the six relaxations are independent: issue every atomicMin before looking at any result
            unsigned long long key[6], old[6];
            size_t ni[6];
            int nid[6], kind[6];   // 0 = skip, 1 = outside (hash voxel), 2 = inside (local volume)
#pragma unroll
            for (int d = 0; d < 6; d++) {
                int3 ng = cg + DIRS6[d];
                int3 nb = ng - m.pvt;
                int cand = sqd3(coc, ng);
                key[d] = gie_mk_pair(cand, sid);
                kind[d] = 0; old[d] = 0; ni[d] = 0; nid[d] = 0;
                if (!gie_inside_loc(m, nb)) {
                    if (!vox_ref(h, ng, ni[d])) continue;
                    if (h.vox_type[ni[d]] == GIE_VOX_UNKNOWN) continue;
                    if (gie_invalid_coc_glb(gie_unpack_coc(__ldcg(&h.coc_glb[ni[d]])))) continue;
                    kind[d] = 1;
                    old[d] = atomicMin(&h.pair[ni[d]], key[d]);
                } else {
                    nid[d] = gie_lidx(m, nb);
                    if (*gie_aux_ptr(m, nb) > cand) { kind[d] = 2; atomicMin(&m.pair[nid[d]], key[d]); }
                }
            }
            int col[6];
#pragma unroll
            for (int d = 0; d < 6; d++) {   // colour swaps all in flight before any is looked at
                col[d] = kind[d] == 1 ? gray : w.epoch;
                if (kind[d] == 1 && key[d] < old[d]) col[d] = atomicExch(&h.wave_layer[ni[d]], gray);
                else if (kind[d] == 2) col[d] = atomicExch(&m.wave_layer[nid[d]], w.epoch);
            }
#pragma unroll
            for (int d = 0; d < 6; d++) {
                if (kind[d] == 1 && col[d] != gray) {
                    h.update_ct[ni[d]] = map_ct;
                    out_items[n_out++] = pack_glb(cg + DIRS6[d]);
                } else if (kind[d] == 2 && col[d] != w.epoch) in_items[n_in++] = c_entry(cg + DIRS6[d] - m.pvt, C_ALWAYS);
            }
        }
        q_push_warp(next, next_cnt, w.cap, out_items, n_out, h.status);
        q_push_warp(w.qC[0], &w.cnt[C_C0], w.cap, in_items, n_in, h.status);
    }
}

// wave C: lower_inside (wave_core.cuh:353-393)
// the six relaxations of one frontier voxel: keys from the voxel's snapshot coc, all atomicMins in flight together
struct CRelax { unsigned long long key[6], old[6]; int3 nb[6]; bool in[6]; };
__device__ __forceinline__ void c_relax(const LocDev &m, int3 c, uint32_t sid, CRelax &r)
{
    const int id = gie_lidx(m, c);
    const int XY = m.X * m.Y;
    const int3 dlt = gie_id2wr(sid) + m.upvt - m.pvt - c;           // coc - voxel
    const int d0 = dlt.x * dlt.x + dlt.y * dlt.y + dlt.z * dlt.z;
    // |coc - (c + e)|^2 = d0 - 2 e.dlt + 1 for a unit step e
    const int dd[6] = { d0 + 2 * dlt.x + 1, d0 - 2 * dlt.x + 1, d0 + 2 * dlt.y + 1, d0 - 2 * dlt.y + 1, d0 + 2 * dlt.z + 1, d0 - 2 * dlt.z + 1 };
    const int off[6] = { -1, 1, -m.X, m.X, -XY, XY };
    const bool ok[6] = { c.x > 0, c.x < m.X - 1, c.y > 0, c.y < m.Y - 1, c.z > 0, c.z < m.Z - 1 };
#pragma unroll
    for (int d = 0; d < 6; d++) {
        r.in[d] = ok[d];
        r.nb[d] = c + DIRS6[d];
        r.key[d] = gie_mk_pair(dd[d], sid);
        r.old[d] = ok[d] ? atomicMin(&m.pair[id + off[d]], r.key[d]) : 0ULL;
    }
}
__device__ void waveC_phase1(const LocDev &m, const WaveDev &w, const unsigned long long *cur, int n, int tid, int nthreads)
{
    for (int i = tid; i < n; i += nthreads) {
        unsigned long long e = __ldcg(&cur[i]);
        uint32_t want = (uint32_t)(e >> 32);
        uint32_t sid = gie_pair_id(__ldcg(&m.pair[gie_lidx(m, c_entry_coord(e))]));
        w.snap_id[i] = (want == C_ALWAYS || want == sid) ? sid : 0xffffffffu;   // 0xffffffff: a superseded copy, skipped
    }
}
__device__ void waveC_phase2(const LocDev &m, const HashDev &h, const WaveDev &w, const unsigned long long *cur, int n,
                             unsigned long long *next, int *next_cnt, int tid, int nthreads)
{
    const int lane = threadIdx.x & 31;
    for (int i0 = tid - lane; i0 < n; i0 += nthreads) {   // warp-uniform trip count (warp-collective push)
        const int i = i0 + lane;
        unsigned long long items[6];
        int n_items = 0;
        uint32_t sid = i < n ? w.snap_id[i] : 0xffffffffu;
        if (sid != 0xffffffffu) {
            CRelax r;
            c_relax(m, c_entry_coord(__ldcg(&cur[i])), sid, r);
#pragma unroll
            for (int d = 0; d < 6; d++)
                if (r.in[d] && r.key[d] < r.old[d]) items[n_items++] = c_entry(r.nb[d], sid);
        }
        q_push_warp(next, next_cnt, w.cap, items, n_items, h.status);
    }
}

// The tail of wave C: with a frontier of at most SOLO_ENTER voxels even the cluster barriers (2 x ~0.5 us per level) dominate,
// so CTA 0 of the cluster takes the whole frontier and runs the levels alone with __syncthreads and shared-memory queue
// atomics, until the frontier is empty or has grown again.  Same relaxation, same de-duplication.  cb = buffer holding the
// frontier of `level`, on entry and on return.
__device__ int wave_c_solo(const LocDev &m, const HashDev &h, const WaveDev &w, LocalQ &lq, int level, int &cb)
{
    const int t = threadIdx.x, T = blockDim.x, lane = t & 31;
    for (;;) {
        const int n = min(lq.n[cb], LQ_CAP);
        const bool tr = w.trace && t == 0 && level < TRACE_LEVELS;
        if (tr) { w.trace[level * 6] = 2000000000ULL + n; w.trace[level * 6 + 1] = gtimer(); }
        for (int k = t; k < n; k += T) {
            unsigned long long e = lq.q[cb][k];
            uint32_t want = (uint32_t)(e >> 32);
            uint32_t sid = gie_pair_id(__ldcg(&m.pair[gie_lidx(m, c_entry_coord(e))]));
            lq.snap[k] = (want == C_ALWAYS || want == sid) ? sid : 0xffffffffu;
        }
        if (t == 0) lq.n[cb ^ 1] = 0;
        __syncthreads();
        for (int k0 = t - lane; k0 < n; k0 += T) {   // warp-uniform trip count
            const int k = k0 + lane;
            uint32_t sid = k < n ? lq.snap[k] : 0xffffffffu;
            CRelax rx;
#pragma unroll
            for (int d = 0; d < 6; d++) { rx.in[d] = false; rx.key[d] = 0; rx.old[d] = 0; rx.nb[d] = make_int3(0, 0, 0); }
            if (sid != 0xffffffffu) c_relax(m, c_entry_coord(lq.q[cb][k]), sid, rx);
            int mine = 0;
#pragma unroll
            for (int d = 0; d < 6; d++) { rx.in[d] = rx.in[d] && rx.key[d] < rx.old[d]; mine += rx.in[d]; }
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                int base = 0;
                if (lane == 31) base = atomicAdd(&lq.n[cb ^ 1], total);
                base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    if (!rx.in[d]) continue;
                    if (base < LQ_CAP) lq.q[cb ^ 1][base] = c_entry(rx.nb[d], sid);
                    else atomicOr(h.status, GIE_DEV_ERR_QUEUE_OVERFLOW);   // cannot happen: at most 6 * SOLO_ENTER pushes
                    base++;
                }
            }
        }
        __syncthreads();
        if (tr) w.trace[level * 6 + 5] = gtimer();
        level++; cb ^= 1;
        const int nn = lq.n[cb];
        if (nn == 0 || nn > SOLO_ENTER) break;
    }
    return level;
}

// Wave C while the frontier is small (the common case: a few hundred to a few thousand voxels per level for hundreds of
// levels).  A grid-wide level costs two grid barriers plus global-queue round trips (~11 us measured); here ONE thread-block
// cluster runs the levels on its own: every CTA keeps its share of the frontier in shared memory, pushes go to a peer CTA's
// queue through distributed shared memory (one remote atomicAdd per warp + remote stores), and the two per-level
// synchronisations are hardware cluster barriers.  The relaxation itself is c_relax, as in the grid-wide phases, so results
// do not depend on the mode.  Returns the level reached; on return either the frontier is empty (all three queue counters
// are 0) or it was spilled to the global queue of that level (too large for the cluster).
__device__ int wave_c_local(const LocDev &m, const HashDev &h, const WaveDev &w, LocalQ &lq, int level, int n_in, int spill_at)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int R = (int)cl.num_blocks(), r = (int)cl.block_rank();
    const int t = threadIdx.x, T = blockDim.x, lane = t & 31, wid = t >> 5;
    if (t == 0) { lq.n[0] = r < n_in ? (n_in - r + R - 1) / R : 0; lq.n[1] = 0; }
    {
        const unsigned long long *src = w.qC[level % 3];
        for (int i = r + t * R; i < n_in; i += T * R) lq.q[0][i / R] = __ldcg(&src[i]);
    }
    cl.sync();   // every CTA of the cluster has taken its share: the global counters can be cleared
    if (r == 0 && t == 0) { w.cnt[C_C0] = 0; w.cnt[C_C0 + 1] = 0; w.cnt[C_C0 + 2] = 0; }
    int cb = 0;
    for (;;) {
        const int nloc = min(lq.n[cb], LQ_CAP);
        unsigned long long *ovf_q = w.qC[(level + 1) % 3];
        int *ovf_cnt = &w.cnt[C_C0 + (level + 1) % 3];
        const bool tr = w.trace && r == 0 && t == 0 && level < TRACE_LEVELS;
        if (tr) { w.trace[level * 6] = 1000000000ULL + nloc; w.trace[level * 6 + 1] = gtimer(); }
        // phase 1: snapshot (waveC_phase1)
        for (int k = t; k < nloc; k += T) {
            unsigned long long e = lq.q[cb][k];
            uint32_t want = (uint32_t)(e >> 32);
            uint32_t sid = gie_pair_id(__ldcg(&m.pair[gie_lidx(m, c_entry_coord(e))]));
            lq.snap[k] = (want == C_ALWAYS || want == sid) ? sid : 0xffffffffu;
        }
        if (t == 0) { lq.n[cb ^ 1] = 0; lq.gather = 0; }
        if (tr) w.trace[level * 6 + 2] = gtimer();
        cl.sync();   // all snapshots taken, all next-queue counters zero
        if (tr) w.trace[level * 6 + 3] = gtimer();
        // phase 2: offers (waveC_phase2).  The lowered neighbours of one warp's 32 frontier voxels go, as one block, to the
        // queue of a peer CTA chosen round-robin per warp and level: one remote atomicAdd per warp instead of one per voxel
        // (same-address atomics on a queue counter serialise at a few ns each).
        for (int k0 = t - lane; k0 < nloc; k0 += T) {   // warp-uniform trip count (warp collectives below)
            const int k = k0 + lane;
            uint32_t sid = k < nloc ? lq.snap[k] : 0xffffffffu;
            CRelax rx;
#pragma unroll
            for (int d = 0; d < 6; d++) { rx.in[d] = false; rx.key[d] = 0; rx.old[d] = 0; rx.nb[d] = make_int3(0, 0, 0); }
            if (sid != 0xffffffffu) c_relax(m, c_entry_coord(lq.q[cb][k]), sid, rx);
            if (tr) {
#pragma unroll
                for (int d = 0; d < 6; d++) asm volatile("" ::"l"(rx.old[d]));
                w.trace[TRACE_LEVELS * 6 + level * 4 + 0] = gtimer();
            }
            int mine = 0;
#pragma unroll
            for (int d = 0; d < 6; d++) {
                rx.in[d] = rx.in[d] && rx.key[d] < rx.old[d];   // lowered -> joins the next frontier
                mine += rx.in[d];
            }
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                const int dst = (wid + r + level) % R;
                int base = 0;
                if (lane == 31) base = atomicAdd(cl.map_shared_rank(&lq.n[cb ^ 1], dst), total);
                base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
                if (tr) { asm volatile("" ::"r"(base)); w.trace[TRACE_LEVELS * 6 + level * 4 + 1] = gtimer(); }
                unsigned long long *dq = cl.map_shared_rank(&lq.q[cb ^ 1][0], dst);
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    if (!rx.in[d]) continue;
                    unsigned long long e = c_entry(rx.nb[d], sid);
                    if (base < LQ_CAP) dq[base] = e;
                    else {   // that CTA's queue is full: the voxel goes to the global queue of the next level
                        int g = atomicAdd(ovf_cnt, 1);
                        if (g < w.cap) ovf_q[g] = e; else atomicOr(h.status, GIE_DEV_ERR_QUEUE_OVERFLOW);
                    }
                    base++;
                }
            }
        }
        if (tr) w.trace[level * 6 + 4] = gtimer();
        cl.sync();   // all pushes have landed
        if (tr) w.trace[TRACE_LEVELS * 6 + level * 4 + 2] = gtimer();
        level++; cb ^= 1;
        if (t < 32) {   // frontier size of the new level: every CTA reads all queue counters and reaches the same decision
            int c = t < R ? *cl.map_shared_rank(&lq.n[cb], t) : 0;
            int sum = c, mx = c;
#pragma unroll
            for (int o = 16; o; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
            const int c0 = __shfl_sync(0xffffffffu, c, 0);
            if (t == 0) { lq.total = sum; lq.nmax = mx; lq.n0 = c0; }
        }
        __syncthreads();
        if (tr) w.trace[(level - 1) * 6 + 5] = gtimer();
        int total = lq.total, nmax = lq.nmax;
        if (total > 0 && total <= SOLO_ENTER && nmax <= LQ_CAP && !w.no_solo) {
            // hand the whole frontier to CTA 0 and let it run alone (wave_c_solo); the other CTAs park at the cluster barrier
            // Nothing below may touch a queue counter of this level (lq.n[cb]) before the next cluster barrier: a slower
            // CTA can still be reading all of them (above) and must reach the same decision.  The hand-off therefore counts
            // in a separate word of CTA 0 (lq.gather, zeroed by CTA 0 in phase 1) on top of CTA 0's own count n0.
            if (r != 0) {
                const int mine = lq.n[cb];
                if (t == 0) lq.spill_base = lq.n0 + (mine ? atomicAdd(cl.map_shared_rank(&lq.gather, 0), mine) : 0);
                __syncthreads();
                unsigned long long *q0 = cl.map_shared_rank(&lq.q[cb][0], 0);
                for (int k = t; k < mine; k += T) q0[lq.spill_base + k] = lq.q[cb][k];
            }
            cl.sync();   // every CTA has decided, all copies have landed
            if (r != 0) { if (t == 0) { lq.n[0] = 0; lq.n[1] = 0; } }
            else { if (t == 0) lq.n[cb] += lq.gather; __syncthreads(); }
            if (r == 0) {
                level = wave_c_solo(m, h, w, lq, level, cb);
                if (t == 0) { lq.solo_level = level; lq.solo_cb = cb; }
            }
            cl.sync();
            level = *cl.map_shared_rank(&lq.solo_level, 0);
            cb = *cl.map_shared_rank(&lq.solo_cb, 0);
            total = *cl.map_shared_rank(&lq.n[cb], 0);
            nmax = total;
            cl.sync();   // everyone has read CTA 0's result before it is overwritten
        }
        if (total == 0) break;
        if (nmax > LQ_CAP || total > spill_at) {
            // hand the frontier to the grid-wide mode: append the local queues to the global queue of this level
            const int mine = min(lq.n[cb], LQ_CAP);
            if (t == 0) lq.spill_base = atomicAdd(&w.cnt[C_C0 + level % 3], mine);
            __syncthreads();
            unsigned long long *dst = w.qC[level % 3];
            for (int k = t; k < mine; k += T) {
                int g = lq.spill_base + k;
                if (g < w.cap) dst[g] = lq.q[cb][k]; else atomicOr(h.status, GIE_DEV_ERR_QUEUE_OVERFLOW);
            }
            break;
        }
    }
    if (r == 0 && t == 0) w.cnt[C_CUR_LEVEL] = level;
    cl.sync();   // no CTA leaves while a peer may still read its shared memory
    return level;
}

__global__ void __launch_bounds__(WAVE_THREADS, 1) k_waves(LocDev m, HashDev h, WaveDev w, int map_ct)
{
    __shared__ LocalQ lq;
    unsigned int gen = 0;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    volatile int *cnt = w.cnt;
    // phase 0: apply the deferred frontier-C seeds
    {
        int n = min(cnt[C_C0], w.cap);
        for (int i = tid; i < n; i += nthreads) m.pair[gie_lidx(m, c_entry_coord(__ldcg(&w.qC[0][i])))] = __ldcg(&w.cseed_key[i]);
        if (tid == 0) { w.cnt[C_FA] = min(cnt[C_A0], w.cap); w.cnt[C_FB] = min(cnt[C_B0], w.cap); w.cnt[C_FC] = n; }
    }
    grid_barrier(w.barrier, gen);
    if (!m.fast) {
        int level = 0;
        for (;; level++) {
            int ci = level % 3, ni = (level + 1) % 3, zi = (level + 2) % 3;
            int n = min(cnt[C_A0 + ci], w.cap);
            if (n == 0) break;
            if (tid == 0) w.cnt[C_A0 + zi] = 0;
            waveA_phase1(m, h, w, map_ct, w.qA[ci], n, w.qA[ni], &w.cnt[C_A0 + ni], tid, nthreads);
            grid_barrier(w.barrier, gen);
            waveA_phase2(m, h, w, map_ct, w.qA[ci], n, w.qA[ni], min(cnt[C_A0 + ni], w.cap), tid, nthreads);
            grid_barrier(w.barrier, gen);
        }
        if (tid == 0) { w.cnt[C_LEVA] = level; w.cnt[C_FB_AFTER_A] = min(cnt[C_B0], w.cap); }
        for (level = 0;; level++) {
            int ci = level % 3, ni = (level + 1) % 3, zi = (level + 2) % 3;
            int n = min(cnt[C_B0 + ci], w.cap);
            if (n == 0) break;
            if (tid == 0) w.cnt[C_B0 + zi] = 0;
            int gray = (level & 1) ? GIE_WL_GRAY1 : GIE_WL_GRAY0;
            waveB_phase1(m, h, w, w.qB[ci], n, tid, nthreads);
            grid_barrier(w.barrier, gen);
            waveB_phase2(m, h, w, map_ct, w.qB[ci], n, w.qB[ni], &w.cnt[C_B0 + ni], gray, tid, nthreads);
            grid_barrier(w.barrier, gen);
        }
        if (tid == 0) { w.cnt[C_LEVB] = level; w.cnt[C_FC_AFTER_B] = min(cnt[C_C0], w.cap); }
    }
    {
        int level = 0;
        for (;; level++) {
            int ci = level % 3, ni = (level + 1) % 3, zi = (level + 2) % 3;
            int n = min(cnt[C_C0 + ci], w.cap);
            if (n == 0) break;
            if (n <= w.local_enter) {
                grid_barrier(w.barrier, gen);   // every CTA has read n: the counters may change now
                if (blockIdx.x < w.cluster_size) wave_c_local(m, h, w, lq, level, n, w.local_spill);
                grid_barrier(w.barrier, gen);
                level = cnt[C_CUR_LEVEL] - 1;   // the loop increment brings it to the level the local mode stopped at
                continue;
            }
            if (tid == 0) w.cnt[C_C0 + zi] = 0;
            const bool tr = w.trace && tid == 0 && level < TRACE_LEVELS;
            if (tr) { w.trace[level * 6] = n; w.trace[level * 6 + 1] = gtimer(); }
            waveC_phase1(m, w, w.qC[ci], n, tid, nthreads);
            if (tr) w.trace[level * 6 + 2] = gtimer();
            grid_barrier(w.barrier, gen);
            if (tr) w.trace[level * 6 + 3] = gtimer();
            waveC_phase2(m, h, w, w.qC[ci], n, w.qC[ni], &w.cnt[C_C0 + ni], tid, nthreads);
            if (tr) w.trace[level * 6 + 4] = gtimer();
            grid_barrier(w.barrier, gen);
            if (tr) w.trace[level * 6 + 5] = gtimer();
        }
        if (tid == 0) w.cnt[C_LEVC] = level;
    }
}

// UpdateHashBatch (unify_helper.cuh:448-523), one allocated block per CTA pass (see k_list_blocks)
__global__ void __launch_bounds__(256) k_commit(LocDev m, HashDev h, int display, const int4 *__restrict__ org, const int *__restrict__ count,
                                                int *cnt, long long *stats_out)
{
    gie_pdl_sync();
    const int n = __ldcg(count);
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const int4 o = __ldcg(&org[b]);
        const int blk = o.w;
        for (int v = threadIdx.x; v < 512; v += blockDim.x) {
            int3 c;
            if (!block_voxel_local(m, o, v, c)) continue;
            const int id = gie_lidx(m, c);
            const int8_t type = m.glb_type[id];
            if (type == GIE_VOX_UNKNOWN) continue;
            unsigned long long pr = m.pair[id];
            int dist = gie_pair_dist(pr);
            uint32_t pid = gie_pair_id(pr);
            if (dist == GIE_EMPTY_VALUE) {
                if (pid == 0xffffffffu) m.edt[id] = (float)m.max_loc_dist_sq;
                continue;
            }
            size_t vi = (size_t)blk * 512 + v;
            h.coc_glb[vi] = gie_pack_coc(gie_id2wr(pid) + m.upvt);
            if (display && h.dist_sq[vi] != dist) h.dirty[blk] = 1;   // unify_helper.cuh:510-520
            h.dist_sq[vi] = dist;
            m.edt[id] = sqrtf((float)dist);
            h.pair[vi] = pr;
            if (type == GIE_VOX_FNT) h.vox_type[vi] = GIE_VOX_FNT;
        }
    }
    // the last CTA to finish publishes the frame's wavefront statistics and the sticky device status to pinned host memory
    // (read by the next frame's entry points without a synchronisation); this was a one-thread kernel of its own
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&cnt[C_DONE], 1) == (int)gridDim.x - 1) {
            __threadfence();
            stats_out[8] = *(volatile int *)h.status;
            stats_out[0] = __ldcg(&cnt[C_FA]); stats_out[1] = __ldcg(&cnt[C_FB]); stats_out[2] = __ldcg(&cnt[C_FC]);
            stats_out[3] = __ldcg(&cnt[C_LEVA]); stats_out[4] = __ldcg(&cnt[C_LEVB]); stats_out[5] = __ldcg(&cnt[C_LEVC]);
            stats_out[6] = __ldcg(&cnt[C_FB_AFTER_A]); stats_out[7] = __ldcg(&cnt[C_FC_AFTER_B]);
        }
    }
}

WaveDev make_wave_dev(gie_hashmap *hm)
{
    WaveDev w{};
    for (int i = 0; i < 3; i++) { w.qA[i] = hm->qA[i]; w.qB[i] = hm->qB[i]; w.qC[i] = hm->qC[i]; }
    w.cseed_key = hm->cseed_key; w.cap = hm->queue_cap; w.cnt = hm->counters; w.barrier = hm->barrier;
    w.dec_dist = hm->decA_dist; w.dec_coc = hm->decA_coc; w.dec_pair = hm->decA_pair; w.dec_flags = hm->decA_flags;
    w.snap_id = hm->snap_id;
    w.trace = hm->wave_trace;
    w.cluster_size = hm->wave_cluster;
    w.epoch = hm->merge_epoch;
    // enter the local mode when the frontier fits comfortably; leave it when the queues are more than half full
    w.local_enter = hm->wave_cluster * LQ_CAP / 4;
    w.local_spill = hm->wave_cluster * LQ_CAP / 2;
    if (getenv("GIE_WAVE_NO_LOCAL")) w.local_enter = 0;
    w.no_solo = getenv("GIE_WAVE_NO_SOLO") ? 1 : 0;

    return w;
}

}  // namespace

int gie_wave_prepare(gie_hashmap *hm)
{
    gie_locmap *lm = hm->lm;
    const LocDev &m = lm->d;
    long long bdr = 2LL * ((long long)m.X * m.Y + (long long)m.Y * m.Z + (long long)m.X * m.Z);   // reference _bdr_num
    long long cap = bdr * 4;
    if (cap < (1 << 20)) cap = 1 << 20;
    if (cap > m.N && m.N > (1 << 20)) cap = m.N;
    hm->queue_cap = (int)cap;
    for (int i = 0; i < 3; i++) {
        GIE_CUDA_CHECK(cudaMalloc(&hm->qA[i], cap * 8));
        GIE_CUDA_CHECK(cudaMalloc(&hm->qB[i], cap * 8));
        GIE_CUDA_CHECK(cudaMalloc(&hm->qC[i], cap * 8));
    }
    GIE_CUDA_CHECK(cudaMalloc(&hm->cseed_key, cap * 8));
    int per_sm = 0;
    GIE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_waves, WAVE_THREADS, 0));
    if (per_sm < 1) { gie_set_error("wave kernel does not fit on an SM"); return GIE_ERR_CUDA; }
    hm->wave_ctas = lm->num_sms;
    hm->wave_cluster = 1;
    // one CTA per SM in clusters of 16 (non-portable size) or 8 when the device can keep such a grid co-resident
    // (cooperative launch); GIE_WAVE_CLUSTER overrides the size, 1 disables clusters
    {
        int want = getenv("GIE_WAVE_CLUSTER") ? atoi(getenv("GIE_WAVE_CLUSTER")) : 16;
        for (int cs = want; cs >= 2; cs >>= 1) {
            if (cs > 8 && cudaFuncSetAttribute(k_waves, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs); cfg.blockDim = dim3(WAVE_THREADS);
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension; attr.val.clusterDim.x = cs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, (void *)k_waves, &cfg) == cudaSuccess && nclusters >= 1) {
                hm->wave_cluster = cs;
                hm->wave_ctas = std::min(nclusters * cs, (lm->num_sms / cs) * cs);
                break;
            }
            cudaGetLastError();
        }
    }
    // sized for the larger of the cluster grid and the plain one-CTA-per-SM grid the launch falls back to (gie_launch_merge)
    hm->barrier_words = (size_t)(std::max(hm->wave_ctas, lm->num_sms) + 1) * 32;
    // barrier lines, wave counters and the block-list counter share one allocation: one memset per frame clears all three
    static_assert(C_COUNT <= 64, "wave counters outgrew their slot");
    GIE_CUDA_CHECK(cudaMalloc(&hm->barrier, (hm->barrier_words + 64 + 32) * sizeof(unsigned int)));
    hm->counters = (int *)(hm->barrier + hm->barrier_words);
    hm->blk_count = hm->counters + 64;
    GIE_CUDA_CHECK(cudaMalloc(&hm->decA_dist, cap * 4));
    GIE_CUDA_CHECK(cudaMalloc(&hm->decA_coc, cap * 8));
    GIE_CUDA_CHECK(cudaMalloc(&hm->decA_pair, cap * 8));
    GIE_CUDA_CHECK(cudaMalloc(&hm->decA_flags, cap * 4));
    GIE_CUDA_CHECK(cudaMalloc(&hm->snap_id, cap * 4));
    GIE_CUDA_CHECK(cudaMalloc(&hm->blk_list, hm->tab_entries * sizeof(int)));
    GIE_CUDA_CHECK(cudaMalloc(&hm->blk_org, hm->tab_entries * sizeof(int4)));
    if (getenv("GIE_WAVE_TRACE")) {
        GIE_CUDA_CHECK(cudaMalloc(&hm->wave_trace, (size_t)TRACE_LEVELS * 10 * 8));
        GIE_CUDA_CHECK(cudaMemset(hm->wave_trace, 0, (size_t)TRACE_LEVELS * 10 * 8));
    }
    return GIE_OK;
}

// refreshes blk_list / blk_count for callers outside the merge (check.cu)
int gie_wave_list_blocks(gie_hashmap *hm)
{
    gie_locmap *lm = hm->lm;
    const int entries = (int)hm->tab_entries;
    GIE_CUDA_CHECK(cudaMemsetAsync(hm->blk_count, 0, sizeof(int), lm->stream));
    gie_launch(k_list_blocks, dim3(std::min((entries + 255) / 256, lm->num_sms * 8)), dim3(256), 0, lm->stream, lm->d, hm->d, entries, hm->blk_list, hm->blk_org, hm->blk_count);
    lm->launches++;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_merge(gie_hashmap *hm, int map_ct, int display)
{
    gie_locmap *lm = hm->lm;
    const LocDev &m = lm->d;
    hm->merge_epoch++;
    WaveDev w = make_wave_dev(hm);
    w.display = display;
    {
        StageTimer t(lm, GIE_ST_MARK_FRONTIER);
        GIE_CUDA_CHECK(cudaMemsetAsync(hm->barrier, 0, (hm->barrier_words + 64 + 32) * sizeof(unsigned int), lm->stream));   // + counters + blk_count
        const int entries = (int)hm->tab_entries;
        gie_launch(k_list_blocks, dim3(std::min((entries + 255) / 256, lm->num_sms * 8)), dim3(256), 0, lm->stream, m, hm->d, entries, hm->blk_list, hm->blk_org, hm->blk_count);
        gie_launch(k_mark_blocks, dim3(lm->num_sms * 8), dim3(256), 0, lm->stream, m, hm->d, hm->blk_org, hm->blk_count);
        gie_launch(k_frontiers, dim3(lm->num_sms * 8), dim3(256), 0, lm->stream, m, hm->d, w, map_ct, hm->blk_org, hm->blk_count);
    }
    {
        StageTimer t(lm, GIE_ST_WAVES);
        LocDev md = m; HashDev hd = hm->d; int ct = map_ct;
        void *args[] = { &md, &hd, &w, &ct };
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(hm->wave_ctas); cfg.blockDim = dim3(WAVE_THREADS); cfg.stream = lm->stream;
        cudaLaunchAttribute attrs[2];
        attrs[0].id = cudaLaunchAttributeCooperative; attrs[0].val.cooperative = 1;
        attrs[1].id = cudaLaunchAttributeClusterDimension;
        attrs[1].val.clusterDim.x = hm->wave_cluster; attrs[1].val.clusterDim.y = 1; attrs[1].val.clusterDim.z = 1;
        cfg.attrs = attrs; cfg.numAttrs = hm->wave_cluster > 1 ? 2 : 1;
        cudaError_t le = cudaLaunchKernelExC(&cfg, (void *)k_waves, args);
        if (le != cudaSuccess && hm->wave_cluster > 1) {
            // cooperative + cluster launch refused by this driver: fall back to single-CTA "clusters" for good
            cudaGetLastError();
            hm->wave_cluster = 1;
            hm->wave_ctas = lm->num_sms;
            w = make_wave_dev(hm);
            w.display = display;
            cfg.gridDim = dim3(hm->wave_ctas); cfg.numAttrs = 1;
            le = cudaLaunchKernelExC(&cfg, (void *)k_waves, args);
        }
        GIE_CUDA_CHECK(le);
    }
    {
        StageTimer t(lm, GIE_ST_COMMIT);
        gie_launch(k_commit, dim3(lm->num_sms * 8), dim3(256), 0, lm->stream, m, hm->d, display, hm->blk_org, hm->blk_count, hm->counters, hm->stats_host);
    }
    lm->launches += 5;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}
