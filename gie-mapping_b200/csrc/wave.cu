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
// is schedule dependent here; the deterministic rules D1-D4 are listed in DESIGN.md §5 and are restated
// identically by the CPU oracle.
#include "engine.h"
#include <cooperative_groups.h>

namespace {

// counters layout (ints)
enum { C_A0 = 0, C_B0 = 3, C_C0 = 6, C_LEVA = 10, C_LEVB = 11, C_LEVC = 12, C_FA = 13, C_FB = 14, C_FC = 15,
       C_FB_AFTER_A = 16, C_FC_AFTER_B = 17, C_COUNT = 32 };

struct WaveDev {
    unsigned long long *qA[3];
    unsigned long long *qB[3];
    int32_t *qC[3];
    unsigned long long *cseed_key;
    int cap;
    int *cnt;
    unsigned int *barrier;
    int32_t *dec_dist;
    unsigned long long *dec_coc;
    unsigned long long *dec_pair;
    int32_t *dec_flags;
    uint32_t *snap_id;
};

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
// MarkLimitedObserve (unify_helper.cuh:201-273).  Each thread owns VEC consecutive voxels along x (vector loads/stores),
// grid-stride over a grid sized to the SM count.  D4: UNKNOWN voxels get their batch values instead of stale memory.
template <int VEC>
__global__ void __launch_bounds__(256) k_mark(LocDev m, HashDev h)
{
    const int nq = m.N / VEC;
    const int mw = m.max_width;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const int id0 = q * VEC;
        const int x0 = id0 % m.X, yz = id0 / m.X;
        const int y = yz % m.Y, z = yz / m.Y;
        int8_t type[VEC]; int dist[VEC], cocv[VEC];
        unsigned long long pr[VEC];
        if (VEC == 4) {
            char4 t4 = *reinterpret_cast<const char4 *>(m.glb_type + id0);
            int4 a4 = *reinterpret_cast<const int4 *>(m.aux + id0);
            int4 c4 = *reinterpret_cast<const int4 *>(m.coc_aux + id0);
            type[0] = t4.x; type[1] = t4.y; type[2] = t4.z; type[3] = t4.w;
            dist[0] = a4.x; dist[1] = a4.y; dist[2] = a4.z; dist[3] = a4.w;
            cocv[0] = c4.x; cocv[1] = c4.y; cocv[2] = c4.z; cocv[3] = c4.w;
        } else { type[0] = m.glb_type[id0]; dist[0] = m.aux[id0]; cocv[0] = m.coc_aux[id0]; }
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            const int3 c = make_int3(x0 + k, y, z);
            int3 coc_new = gie_id2wr((uint32_t)cocv[k]);   // same 11/11/10 packing, local coords
            const int dist_new = dist[k];
            int aux = dist_new;
            uint32_t pid = 0;
            int pdist = 0;
            const bool see_nothing = coc_new.x > mw || coc_new.y > mw || coc_new.z > mw;   // invalid_coc_buf, voxmap_utils.cuh:174-179
            if (see_nothing) { pdist = GIE_EMPTY_VALUE; pid = 0xffffffffu; aux = GIE_EMPTY_VALUE; }
            if (type[k] != GIE_VOX_UNKNOWN) {
                int3 glb = c + m.pvt;
                int blk = gie_block_of(h, glb);
                if (blk >= 0) {   // a known voxel always has a block
                    size_t vi = (size_t)blk * 512 + gie_vox_in_block(glb);
                    int dist_old = h.dist_sq[vi];
                    int3 coc_buf_old = gie_unpack_coc(h.coc_glb[vi]) - m.pvt;
                    if (dist_new > dist_old && !gie_inside_loc(m, coc_buf_old)) { coc_new = coc_buf_old; aux = dist_old; }
                }
            }
            int3 wr = coc_new + m.pvt - m.upvt;
            if (!gie_inside_wr(wr)) { pdist = GIE_EMPTY_VALUE; aux = GIE_EMPTY_VALUE; if (!see_nothing) pid = GIE_INVALID_ID_STALE; }
            else { pdist = aux; pid = gie_wr2id(wr); }
            pr[k] = gie_mk_pair(pdist, pid);
            if (aux != dist_new) m.aux[id0 + k] = aux;
        }
        if (VEC == 4) {
            ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(m.pair + id0);
            dst[0] = make_ulonglong2(pr[0], pr[1]);
            dst[1] = make_ulonglong2(pr[2], pr[3]);
        } else m.pair[id0] = pr[0];
    }
}

// obtainFrontiers (unify_helper.cuh:275-446).  pair[] is read-only here: a lowered own pair (frontier C seed) is
// deferred into cseed_key and applied by the wave kernel, which removes the reference's _g/_coc_idx backup arrays.
template <int VEC>
__global__ void __launch_bounds__(256) k_frontiers(LocDev m, HashDev h, WaveDev w, int map_ct)
{
    const int nq = m.N / VEC;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const int id0 = q * VEC;
        const int x0 = id0 % m.X, yz = id0 / m.X;
        const int y = yz % m.Y, z = yz / m.Y;
        int8_t types[VEC]; int wls[VEC];
        if (VEC == 4) {
            char4 t4 = *reinterpret_cast<const char4 *>(m.glb_type + id0);
            types[0] = t4.x; types[1] = t4.y; types[2] = t4.z; types[3] = t4.w;
        } else types[0] = m.glb_type[id0];
#pragma unroll
        for (int kk = 0; kk < VEC; kk++) {
            const int3 c = make_int3(x0 + kk, y, z);
            const int id = id0 + kk;
            const int8_t type = types[kk];
            int wl = GIE_EMPTY_VALUE;
            if (type != GIE_VOX_UNKNOWN) {
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
                            int3 nwr = gie_id2wr(gie_pair_id(m.pair[nid]));
                            int3 ncb = nwr + m.upvt - m.pvt;
                            if (!gie_inside_loc(m, ncb) && gie_inside_wr(nwr)) {
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
                        wl = 1;
                        int i = atomicAdd(&w.cnt[C_C0], 1);
                        if (i < w.cap) { w.qC[0][i] = id; w.cseed_key[i] = new_key; }
                        else atomicOr(h.status, GIE_DEV_ERR_QUEUE_OVERFLOW);
                    }
                    if (type == GIE_VOX_FREE && nbr_unknown) m.glb_type[id] = GIE_VOX_FNT;
                }
            }

            wls[kk] = wl;
        }
        if (VEC == 4) *reinterpret_cast<int4 *>(m.wave_layer + id0) = make_int4(wls[0], wls[1], wls[2], wls[3]);
        else m.wave_layer[id0] = wls[0];
    }
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &gen)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        gen++;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int target = gen * gridDim.x;
        while (*((volatile unsigned int *)bar) < target) { __nanosleep(20); }
        __threadfence();
    }
    __syncthreads();
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
            if (gie_inside_loc(m, ncb) && m.aux[gie_lidx(m, ncb)] != 0) {
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
    for (int i = tid; i < n; i += nthreads) {
        uint32_t sid = w.snap_id[i];
        if (sid == 0xffffffffu) continue;
        int3 cg = unpack_glb(__ldcg(&cur[i]));
        int3 coc = gie_id2wr(sid) + m.upvt;
        for (int d = 0; d < 6; d++) {
            int3 ng = cg + DIRS6[d];
            int3 nb = ng - m.pvt;
            int cand = sqd3(coc, ng);
            unsigned long long key = gie_mk_pair(cand, sid);
            if (!gie_inside_loc(m, nb)) {
                size_t ni;
                if (!vox_ref(h, ng, ni)) continue;
                if (h.vox_type[ni] == GIE_VOX_UNKNOWN) continue;
                if (gie_invalid_coc_glb(gie_unpack_coc(__ldcg(&h.coc_glb[ni])))) continue;
                unsigned long long old = atomicMin(&h.pair[ni], key);
                if (key < old) {
                    int color = atomicExch(&h.wave_layer[ni], gray);
                    if (color == gray) continue;
                    h.update_ct[ni] = map_ct;
                    q_push(next, next_cnt, w.cap, pack_glb(ng), h.status);
                }
            } else {
                int nid = gie_lidx(m, nb);
                if (m.aux[nid] > cand) {
                    atomicMin(&m.pair[nid], key);
                    if (atomicExch(&m.wave_layer[nid], 1) != 1)
                        q_push(w.qC[0], &w.cnt[C_C0], w.cap, (int32_t)nid, h.status);
                }
            }
        }
    }
}

// wave C: lower_inside (wave_core.cuh:353-393)
__device__ void waveC_phase1(const LocDev &m, const WaveDev &w, const int32_t *cur, int n, int tid, int nthreads)
{
    for (int i = tid; i < n; i += nthreads) {
        int id = __ldcg(&cur[i]);
        m.wave_layer[id] = GIE_WL_BLACK;
        w.snap_id[i] = gie_pair_id(__ldcg(&m.pair[id]));
    }
}
__device__ void waveC_phase2(const LocDev &m, const HashDev &h, const WaveDev &w, const int32_t *cur, int n, int32_t *next,
                             int *next_cnt, int gray, int tid, int nthreads)
{
    const int XY = m.X * m.Y;
    for (int i = tid; i < n; i += nthreads) {
        int id = __ldcg(&cur[i]);
        uint32_t sid = w.snap_id[i];
        int3 cb = make_int3(id % m.X, (id / m.X) % m.Y, id / XY);
        int3 coc_buf = gie_id2wr(sid) + m.upvt - m.pvt;
        for (int d = 0; d < 6; d++) {
            int3 nb = cb + DIRS6[d];
            if (!gie_inside_loc(m, nb)) continue;
            int nid = gie_lidx(m, nb);
            unsigned long long key = gie_mk_pair(sqd3(coc_buf, nb), sid);
            unsigned long long old = atomicMin(&m.pair[nid], key);
            if (key < old) {
                if (atomicExch(&m.wave_layer[nid], gray) == gray) continue;
                q_push(next, next_cnt, w.cap, (int32_t)nid, h.status);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_waves(LocDev m, HashDev h, WaveDev w, int map_ct)
{
    unsigned int gen = 0;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    volatile int *cnt = w.cnt;
    // phase 0: apply the deferred frontier-C seeds
    {
        int n = min(cnt[C_C0], w.cap);
        for (int i = tid; i < n; i += nthreads) m.pair[__ldcg(&w.qC[0][i])] = __ldcg(&w.cseed_key[i]);
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
            if (tid == 0) w.cnt[C_C0 + zi] = 0;
            int gray = (level & 1) ? GIE_WL_GRAY1 : GIE_WL_GRAY0;
            waveC_phase1(m, w, w.qC[ci], n, tid, nthreads);
            grid_barrier(w.barrier, gen);
            waveC_phase2(m, h, w, w.qC[ci], n, w.qC[ni], &w.cnt[C_C0 + ni], gray, tid, nthreads);
            grid_barrier(w.barrier, gen);
        }
        if (tid == 0) w.cnt[C_LEVC] = level;
    }
}

// UpdateHashBatch (unify_helper.cuh:448-523)
template <int VEC>
__global__ void __launch_bounds__(256) k_commit(LocDev m, HashDev h)
{
    const int nq = m.N / VEC;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const int id0 = q * VEC;
        int8_t types[VEC];
        if (VEC == 4) {
            char4 t4 = *reinterpret_cast<const char4 *>(m.glb_type + id0);
            types[0] = t4.x; types[1] = t4.y; types[2] = t4.z; types[3] = t4.w;
            if ((t4.x | t4.y | t4.z | t4.w) == 0) continue;
        } else types[0] = m.glb_type[id0];
        const int x0 = id0 % m.X, yz = id0 / m.X;
        const int y = yz % m.Y, z = yz / m.Y;
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            const int8_t type = types[k];
            if (type == GIE_VOX_UNKNOWN) continue;
            const int id = id0 + k;
            unsigned long long pr = m.pair[id];
            int dist = gie_pair_dist(pr);
            uint32_t pid = gie_pair_id(pr);
            if (dist == GIE_EMPTY_VALUE) {
                if (pid == 0xffffffffu) m.edt[id] = (float)m.max_loc_dist_sq;
                continue;
            }
            int3 glb = make_int3(x0 + k, y, z) + m.pvt;
            int blk = gie_block_of(h, glb);
            if (blk < 0) continue;
            size_t vi = (size_t)blk * 512 + gie_vox_in_block(glb);
            h.coc_glb[vi] = gie_pack_coc(gie_id2wr(pid) + m.upvt);
            h.dist_sq[vi] = dist;
            m.edt[id] = sqrtf((float)dist);
            h.pair[vi] = pr;
            if (type == GIE_VOX_FNT) h.vox_type[vi] = GIE_VOX_FNT;
        }
    }
}

__global__ void k_wave_stats(WaveDev w, long long *out)
{
    out[0] = w.cnt[C_FA]; out[1] = w.cnt[C_FB]; out[2] = w.cnt[C_FC]; out[3] = w.cnt[C_LEVA]; out[4] = w.cnt[C_LEVB];
    out[5] = w.cnt[C_LEVC]; out[6] = w.cnt[C_FB_AFTER_A]; out[7] = w.cnt[C_FC_AFTER_B];
}

WaveDev make_wave_dev(gie_hashmap *hm)
{
    WaveDev w{};
    for (int i = 0; i < 3; i++) { w.qA[i] = hm->qA[i]; w.qB[i] = hm->qB[i]; w.qC[i] = hm->qC[i]; }
    w.cseed_key = hm->cseed_key; w.cap = hm->queue_cap; w.cnt = hm->counters; w.barrier = hm->barrier;
    w.dec_dist = hm->decA_dist; w.dec_coc = hm->decA_coc; w.dec_pair = hm->decA_pair; w.dec_flags = hm->decA_flags;
    w.snap_id = hm->snap_id;
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
        GIE_CUDA_CHECK(cudaMalloc(&hm->qC[i], cap * 4));
    }
    GIE_CUDA_CHECK(cudaMalloc(&hm->cseed_key, cap * 8));
    GIE_CUDA_CHECK(cudaMalloc(&hm->counters, C_COUNT * sizeof(int)));
    GIE_CUDA_CHECK(cudaMalloc(&hm->barrier, sizeof(unsigned int)));
    GIE_CUDA_CHECK(cudaMalloc(&hm->decA_dist, cap * 4));
    GIE_CUDA_CHECK(cudaMalloc(&hm->decA_coc, cap * 8));
    GIE_CUDA_CHECK(cudaMalloc(&hm->decA_pair, cap * 8));
    GIE_CUDA_CHECK(cudaMalloc(&hm->decA_flags, cap * 4));
    GIE_CUDA_CHECK(cudaMalloc(&hm->snap_id, cap * 4));
    int per_sm = 0;
    GIE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_waves, 256, 0));
    if (per_sm < 1) { gie_set_error("wave kernel does not fit on an SM"); return GIE_ERR_CUDA; }
    if (per_sm > 2) per_sm = 2;
    hm->wave_ctas = per_sm * lm->num_sms;
    return GIE_OK;
}

int gie_launch_merge(gie_hashmap *hm, int map_ct)
{
    gie_locmap *lm = hm->lm;
    const LocDev &m = lm->d;
    WaveDev w = make_wave_dev(hm);
    const int vec = (m.X % 4 == 0) ? 4 : 1;
    long long groups = (long long)m.N / vec;
    long long want = (groups + 255) / 256, cap = (long long)lm->num_sms * 16;
    const int grid = (int)(want < cap ? want : cap);
    {
        StageTimer t(lm, GIE_ST_MARK_FRONTIER);
        GIE_CUDA_CHECK(cudaMemsetAsync(hm->counters, 0, C_COUNT * sizeof(int), lm->stream));
        GIE_CUDA_CHECK(cudaMemsetAsync(hm->barrier, 0, sizeof(unsigned int), lm->stream));
        if (vec == 4) {
            k_mark<4><<<grid, 256, 0, lm->stream>>>(m, hm->d);
            k_frontiers<4><<<grid, 256, 0, lm->stream>>>(m, hm->d, w, map_ct);
        } else {
            k_mark<1><<<grid, 256, 0, lm->stream>>>(m, hm->d);
            k_frontiers<1><<<grid, 256, 0, lm->stream>>>(m, hm->d, w, map_ct);
        }
    }
    {
        StageTimer t(lm, GIE_ST_WAVES);
        LocDev md = m; HashDev hd = hm->d; int ct = map_ct;
        void *args[] = { &md, &hd, &w, &ct };
        GIE_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_waves, dim3(hm->wave_ctas), dim3(256), args, 0, lm->stream));
    }
    {
        StageTimer t(lm, GIE_ST_COMMIT);
        if (vec == 4) k_commit<4><<<grid, 256, 0, lm->stream>>>(m, hm->d);
        else k_commit<1><<<grid, 256, 0, lm->stream>>>(m, hm->d);
    }
    k_wave_stats<<<1, 1, 0, lm->stream>>>(w, hm->stats_host);
    lm->launches += 5;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}
