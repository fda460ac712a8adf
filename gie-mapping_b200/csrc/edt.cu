// edt.cu — exact batch squared-EDT (+ closest-obstacle coordinate) of the dense local volume.
//
// Replaces EDT_OCC::batchEDTUpdate (reference src/kernel/edt/local_edt.cu:7-28): EDTphase1/2/3
// (src/kernel/edt/local_edt_core.h:14-193, f/sep at include/map_structure/local_batch.h:494-520) and the six cuTT
// transposes (include/cutt/cutt.h:57-101, plans at src/volumetric_mapper.cpp:344-373).
//
// Contract reproduced: exact, un-truncated squared EDT over OCCUPIED voxels; closest obstacle
// chosen as: along y nearest with ties -> larger y; along x argmin_i (x-i)^2 + g1(i)^2 ties -> smallest i; along z
// argmin_k (z-k)^2 + g2(k) ties -> smallest k.  Output _aux = dist_sq, _coc_idx_aux = x | y<<11 | z<<22 (local).
//
// Mechanism (no transposes, no global s/t/g arrays):
//   y bits      : OCCUPIED bits packed along y (low word of ytab[z][y/32][x]) are kept in step by the OGM merge (hashmap.cu);
//                 k_edt_ybits rebuilds them from glb_type only when the array was written from outside.
//   k_edt_ycols : one CTA per slice z the merge flagged, one thread per column x.  Links every word to the nearest set bit
//                 below / above it (0.25 B/voxel); the y pass is never materialised.  Also compacts the columns of the slice
//                 that hold an obstacle.  A column without one is the reference's "_max_width" sentinel for EVERY row of
//                 the slice and can never win the x sweep while a real column exists ((X+Y+Z)^2 exceeds any real candidate),
//                 so the x sweep only visits real columns, and slices without any obstacle are skipped altogether.
//   k_edt_xsweep: banded lower envelope.  Item = (real slice, 16 rows); thread = (band of real columns, row).  Forward per
//                 band with the reference's sequential push, bands merged pairwise by searching the cut between two
//                 envelopes, backward per x range through a 16 x 8 tile so that global stores are x-contiguous.
//   k_edt_zsweep: one warp = 32 consecutive x of one row y; the sequential scan along z over the real slices only, naturally
//                 coalesced; an entry carries the closest-obstacle word it will store; streaming stores of aux / coc_aux.
// Persistent CTAs (a multiple of the SM count) pull work items from an atomic counter.
#include "engine.h"
#include <algorithm>

namespace {

constexpr int INV_Y = 2045;   // INVALID_LOC_COC.y, local_batch.h:59

// y pass, step 1: OCCUPIED bits of 32 consecutive y per (z, wy, x) -> low word of ytab.  One thread per VEC adjacent words
// (x fastest): 32 coalesced loads of VEC bytes in flight per thread, 128 bytes per warp instruction at VEC = 4.
template <int VEC>
__global__ void __launch_bounds__(128) k_edt_ybits(LocDev m, unsigned long long *__restrict__ ytab, int WY)
{
    gie_pdl_sync();
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * VEC, wy = blockIdx.y, z = blockIdx.z;
    if (x >= m.X) return;
    const int ybase = wy * 32, n = min(32, m.Y - ybase);
    const int8_t *col = m.glb_type + ((size_t)z * m.Y + ybase) * m.X + x;
    uint32_t w[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) w[v] = 0;
#pragma unroll 8
    for (int b = 0; b < n; b++) {
        if (VEC == 4) {
            const char4 t = *reinterpret_cast<const char4 *>(col + (size_t)b * m.X);   // X % 4 == 0: aligned
            w[0] |= (uint32_t)(t.x == GIE_VOX_OCCUPIED) << b; w[1] |= (uint32_t)(t.y == GIE_VOX_OCCUPIED) << b;
            w[2] |= (uint32_t)(t.z == GIE_VOX_OCCUPIED) << b; w[3] |= (uint32_t)(t.w == GIE_VOX_OCCUPIED) << b;
        } else w[0] |= (uint32_t)(col[(size_t)b * m.X] == GIE_VOX_OCCUPIED) << b;
    }
    unsigned long long *out = ytab + ((size_t)z * WY + wy) * m.X + x;
#pragma unroll
    for (int v = 0; v < VEC; v++) out[v] = w[v];
}

// y pass, step 1, per frame: nothing in this file.  Every OCCUPIED voxel of the volume passes through the OGM merge of the
// frame (hashmap.cu), which holds each visited block's types in shared memory: k_merge_ogm sets the bits of the block's columns
// and k_clear_prev_blocks clears those of the previous merge's blocks (134 MB of glb_type at 512^3, 1 GB at 1024^3, are not read
// back).  k_edt_ybits above remains for volumes whose glb_type was written from outside and for the first frame's state.

// y pass, step 2: per column (z, x) link every word to the nearest set bit below / above it, and compact the columns of the
// slice that hold an obstacle.  One CTA per slice, one thread per column; touches only ytab (0.25 B/voxel).
__global__ void __launch_bounds__(1024) k_edt_ycols(LocDev m, unsigned long long *__restrict__ ytab, int WY,
                                                    int *__restrict__ col_list, int *__restrict__ n_cols, const int *__restrict__ slice_has)
{
    gie_pdl_sync();
    __shared__ int warp_cnt[32];
    const int x = threadIdx.x, z = blockIdx.x;
    // the OGM merge flags the slices it set a bit in: nine slices in ten hold none and their 16 x X words need not be read
    if (slice_has && __ldcg(&slice_has[z]) == 0) {
        if (x == 0) n_cols[z] = 0;
        return;
    }
    const int lane = x & 31, wid = x >> 5;
    bool any = false;
    if (x < m.X) {
        unsigned long long *out = ytab + (size_t)z * WY * m.X + x;
        uint32_t any_bits = 0;
        for (int wy = 0; wy < WY; wy++) any_bits |= (uint32_t)out[(size_t)wy * m.X];
        any = any_bits != 0;
        if (any) {   // a column without obstacles is never visited by the x sweep
            int lo_prev = 0xffff;
            for (int wy = 0; wy < WY; wy++) {
                uint32_t w = (uint32_t)out[(size_t)wy * m.X];
                out[(size_t)wy * m.X] = (unsigned long long)w | ((unsigned long long)lo_prev << 32);
                if (w) lo_prev = wy * 32 + 31 - __clz(w);
            }
            int hi_next = 0xffff;
            for (int wy = WY - 1; wy >= 0; wy--) {
                unsigned long long e = out[(size_t)wy * m.X];
                out[(size_t)wy * m.X] = e | ((unsigned long long)hi_next << 48);
                uint32_t w = (uint32_t)e;
                if (w) hi_next = wy * 32 + __ffs(w) - 1;
            }
        }
    }
    // ordered compaction of the real columns of this slice
    unsigned bal = __ballot_sync(0xffffffffu, any);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    int before = 0, total = 0;
    for (int i = 0; i < nw; i++) { int c = warp_cnt[i]; if (i < wid) before += c; total += c; }
    if (any) col_list[(size_t)z * m.X + before + __popc(bal & ((1u << lane) - 1))] = x;
    if (x == 0) n_cols[z] = total;
}

// ordered list of the slices that hold at least one obstacle
__global__ void __launch_bounds__(1024) k_edt_slices(int Z, const int *__restrict__ n_cols, int *__restrict__ slice_list,
                                                     int *__restrict__ n_slices, int *__restrict__ zero4)
{
    gie_pdl_sync();
    if (zero4 && threadIdx.x < 4) zero4[threadIdx.x] = 0;   // the sweeps' work counters (saves a memset node per frame)
    __shared__ int warp_cnt[32];
    const int z = threadIdx.x, lane = z & 31, wid = z >> 5;
    bool any = z < Z && n_cols[z] > 0;
    unsigned bal = __ballot_sync(0xffffffffu, any);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    int before = 0, total = 0;
    for (int i = 0; i < nw; i++) { int c = warp_cnt[i]; if (i < wid) before += c; total += c; }
    if (any) slice_list[before + __popc(bal & ((1u << lane) - 1))] = z;
    if (z == 0) *n_slices = total;
}

// exact floor(num/den) for 0 <= num < 2^24, 0 < den < 2^12 (num >= 0 is guaranteed)
__device__ __forceinline__ int floor_div(int num, int den)
{
    int q = (int)__fdividef((float)num, (float)den);
    int rem = num - q * den;
    if (rem < 0) q--; else if (rem >= den) q++;
    return q;
}

// Envelope stack of one lane.  Entry = (cw, h | t<<21) with cw = cocx | cocy<<11 | s<<22, the closest-obstacle word the
// backward pass stores as it is (h < 2^21, t < 2^10).  The top lives in registers; entries [base, q] live in a
// shared-memory ring (conflict free: bank == lane), entries below `base` in the per-warp global scratch.
struct Top { int s, t, h, cw; };   // site, start, height, closest-obstacle word of the site (carries s in its top 10 bits)
template <int RING>
struct LaneStack {
    int *sh;          // smem: sh[slot*32] = cw words, sb = sh + RING*32 the (h, t) words (lane offset applied)
    int *sb;
    uint2 *g;         // global scratch, lane offset applied, stride 32
    int base;
    __device__ __forceinline__ void put(int q, const Top &e)
    {
        if (q < base) base = q;
        else if (q - base >= RING) {
            int sl = (base % RING) * 32;
            g[base * 32] = make_uint2((uint32_t)sh[sl], (uint32_t)sb[sl]);
            base++;
        }
        int sl = (q % RING) * 32;
        sh[sl] = e.cw;
        sb[sl] = e.h | (e.t << 21);
    }
    __device__ __forceinline__ Top get(int q)
    {
        if (q < base) {   // refill half a ring with independent loads
            int nb = max(0, q - RING / 2 + 1);
            for (int i = nb; i <= q; i++) {
                uint2 v = g[i * 32];
                int sl = (i % RING) * 32;
                sh[sl] = (int)v.x; sb[sl] = (int)v.y;
            }
            base = nb;
        }
        int sl = (q % RING) * 32;
        const uint32_t b = (uint32_t)sb[sl];
        Top e;
        e.cw = sh[sl]; e.s = (int)((uint32_t)e.cw >> 22); e.h = (int)(b & 0x1fffff); e.t = (int)(b >> 21);
        return e;
    }
};

// one step of the lower-envelope construction (EDTphase2/3 forward loops, local_edt_core.h:93-115 / :146-168)
template <int RING>
__device__ __forceinline__ void envelope_push(int u, int h_u, int cw_u, int L, int &q, Top &top, LaneStack<RING> &st)
{
    while (q >= 0) {
        int a = top.t - top.s, b = top.t - u;
        if (a * a + top.h > b * b + h_u) {
            q--;
            if (q >= 0) top = st.get(q);
        } else break;
    }
    if (q < 0) {
        q = 0;
        top.s = u; top.t = 0; top.h = h_u; top.cw = cw_u;
        st.put(0, top);
    } else {
        int num = u * u - top.s * top.s + h_u - top.h;
        int den = 2 * (u - top.s);
        int w = 1 + floor_div(num, den);
        if (w < L) {
            q++;
            top.s = u; top.t = w; top.h = h_u; top.cw = cw_u;
            st.put(q, top);
        }
    }
}

// ---- banded envelope machinery shared by the x sweep and the dense z sweep ---------------------------------------------
// The candidates of one scan line (real columns of a row / real slices of a z column) are cut into NB bands of at most CAP
// candidates; thread = (band b, line r), R lines per work item.
//   1. forward: every (band, line) thread builds the lower envelope of ITS band with the reference's sequential push, the
//      whole stack in shared memory ([band][entry][line], conflict free).
//   2. merge: parabolas of equal curvature cross exactly once, so the envelope of (left bands) U (right bands) is a prefix
//      of the left envelope followed by a suffix of the right one.  Re-pushing the right envelope onto the left with the
//      sequential algorithm would pop some top entries of the left, drop some first entries of the right, and then copy the
//      rest unchanged — so only the cut is searched (a two-pointer loop of typically 1-3 steps) and recorded per band as
//      (lo, hi, tfirst): the surviving entry range of the band's stack and the start of its first survivor.  Bands are merged
//      pairwise in a tree (log2 NB rounds).  Same comparisons, same integer divisions, same tie behaviour as the push loop.
//   3. backward: the scan line is cut into NB ranges; every (range, line) thread finds the entry that covers the end of its
//      range in the composite envelope (EnvCursor) and walks down.
// A serial scan of a 512-long line is ~1000 dependent steps; here the dependent chain of an item is CAP + range (+ merges)
// ~ 70 steps, and a line with few candidates still keeps NB x 32 threads busy.
struct BandCfg { int NB, CAP, BW; };   // bands, stack capacity per band, output range per band

__device__ __forceinline__ int bd_meta_pack(int lo, int hi1, int tf) { return lo | (hi1 << 8) | (tf << 16); }
#define BD_LO(mw) ((mw) & 0xff)
#define BD_HI1(mw) (((mw) >> 8) & 0xff)      // hi + 1; the range [lo, hi] is empty when hi1 <= lo
#define BD_TF(mw) ((mw) >> 16)
#define BD_EMPTY(mw) (BD_HI1(mw) <= BD_LO(mw))

// Shared-memory view of the stacks of one work item for the thread's line r.  Entry = (h, s | t << 10 [| cy << 20]).
template <int R>
struct BandStacks {
    int *stH, *stB, *meta;
    int sk, si, r;           // element (band k, entry i) of line r lives at k * sk + i * si + r
    __device__ __forceinline__ int &M(int k) const { return meta[k * R + r]; }
    __device__ __forceinline__ int H(int k, int i) const { return stH[k * sk + i * si + r]; }
    __device__ __forceinline__ int B(int k, int i) const { return stB[k * sk + i * si + r]; }
};

// one push of the sequential algorithm (EDTphase2/3 forward loops, local_edt_core.h:93-115 / :146-168) onto the stack of band
// `b`; (ts, tt, th) = top entry in registers, q = its index.  `extra` is or-ed into the packed word (the x sweep keeps cy there).
template <int R>
__device__ __forceinline__ void band_push(const BandStacks<R> &S, int b, int u, int h_u, int extra, int L, int &q, int &ts, int &tt, int &th)
{
    int *myH = S.stH + b * S.sk + S.r, *myB = S.stB + b * S.sk + S.r;
    const int si = S.si;
    while (q >= 0) {
        const int a = tt - ts, c = tt - u;
        if (a * a + th > c * c + h_u) {
            q--;
            if (q >= 0) { const int bb = myB[q * si]; th = myH[q * si]; ts = bb & 0x3ff; tt = (bb >> 10) & 0x3ff; }
        } else break;
    }
    int w = 0;
    if (q >= 0) w = 1 + floor_div(u * u - ts * ts + h_u - th, 2 * (u - ts));
    if (w < L) {
        q++;
        ts = u; tt = w; th = h_u;
        myH[q * si] = h_u;
        myB[q * si] = u | (w << 10) | extra;
    }
}

// merge of the composite envelope of bands [b, b + stride) with that of bands [b + stride, b + 2 stride): see above
template <int R>
__device__ __forceinline__ void band_merge(const BandStacks<R> &S, int b, int stride, int NB, int L)
{
    const int l_end = b + stride, r_end = min(b + 2 * stride, NB);
    int kl = l_end - 1, kr = l_end;
    int ml = 0, mr = 0;
    while (kl >= b && (ml = S.M(kl), BD_EMPTY(ml))) kl--;
    while (kr < r_end && (mr = S.M(kr), BD_EMPTY(mr))) kr++;
    if (kl < b || kr >= r_end) return;            // one side is empty: the other stands as it is (its first entry starts at 0)
    int llo = BD_LO(ml), lhi = BD_HI1(ml) - 1, ltf = BD_TF(ml);
    int rlo = BD_LO(mr), rhi = BD_HI1(mr) - 1;
    for (;;) {
        const int lb = S.B(kl, lhi), lh = S.H(kl, lhi);
        const int rb = S.B(kr, rlo), rh = S.H(kr, rlo);
        const int ls = lb & 0x3ff, lt = (lhi == llo) ? ltf : ((lb >> 10) & 0x3ff);
        const int rs = rb & 0x3ff;
        const int a = lt - ls, c = lt - rs;
        if (a * a + lh > c * c + rh) {            // the left top is dominated from its own start on: pop it
            if (lhi > llo) { lhi--; continue; }
            S.M(kl) = bd_meta_pack(0, 0, 0);
            kl--;
            while (kl >= b && (ml = S.M(kl), BD_EMPTY(ml))) kl--;
            if (kl < b) { S.M(kr) = bd_meta_pack(rlo, rhi + 1, 0); return; }   // nothing left on the left: starts at 0
            llo = BD_LO(ml); lhi = BD_HI1(ml) - 1; ltf = BD_TF(ml);
            continue;
        }
        const int w = 1 + floor_div(rs * rs - ls * ls + rh - lh, 2 * (rs - ls));
        // the first entry of the right envelope ends where its successor starts
        int rend = L;
        if (rlo < rhi) rend = (S.B(kr, rlo + 1) >> 10) & 0x3ff;
        else {
            int kn = kr + 1, mn = 0;
            while (kn < r_end && (mn = S.M(kn), BD_EMPTY(mn))) kn++;
            if (kn < r_end) rend = BD_TF(mn);
        }
        if (w >= rend) {                          // squeezed out between the left top and its own successor: drop it
            if (rlo < rhi) { rlo++; continue; }
            S.M(kr) = bd_meta_pack(0, 0, 0);
            kr++;
            while (kr < r_end && (mr = S.M(kr), BD_EMPTY(mr))) kr++;
            if (kr >= r_end) { S.M(kl) = bd_meta_pack(llo, lhi + 1, ltf); return; }
            rlo = BD_LO(mr); rhi = BD_HI1(mr) - 1;
            continue;
        }
        S.M(kl) = bd_meta_pack(llo, lhi + 1, ltf);
        S.M(kr) = bd_meta_pack(rlo, rhi + 1, w);
        return;
    }
}

// position in the composite envelope during the backward pass
template <int R>
struct EnvCursor {
    int k, i, lo, tf;        // band, entry, first surviving entry of the band and its start
    int es, eh, eb, et;      // current entry: site, height, packed word, start
    __device__ __forceinline__ void load(const BandStacks<R> &S)
    {
        eb = S.B(k, i); eh = S.H(k, i);
        es = eb & 0x3ff;
        et = (i == lo) ? tf : ((eb >> 10) & 0x3ff);
    }
    // the entry that covers position p (the last one whose start is <= p)
    __device__ __forceinline__ void seek(const BandStacks<R> &S, int NB, int p)
    {
        int mw = 0;
        for (k = NB - 1; k > 0; k--) { mw = S.M(k); if (!BD_EMPTY(mw) && BD_TF(mw) <= p) break; }
        if (k == 0) mw = S.M(0);                  // the first non-empty band starts at 0: the scan cannot fall through
        while (BD_EMPTY(mw)) { k++; mw = S.M(k); }   // (band 0 itself may be empty)
        lo = BD_LO(mw); tf = BD_TF(mw);
        i = BD_HI1(mw) - 1;
        while (i > lo && ((S.B(k, i) >> 10) & 0x3ff) > p) i--;
        load(S);
    }
    // step to the previous entry of the composite envelope (the caller guarantees there is one)
    __device__ __forceinline__ void prev(const BandStacks<R> &S)
    {
        if (i > lo) i--;
        else {
            int mw;
            do { k--; mw = S.M(k); } while (BD_EMPTY(mw));
            lo = BD_LO(mw); tf = BD_TF(mw); i = BD_HI1(mw) - 1;
        }
        load(S);
    }
};

// ---- x sweep (EDTphase2, local_edt_core.h:84-135) ---------------------------------------------------------------------
// Work item = (obstacle-bearing slice z, 16 consecutive rows y); thread = (band of real columns, row), two bands per warp.
//   0. the CTA stages the item's candidates once: column index and ytab entry of every real column of the slice, one coalesced
//      gather (two dependent global loads per thread per item instead of two per candidate inside the scan);
//   1-3. forward per band, tree of merges, backward per x range (see above).  The backward pass emits through a [16][8]
//      shared-memory tile so that global stores run along x: the in-kernel transpose that replaces cuTT's {1,0,2} permutation
//      and its inverse.
// The next item's index is fetched while the current one is processed.  (A variant with one WARP per row — lane = band, no CTA
// barrier at all — executed 2.9x the instructions: lanes of different bands diverge in every pop loop and a merge round keeps
// 16, 8, 4, 2, 1 lanes busy; 0.28 ms against 0.20 for this shape, profiles/r02_edt_experiments.md.)
constexpr int XS_RPI = 16, XS_TW = 8;
__global__ void __launch_bounds__(512)
k_edt_xsweep(LocDev m, const unsigned long long *__restrict__ ytab, int WY, const int *__restrict__ col_list,
             const int *__restrict__ n_cols, const int *__restrict__ slice_list, const int *__restrict__ n_slices,
             int32_t *__restrict__ g2, int32_t *__restrict__ cxy, BandCfg cfg, int *__restrict__ work_counter, int compact)
{
    gie_pdl_sync();
    constexpr int RPI = XS_RPI, TW = XS_TW;
    extern __shared__ int xs_smem[];
    __shared__ int s_item[2];
    const int NB = cfg.NB, CAP = cfg.CAP, BW = cfg.BW;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r = lane % RPI, b = wid * (32 / RPI) + lane / RPI;
    const int X = m.X;
    BandStacks<RPI> S;
    S.stH = xs_smem;                                     // [NB][CAP][RPI]  h
    S.stB = S.stH + NB * CAP * RPI;                      // [NB][CAP][RPI]  s | t << 10 | cy << 20
    S.meta = S.stB + NB * CAP * RPI;                     // [NB][RPI]
    S.sk = CAP * RPI; S.si = RPI; S.r = r;
    int *tile_g = S.meta + NB * RPI + b * 2 * RPI * (TW + 1), *tile_c = tile_g + RPI * (TW + 1);   // [NB][2][RPI][TW + 1]
    int *cand_u = S.meta + NB * RPI + NB * 2 * RPI * (TW + 1);                                      // [X] column of candidate j
    unsigned long long *cand_e = (unsigned long long *)(cand_u + ((X + 1) & ~1));                   // [X] its ytab entry for this row group
    const int YS0 = m.ys0, YSN = m.ysn;                  // rows of this map (a slab of a sharded volume, or all of them)
    const int RG = (YSN + RPI - 1) / RPI;                // row groups per slice
    const int n_items = __ldg(n_slices) * RG;
    if (threadIdx.x == 0) s_item[0] = atomicAdd(work_counter, 1);
    __syncthreads();
    for (int it = 0;; it ^= 1) {
        const int item = s_item[it];
        if (item >= n_items) break;
        if (threadIdx.x == 0) s_item[it ^ 1] = atomicAdd(work_counter, 1);   // visible after the barriers below
        const int zi = item / RG, rg = item - zi * RG;
        const int z = __ldg(&slice_list[zi]);
        const int y = YS0 + rg * RPI + r;
        const int wy = y >> 5, p = y & 31;               // ytab word and bit of this row (a row group never straddles a word)
        const int zp = compact ? zi : z;                 // plane of ytab / col_list: compacted to the real slices when received from a peer
        const int nc = __ldg(&n_cols[z]);
        // ---- 0. stage the candidates of this (slice, word of rows)
        {
            const unsigned long long *trow = ytab + ((size_t)zp * WY + wy) * X;
            const int *cols = col_list + (size_t)zp * X;
            for (int j = threadIdx.x; j < nc; j += blockDim.x) {
                const int u = __ldg(&cols[j]);
                cand_u[j] = u;
                cand_e[j] = __ldg(&trow[u]);
            }
        }
        __syncthreads();   // also: the previous item's stacks, meta words and tiles are no longer read
        // ---- 1. forward pass over this band's real columns
        {
            const int jb = (int)((long long)nc * b / NB), je = (int)((long long)nc * (b + 1) / NB);
            int q = -1, ts = 0, tt = 0, th = 0;
            const uint32_t lomask = 0xffffffffu >> (31 - p);
            for (int j = jb; j < je; j++) {
                const int u = cand_u[j];
                const unsigned long long e = cand_e[j];
                const uint32_t w = (uint32_t)e;
                const int lo_prev = (int)((e >> 32) & 0xffff), hi_next = (int)(e >> 48);
                const uint32_t mlo = w & lomask, mhi = w >> p;
                const int lo = mlo ? (wy * 32 + 31 - __clz(mlo)) : (lo_prev == 0xffff ? -1 : lo_prev);
                const int hi = mhi ? (y + __ffs(mhi) - 1) : (hi_next == 0xffff ? -1 : hi_next);
                int g1, cy;
                if (hi >= 0 && (lo < 0 || hi - y <= y - lo)) { g1 = hi - y; cy = hi; }   // ties -> larger y
                else { g1 = y - lo; cy = lo; }                                          // a real column always has lo or hi
                band_push<RPI>(S, b, u, g1 * g1, cy << 20, X, q, ts, tt, th);
            }
            S.M(b) = bd_meta_pack(0, q + 1, 0);
        }
        // ---- 2. pairwise merges: (0 <- 1), (2 <- 3), ... then (0..1 <- 2..3), ... until the row's envelope is described.
        // The two bands of the first round live in one warp.
        __syncwarp();
        if ((b & 1) == 0 && b + 1 < NB) band_merge<RPI>(S, b, 1, NB, X);
        for (int stride = 2; stride < NB; stride <<= 1) {
            __syncthreads();
            if ((b & (2 * stride - 1)) == 0 && b + stride < NB) band_merge<RPI>(S, b, stride, NB, X);
        }
        __syncthreads();
        // ---- 3. backward pass (local_edt_core.h:116-134) over this band's x range, emitted through a RPI x TW tile
        {
            const int x_lo = b * BW, x_hi = min(X, x_lo + BW) - 1;
            const bool act = x_lo <= x_hi;
            EnvCursor<RPI> cur{};
            if (act) cur.seek(S, NB, x_hi);
            // flush geometry: thread r stores column (r % TW) of rows r / TW, r / TW + RPI / TW, ...
            const int col = r % TW, r0 = r / TW;
            const bool rows_full = rg * RPI + RPI <= YSN;
            const size_t row0 = ((size_t)z * YSN + rg * RPI + r0) * X + col;
            for (int u = x_lo + BW - 1; u >= x_lo; u--) {
                if (act && u <= x_hi) {
                    const int d = u - cur.es;
                    tile_g[r * (TW + 1) + (u & (TW - 1))] = d * d + cur.eh;
                    tile_c[r * (TW + 1) + (u & (TW - 1))] = cur.es | ((cur.eb >> 20) << 16);
                    if (u == cur.et && u > 0) cur.prev(S);
                }
                if ((u & (TW - 1)) == 0) {
                    __syncwarp();
                    if (act && u + col < X) {
                        int32_t *pg = g2 + row0 + u, *pc = cxy + row0 + u;
                        const int *tg = tile_g + r0 * (TW + 1) + col, *tc = tile_c + r0 * (TW + 1) + col;
#pragma unroll
                        for (int k = 0; k < TW; k++) {        // RPI / (RPI / TW) = TW rows per thread
                            if (rows_full || rg * RPI + r0 + k * (RPI / TW) < YSN) {
                                pg[(size_t)k * (RPI / TW) * X] = tg[k * (RPI / TW) * (TW + 1)];
                                pc[(size_t)k * (RPI / TW) * X] = tc[k * (RPI / TW) * (TW + 1)];
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
}

// ---- z sweep, dense regime (EDTphase3, local_edt_core.h:137-193) ---------------------------------------------------------
// Same banded machinery along z for volumes with obstacles in most slices, where the envelope of a z column is deep and the
// serial scan of k_edt_zsweep (one lane = one column, 2 x Z dependent steps, stack spilling past 16 entries) runs at a small
// fraction of the write bandwidth.  Work item = (row y, 32 consecutive x); thread = (band of real slices, x).  Loads and stores
// are naturally coalesced along x (no tiles).  Runs only when more than a quarter of the slices hold obstacles; otherwise
// k_edt_zsweep (below), which is bound by its writes and orders them for DRAM page locality, does the work.
__global__ void __launch_bounds__(512)
k_edt_zsweep_banded(LocDev m, const int32_t *__restrict__ g2, const int32_t *__restrict__ cxy, const int *__restrict__ slice_list,
                    const int *__restrict__ n_slices, BandCfg cfg, int *__restrict__ work_counter)
{
    gie_pdl_sync();
    extern __shared__ int zs_smem[];
    __shared__ int s_item;
    const int X = m.X, Z = m.Z;
    const int ns = __ldg(n_slices);
    if (ns * 4 <= Z) return;                             // sparse regime: k_edt_zsweep
    const int NB = cfg.NB, CAP = cfg.CAP, BW = cfg.BW;
    const int lane = threadIdx.x & 31, b = threadIdx.x >> 5;
    BandStacks<32> S;
    S.stH = zs_smem; S.stB = S.stH + NB * CAP * 32; S.meta = S.stB + NB * CAP * 32;
    S.sk = CAP * 32; S.si = 32; S.r = lane;
    const int XG = (X + 31) / 32;
    const int n_items = m.ysn * XG;                      // rows of this map only; the arrays are [Z][ysn][X]
    const size_t slice = (size_t)X * m.ysn;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        const int y = item / XG, x = (item - y * XG) * 32 + lane;
        const bool valid = x < X;
        const size_t base = (size_t)y * X + (valid ? x : 0);
        // ---- forward over this band's real slices; the loads run ahead of the scan
        {
            const int jb = (int)((long long)ns * b / NB), je = (int)((long long)ns * (b + 1) / NB);
            int q = -1, ts = 0, tt = 0, th = 0;
            for (int j0 = jb; j0 < je; j0 += 4) {
                int kk[4], hh[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    kk[k] = __ldg(&slice_list[min(j0 + k, je - 1)]);
                    hh[k] = __ldcs(&g2[base + (size_t)kk[k] * slice]);
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (j0 + k < je) band_push<32>(S, b, kk[k], hh[k], 0, Z, q, ts, tt, th);
            }
            S.M(b) = bd_meta_pack(0, q + 1, 0);
        }
        for (int stride = 1; stride < NB; stride <<= 1) {
            __syncthreads();
            if ((b & (2 * stride - 1)) == 0 && b + stride < NB) band_merge<32>(S, b, stride, NB, Z);
        }
        __syncthreads();
        // ---- backward (local_edt_core.h:169-192) over this band's z range
        {
            const int z_lo = b * BW, z_hi = min(Z, z_lo + BW) - 1;
            if (z_lo <= z_hi) {
                EnvCursor<32> cur{};
                cur.seek(S, NB, z_hi);
                int c = __ldg(&cxy[base + (size_t)cur.es * slice]);
                int coc_word = (c & 0xffff) | ((c >> 16) << 11) | (cur.es << 22);
                int32_t *pa = m.aux + base + (size_t)z_hi * slice, *pc = m.coc_aux + base + (size_t)z_hi * slice;
                for (int u = z_hi; u >= z_lo; u--) {
                    if (valid) {
                        const int d = u - cur.es;
                        __stcs(pa, d * d + cur.eh);
                        __stcs(pc, coc_word);
                    }
                    pa -= slice; pc -= slice;
                    if (u == cur.et && u > 0) {
                        cur.prev(S);
                        c = __ldg(&cxy[base + (size_t)cur.es * slice]);
                        coc_word = (c & 0xffff) | ((c >> 16) << 11) | (cur.es << 22);
                    }
                }
            }
        }
    }
}

constexpr int ZS_WARPS = 8;
constexpr int ZS_RING = 16;    // envelope entries per lane kept in shared memory (deeper ones spill to an L2-resident scratch)
__global__ void __launch_bounds__(ZS_WARPS * 32, 32 / ZS_WARPS)
k_edt_zsweep(LocDev m, const int32_t *__restrict__ g2, const int32_t *__restrict__ cxy, const int *__restrict__ slice_list,
             const int *__restrict__ n_slices, uint2 *__restrict__ scratch, int L, int *__restrict__ work_counter, int n_items,
             int XG, int banded_dense)
{
    gie_pdl_sync();
    __shared__ int zs_ring[ZS_WARPS * 2 * ZS_RING * 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * ZS_WARPS + wid;
    LaneStack<ZS_RING> st;
    st.sh = zs_ring + wid * 2 * ZS_RING * 32 + lane; st.sb = st.sh + ZS_RING * 32;
    st.g = scratch + (size_t)gwarp * L * 32 + lane;
    const int X = m.X, Z = m.Z, S = m.max_width;
    const size_t slice = (size_t)X * m.ysn;            // the arrays are [Z][ysn][X]: a slab of the rows, or all of them
    const int ns = __ldg(n_slices);
    if (ns * 4 > Z && banded_dense) return;   // dense regime: k_edt_zsweep_banded does the work
    // Work item = 32 consecutive x of one row y, one warp each, pulled from a counter per warp (no CTA-wide synchronisation).
    // (Round 1 gave a CTA 8 adjacent x groups and walked z in lockstep for DRAM page locality; with two output arrays instead of
    // three the barrier costs more than it gains — 0.32 vs 0.30 ms — and a synthetic kernel with this store pattern reaches
    // 5.5-6.0 TB/s without any ordering, scratch/write_pattern.cu.)
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int y = item / XG, xg = item - y * XG;
        const int x = xg * 32 + lane;
        const bool valid = x < X;
        const size_t base = (size_t)y * X + (valid ? x : 0);
        if (ns == 0) {   // no obstacle anywhere: every voxel "sees nothing" (D5)
            for (int u = 0; u < Z && valid; u++) {
                m.aux[base + (size_t)u * slice] = S * S;
                m.coc_aux[base + (size_t)u * slice] = x | (INV_Y << 11) | (u << 22);
            }
            continue;
        }
        int q = -1;
        Top top{0, 0, 0, 0};
        st.base = 0;
        // forward (EDTphase3, local_edt_core.h:146-168) over the slices that hold obstacles; the loads run ahead of the scan.
        // The closest obstacle of the slice (cocx | cocy << 16) travels with the entry (20 bits of payload), so that the backward
        // pass never has to fetch it when the owner changes — that dependent load was its largest stall.
        for (int j0 = 0; j0 < ns; j0 += 8) {
            int kk[8], hh[8], cc[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                kk[k] = __ldg(&slice_list[min(j0 + k, ns - 1)]);
                hh[k] = __ldcs(&g2[base + (size_t)kk[k] * slice]);
                cc[k] = __ldcs(&cxy[base + (size_t)kk[k] * slice]);
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (j0 + k < ns) envelope_push(kk[k], hh[k], (cc[k] & 0x7ff) | ((cc[k] >> 16) << 11) | (kk[k] << 22), Z, q, top, st);
        }
        // backward (local_edt_core.h:169-192).  All lanes of the warp walk the same z (the stores stay full 128-byte lines); the
        // walk is cut at the next z where ANY lane's envelope changes owner (warp max of the entries' starts), so that between
        // two such events a z step is d*d + h, two streaming stores and one index bump (the first version tested `u == top.t`
        // and bumped two 64-bit pointers every step: 23 instructions per step).
        // (_dist_id_pair is NOT written here: the reference leaves the pair of UNKNOWN voxels stale and the wavefronts relax
        // against those stale words, so only k_mark_blocks writes it, for known voxels — unify_helper.cuh:217-218.)
        size_t o = base + (size_t)(Z - 1) * slice;
        // Every z step tests this lane's own breakpoint; the pop path runs for the warp whenever any lane changes owner, and the
        // warp is re-converged after it: lanes that drift apart would store partial lines (measured: 6 ms instead of 0.28).
        // (Cutting the walk at the warp-wide maximum of the breakpoints instead — a branch-free inner loop between cuts —
        // executes fewer instructions but times the same on the scene and 5 % slower on dense volumes: the sweep is bound by
        // the memory system, profiles/r02_edt_experiments.md.)
        int32_t *pa = m.aux + o, *pc = m.coc_aux + o;
        int s_ = top.s, h_ = top.h, cw_ = top.cw, t_ = top.t;
        for (int v = Z - 1;; v--) {
            const int d = v - s_;
            if (valid) {
                __stcs(pa, d * d + h_);
                __stcs(pc, cw_);
            }
            if (v == t_) {                       // entry 0 starts at 0, every other entry above it: all lanes leave at v == 0
                if (q == 0) break;
                q--;
                const Top e = st.get(q);
                s_ = e.s; h_ = e.h; cw_ = e.cw; t_ = e.t;
            }
            __syncwarp();
            pa -= slice; pc -= slice;
        }
    }
}

// y-pass bits: kept in step by the OGM merge (hashmap.cu) when the merge was the last to write glb_type, else from the whole array
// slice flags are valid when the merge was the last to write the bits (launch_ybits found them in step)
const int *ybits_slice_flags(gie_locmap *lm)
{
    return lm->hm && lm->ytab_serial >= 0 && lm->ytab_serial == lm->hm->merge_serial ? lm->slice_has : nullptr;
}
int launch_ybits(gie_locmap *lm, int WY)
{
    const LocDev &m = lm->d;
    gie_hashmap *hm = lm->hm;
    // in step with the last OGM merge (hashmap.cu: cleared with the previous merge's blocks, set by k_merge_ogm): nothing to do
    if (hm && hm->merge_serial > 0 && lm->ytab_serial == hm->merge_serial && !lm->glb_type_foreign && !lm->slab_only) return GIE_OK;
    lm->ytab_serial = -1;
    if (m.X % 4 == 0) gie_launch(k_edt_ybits<4>, dim3((m.X / 4 + 127) / 128, WY, m.Z), dim3(128), 0, lm->stream, m, lm->ytab, WY);
    else gie_launch(k_edt_ybits<1>, dim3((m.X + 127) / 128, WY, m.Z), dim3(128), 0, lm->stream, m, lm->ytab, WY);
    return GIE_OK;
}

}  // namespace

int gie_edt_prepare(gie_locmap *lm)
{
    const LocDev &m = lm->d;
    int WY = (m.Y + 31) / 32;
    const size_t slab_voxels = (size_t)m.Z * m.ysn * m.X;   // == N for a whole map
    GIE_CUDA_CHECK(cudaMalloc(&lm->ytab, (size_t)m.Z * WY * m.X * 8));
    GIE_CUDA_CHECK(cudaMalloc(&lm->g2, slab_voxels * 4));
    GIE_CUDA_CHECK(cudaMalloc(&lm->cxy, slab_voxels * 4));
    GIE_CUDA_CHECK(cudaMalloc(&lm->col_list, (size_t)m.Z * m.X * 4));
    GIE_CUDA_CHECK(cudaMalloc(&lm->edt_meta, (size_t)(2 * m.Z + 8) * 4));   // n_cols[Z], slice_list[Z], n_slices
    GIE_CUDA_CHECK(cudaMalloc(&lm->slice_has, (size_t)m.Z * 4));
    GIE_CUDA_CHECK(cudaMemset(lm->slice_has, 0, (size_t)m.Z * 4));
    int L = m.X > m.Z ? m.X : m.Z;
    // serial z sweep: persistent CTAs of ZS_WARPS warps, 32 warps per SM; every warp pulls (row, 32 x) items
    int n_items = m.ysn * ((m.X + 31) / 32);
    int ctas = lm->num_sms * (32 / ZS_WARPS);
    int need = (n_items + ZS_WARPS - 1) / ZS_WARPS;
    if (need < ctas) ctas = need;   // small volumes: no idle persistent CTAs
    lm->edt_ctas = ctas;
    // x sweep: one CTA per (slice, 16 rows) item; bands of ~16 columns, two bands per warp; persistent CTAs, as many as fit
    {
        XsLaunch &x = lm->xs;
        const int nwarps = std::min(16, (m.X + 31) / 32);
        x.rpi = XS_RPI; x.threads = nwarps * 32;
        x.NB = nwarps * (32 / XS_RPI);
        x.CAP = (m.X + x.NB - 1) / x.NB;
        x.BW = ((m.X + x.NB - 1) / x.NB + XS_TW - 1) / XS_TW * XS_TW;
        x.smem = (size_t)(2 * x.NB * x.CAP * XS_RPI + x.NB * XS_RPI + x.NB * 2 * XS_RPI * (XS_TW + 1) + ((m.X + 1) & ~1) + 2 * m.X) * 4;
        GIE_CUDA_CHECK(cudaFuncSetAttribute(k_edt_xsweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)x.smem));
        int xs_per_sm = 1;
        GIE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&xs_per_sm, k_edt_xsweep, x.threads, x.smem));
        if (xs_per_sm < 1) { gie_set_error("x sweep does not fit on an SM"); return GIE_ERR_CUDA; }
        lm->xs_ctas = std::min(lm->num_sms * xs_per_sm, m.Z * ((m.ysn + XS_RPI - 1) / XS_RPI));
    }
    // z sweep, dense regime: one band per warp; only when the stacks of a whole z column fit shared memory (Z <= ~880)
    {
        XsLaunch &zb = lm->zs;
        const int nwarps = std::min(16, (m.Z + 31) / 32);
        zb.rpi = 32; zb.threads = nwarps * 32; zb.NB = nwarps;
        zb.CAP = (m.Z + zb.NB - 1) / zb.NB; zb.BW = zb.CAP;
        zb.smem = (size_t)(2 * zb.NB * zb.CAP * 32 + zb.NB * 32) * 4;
        int dev_max = 0;
        GIE_CUDA_CHECK(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, lm->device));
        lm->zs_banded = zb.smem <= (size_t)dev_max && getenv("GIE_ZS_BANDED") != nullptr;   // off by default: measured slower than the serial sweep (profiles/)
        if (lm->zs_banded) {
            GIE_CUDA_CHECK(cudaFuncSetAttribute(k_edt_zsweep_banded, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zb.smem));
            int per_sm = 1;
            GIE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_edt_zsweep_banded, zb.threads, zb.smem));
            if (per_sm < 1) lm->zs_banded = false;
            lm->zs_ctas = std::min(lm->num_sms * std::max(per_sm, 1), m.ysn * ((m.X + 31) / 32));
        }
    }
    lm->stack_scratch_entries = (size_t)ctas * ZS_WARPS * L * 32;
    GIE_CUDA_CHECK(cudaMalloc(&lm->stack_scratch, lm->stack_scratch_entries * 8));
    GIE_CUDA_CHECK(cudaMalloc(&lm->work_counters, 4 * sizeof(int)));
    return GIE_OK;
}

static void launch_xsweep(gie_locmap *lm, int WY, int *n_cols, int *slice_list, int *n_slices)
{
    const XsLaunch &x = lm->xs;
    const BandCfg cfg{ x.NB, x.CAP, x.BW };
    gie_launch(k_edt_xsweep, dim3(lm->xs_ctas), dim3(x.threads), x.smem, lm->stream, lm->d, lm->ytab, WY, lm->col_list, n_cols, slice_list, n_slices, lm->g2, lm->cxy,
                                                                cfg, lm->work_counters + 0, lm->edt_compact ? 1 : 0);
}

// dense regime first (returns at once when few slices hold obstacles), then the write-ordered serial sweep (returns at once in
// the dense regime when the banded kernel is available)
static void launch_zsweep(gie_locmap *lm, const LocDev &m, int *slice_list, int *n_slices)
{
    const int XG = (m.X + 31) / 32;
    const int L = m.X > m.Z ? m.X : m.Z;
    const int n_items = m.ysn * XG;
    if (lm->zs_banded) {
        const BandCfg cfg{ lm->zs.NB, lm->zs.CAP, lm->zs.BW };
        gie_launch(k_edt_zsweep_banded, dim3(lm->zs_ctas), dim3(lm->zs.threads), lm->zs.smem, lm->stream, m, lm->g2, lm->cxy, slice_list, n_slices, cfg,
                                                                                     lm->work_counters + 2);
        lm->launches++;
    }
    gie_launch(k_edt_zsweep, dim3(lm->edt_ctas), dim3(ZS_WARPS * 32), 0, lm->stream, m, lm->g2, lm->cxy, slice_list, n_slices, (uint2 *)lm->stack_scratch, L,
                                                                 lm->work_counters + 1, n_items, XG, lm->zs_banded ? 1 : 0);
    lm->launches++;
}

// The two halves of the batch EDT as separate launches, for the multi-GPU path (gie-mapping_b200/sharded.py): the y and x
// sweeps are local to a z-slab of the volume, the z sweep needs whole z columns and runs after the slabs were re-partitioned.
int gie_launch_edt_xy(gie_locmap *lm)
{
    const LocDev &m = lm->d;
    const int WY = (m.Y + 31) / 32;
    int *n_cols = lm->edt_meta, *slice_list = lm->edt_meta + m.Z, *n_slices = lm->edt_meta + 2 * m.Z;
    GIE_CUDA_CHECK(cudaMemsetAsync(lm->work_counters, 0, 4 * sizeof(int), lm->stream));
    launch_ybits(lm, WY);
    gie_launch(k_edt_ycols, dim3(m.Z), dim3(((m.X + 31) / 32) * 32), 0, lm->stream, m, lm->ytab, WY, lm->col_list, n_cols, ybits_slice_flags(lm));
    gie_launch(k_edt_slices, dim3(1), dim3(((m.Z + 31) / 32) * 32), 0, lm->stream, m.Z, n_cols, slice_list, n_slices, (int *)nullptr);
    launch_xsweep(lm, WY, n_cols, slice_list, n_slices);
    lm->launches += 4;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}
int gie_launch_edt_z(gie_locmap *lm, int max_width_override)
{
    LocDev m = lm->d;
    if (max_width_override > 0) m.max_width = max_width_override;
    int *n_cols = lm->edt_meta, *slice_list = lm->edt_meta + m.Z, *n_slices = lm->edt_meta + 2 * m.Z;
    GIE_CUDA_CHECK(cudaMemsetAsync(lm->work_counters, 0, 4 * sizeof(int), lm->stream));
    gie_launch(k_edt_slices, dim3(1), dim3(((m.Z + 31) / 32) * 32), 0, lm->stream, m.Z, n_cols, slice_list, n_slices, (int *)nullptr);
    launch_zsweep(lm, m, slice_list, n_slices);
    lm->launches += 1;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

// ---- volumes sharded over GPUs -----------------------------------------------------------------------------
// The map that owns glb_type runs the y pass for the whole volume (gie_launch_edt_pack) and hands ytab / col_list / meta to
// the slab maps, which run the x and z sweeps on their rows (gie_launch_edt_slab).  For a peer GPU only the planes of the
// obstacle-bearing slices travel: k_edt_compact gathers them into [n_slices][...] buffers (the sweeps never read the others).
namespace {
__global__ void __launch_bounds__(256) k_edt_compact(const unsigned long long *__restrict__ ytab, const int *__restrict__ col_list,
                                                     const int *__restrict__ slice_list, const int *__restrict__ n_slices, int WY, int X,
                                                     unsigned long long *__restrict__ ytab_c, int *__restrict__ col_c)
{
    const int ns = __ldg(n_slices);
    const size_t plane = (size_t)WY * X;
    for (int zi = blockIdx.y; zi < ns; zi += gridDim.y) {
        const int z = __ldg(&slice_list[zi]);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x)
            ytab_c[zi * plane + i] = ytab[z * plane + i];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < X; i += gridDim.x * blockDim.x) col_c[(size_t)zi * X + i] = col_list[(size_t)z * X + i];
    }
}
}  // namespace

int gie_launch_edt_pack(gie_locmap *lm, unsigned long long *ytab_compact, int *col_compact)
{
    const LocDev &m = lm->d;
    const int WY = (m.Y + 31) / 32;
    int *n_cols = lm->edt_meta, *slice_list = lm->edt_meta + m.Z, *n_slices = lm->edt_meta + 2 * m.Z;
    StageTimer t(lm, GIE_ST_EDT_PACK);
    launch_ybits(lm, WY);
    gie_launch(k_edt_ycols, dim3(m.Z), dim3(((m.X + 31) / 32) * 32), 0, lm->stream, m, lm->ytab, WY, lm->col_list, n_cols, ybits_slice_flags(lm));
    gie_launch(k_edt_slices, dim3(1), dim3(((m.Z + 31) / 32) * 32), 0, lm->stream, m.Z, n_cols, slice_list, n_slices, (int *)nullptr);
    lm->launches += 3;
    if (ytab_compact && col_compact) {
        k_edt_compact<<<dim3(lm->num_sms, 8), 256, 0, lm->stream>>>(lm->ytab, lm->col_list, slice_list, n_slices, WY, m.X, ytab_compact, col_compact);
        lm->launches++;
    }
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_edt_slab(gie_locmap *lm, int max_width)
{
    LocDev m = lm->d;
    if (max_width > 0) m.max_width = max_width;
    const int WY = (m.Y + 31) / 32;
    int *n_cols = lm->edt_meta, *slice_list = lm->edt_meta + m.Z, *n_slices = lm->edt_meta + 2 * m.Z;
    GIE_CUDA_CHECK(cudaMemsetAsync(lm->work_counters, 0, 4 * sizeof(int), lm->stream));
    {
        StageTimer t(lm, GIE_ST_EDT_X);
        launch_xsweep(lm, WY, n_cols, slice_list, n_slices);
    }
    {
        StageTimer t(lm, GIE_ST_EDT_Z);
        launch_zsweep(lm, m, slice_list, n_slices);
    }
    lm->launches += 1;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_batch_edt(gie_locmap *lm)
{
    const LocDev &m = lm->d;
    const int WY = (m.Y + 31) / 32;
    int *n_cols = lm->edt_meta, *slice_list = lm->edt_meta + m.Z, *n_slices = lm->edt_meta + 2 * m.Z;
    {
        StageTimer t(lm, GIE_ST_EDT_PACK);
        launch_ybits(lm, WY);
        gie_launch(k_edt_ycols, dim3(m.Z), dim3(((m.X + 31) / 32) * 32), 0, lm->stream, m, lm->ytab, WY, lm->col_list, n_cols, ybits_slice_flags(lm));
        gie_launch(k_edt_slices, dim3(1), dim3(((m.Z + 31) / 32) * 32), 0, lm->stream, m.Z, n_cols, slice_list, n_slices, lm->work_counters);
    }
    {
        StageTimer t(lm, GIE_ST_EDT_X);
        launch_xsweep(lm, WY, n_cols, slice_list, n_slices);
    }
    {
        StageTimer t(lm, GIE_ST_EDT_Z);
        launch_zsweep(lm, m, slice_list, n_slices);
    }
    lm->launches += 4;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}
