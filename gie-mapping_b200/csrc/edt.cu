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
//   k_edt_ycols : per (x,z) column packs OCCUPIED bits along y into 32-bit words and, per word, the nearest set bit
//                 below / above the word  -> ytab[z][wy][x] (0.25 B/voxel).  The y pass itself is never materialised.
//   k_edt_xsweep: one warp = 32 consecutive rows y of one slice z.  Lane y derives g1(u,y) from the warp-uniform ytab
//                 entry of column u with two bit scans, runs the lower-envelope scan along x with the stack top in
//                 registers and the body in a per-warp, lane-interleaved scratch ring (L2 resident), and writes its
//                 outputs through a 32x16 shared-memory tile so that global stores are x-contiguous.
//   k_edt_zsweep: one warp = 32 consecutive x of one row y; same scan along z, naturally coalesced; gathers (cocx,cocy)
//                 of the winning slice and emits the final packed result.
// Persistent CTAs (multiple of the SM count) pull work items from an atomic counter.
#include "engine.h"

namespace {

constexpr int WARPS_PER_CTA = 8;
constexpr int INV_Y = 2045;   // INVALID_LOC_COC.y, local_batch.h:59

__global__ void __launch_bounds__(128) k_edt_ycols(LocDev m, unsigned long long *__restrict__ ytab, int WY)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int z = blockIdx.y;
    if (x >= m.X) return;
    const int8_t *col = m.glb_type + (size_t)z * m.X * m.Y + x;
    unsigned long long *out = ytab + (size_t)z * WY * m.X + x;
    int lo_prev = 0xffff;
    for (int wy = 0; wy < WY; wy++) {
        uint32_t w = 0;
        int ybase = wy * 32;
        int n = min(32, m.Y - ybase);
#pragma unroll 8
        for (int b = 0; b < n; b++)
            w |= (uint32_t)(col[(size_t)(ybase + b) * m.X] == GIE_VOX_OCCUPIED) << b;
        out[(size_t)wy * m.X] = (unsigned long long)w | ((unsigned long long)lo_prev << 32);
        if (w) lo_prev = ybase + 31 - __clz(w);
    }
    int hi_next = 0xffff;
    for (int wy = WY - 1; wy >= 0; wy--) {
        unsigned long long e = out[(size_t)wy * m.X];
        out[(size_t)wy * m.X] = e | ((unsigned long long)hi_next << 48);
        uint32_t w = (uint32_t)e;
        if (w) hi_next = wy * 32 + __ffs(w) - 1;
    }
}

// exact floor(num/den) for 0 <= num < 2^24, 0 < den < 2^12 (see the argument in DESIGN.md §4.2: num >= 0 always)
__device__ __forceinline__ int floor_div(int num, int den)
{
    int q = (int)__fdividef((float)num, (float)den);
    int rem = num - q * den;
    if (rem < 0) q--; else if (rem >= den) q++;
    return q;
}

struct Top { int s, t, h, cy; };
__device__ __forceinline__ unsigned long long pack_entry(const Top &e)
{
    return (unsigned long long)(uint32_t)e.h | ((unsigned long long)e.s << 24) | ((unsigned long long)e.t << 35) |
           ((unsigned long long)e.cy << 46);
}
__device__ __forceinline__ Top unpack_entry(unsigned long long p)
{
    Top e;
    e.h = (int)(p & 0xffffff); e.s = (int)((p >> 24) & 0x7ff); e.t = (int)((p >> 35) & 0x7ff); e.cy = (int)((p >> 46) & 0x7ff);
    return e;
}

// one step of the lower-envelope construction (EDTphase2/3 forward loops, local_edt_core.h:93-115 / :146-168)
__device__ __forceinline__ void envelope_push(int u, int h_u, int cy_u, int L, int &q, Top &top,
                                              unsigned long long *__restrict__ stack /* + lane */)
{
    while (q >= 0) {
        int a = top.t - top.s, b = top.t - u;
        if (a * a + top.h > b * b + h_u) {
            q--;
            if (q >= 0) top = unpack_entry(stack[q * 32]);
        } else break;
    }
    if (q < 0) {
        q = 0;
        top.s = u; top.t = 0; top.h = h_u; top.cy = cy_u;
        stack[0] = pack_entry(top);
    } else {
        int num = u * u - top.s * top.s + h_u - top.h;
        int den = 2 * (u - top.s);
        int w = 1 + floor_div(num, den);
        if (w < L) {
            q++;
            top.s = u; top.t = w; top.h = h_u; top.cy = cy_u;
            stack[q * 32] = pack_entry(top);
        }
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_edt_xsweep(LocDev m, const unsigned long long *__restrict__ ytab,
                                                                   int WY, int32_t *__restrict__ g2, int32_t *__restrict__ cxy,
                                                                   unsigned long long *__restrict__ scratch, int L,
                                                                   int *__restrict__ work_counter, int n_items)
{
    __shared__ int tile_g[WARPS_PER_CTA][32][17];
    __shared__ int tile_c[WARPS_PER_CTA][32][17];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * WARPS_PER_CTA + wid;
    unsigned long long *stack = scratch + (size_t)gwarp * L * 32 + lane;
    const int X = m.X, Y = m.Y, S = m.max_width;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int z = item / WY, wy = item - z * WY;
        const int y = wy * 32 + lane;
        const unsigned long long *trow = ytab + ((size_t)z * WY + wy) * X;
        int q = -1;
        Top top{0, 0, 0, 0};
        for (int u = 0; u < X; u++) {
            unsigned long long e = __ldg(&trow[u]);
            uint32_t w = (uint32_t)e;
            int lo_prev = (int)((e >> 32) & 0xffff), hi_next = (int)(e >> 48);
            uint32_t mlo = w & (0xffffffffu >> (31 - lane));
            uint32_t mhi = w >> lane;
            int lo = mlo ? (wy * 32 + 31 - __clz(mlo)) : (lo_prev == 0xffff ? -1 : lo_prev);
            int hi = mhi ? (y + __ffs(mhi) - 1) : (hi_next == 0xffff ? -1 : hi_next);
            int g1, cy;
            if (hi >= 0 && (lo < 0 || hi - y <= y - lo)) { g1 = hi - y; cy = hi; }   // ties -> larger y
            else if (lo >= 0) { g1 = y - lo; cy = lo; }
            else { g1 = S; cy = INV_Y; }
            envelope_push(u, g1 * g1, cy, X, q, top, stack);
        }
        // EDTphase2 backward loop (local_edt_core.h:116-134), emitted through a 32 x 16 tile
        for (int u = X - 1; u >= 0; u--) {
            int d = u - top.s;
            tile_g[wid][lane][u & 15] = d * d + top.h;
            tile_c[wid][lane][u & 15] = top.s | (top.cy << 16);
            if (u == top.t) {
                q--;
                if (q >= 0) top = unpack_entry(stack[q * 32]);
            }
            if ((u & 15) == 0) {
                __syncwarp();
                const int col = lane & 15, r0 = lane >> 4;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    int r = 2 * i + r0;
                    int yy = wy * 32 + r, xx = u + col;
                    if (yy < Y && xx < X) {
                        size_t o = ((size_t)z * Y + yy) * X + xx;
                        g2[o] = tile_g[wid][r][col];
                        cxy[o] = tile_c[wid][r][col];
                    }
                }
                __syncwarp();
            }
        }
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_edt_zsweep(LocDev m, const int32_t *__restrict__ g2,
                                                                   const int32_t *__restrict__ cxy,
                                                                   unsigned long long *__restrict__ scratch, int L,
                                                                   int *__restrict__ work_counter, int n_items, int XG)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * WARPS_PER_CTA + wid;
    unsigned long long *stack = scratch + (size_t)gwarp * L * 32 + lane;
    const int X = m.X, Y = m.Y, Z = m.Z, S = m.max_width;
    const size_t slice = (size_t)X * Y;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int y = item / XG, x = (item - y * XG) * 32 + lane;
        const bool valid = x < X;
        const size_t base = (size_t)y * X + (valid ? x : 0);
        int q = -1;
        Top top{0, 0, 0, 0};
        int k = 0;
        // forward (EDTphase3, local_edt_core.h:146-168); loads run 4 slices ahead of the dependent scan
        for (; k + 4 <= Z; k += 4) {
            int h0 = __ldcs(&g2[base + (size_t)(k + 0) * slice]);
            int h1 = __ldcs(&g2[base + (size_t)(k + 1) * slice]);
            int h2 = __ldcs(&g2[base + (size_t)(k + 2) * slice]);
            int h3 = __ldcs(&g2[base + (size_t)(k + 3) * slice]);
            envelope_push(k + 0, h0, 0, Z, q, top, stack);
            envelope_push(k + 1, h1, 0, Z, q, top, stack);
            envelope_push(k + 2, h2, 0, Z, q, top, stack);
            envelope_push(k + 3, h3, 0, Z, q, top, stack);
        }
        for (; k < Z; k++) envelope_push(k, __ldcs(&g2[base + (size_t)k * slice]), 0, Z, q, top, stack);
        // backward (local_edt_core.h:169-192)
        for (int u = Z - 1; u >= 0; u--) {
            int d = u - top.s;
            int dist = d * d + top.h;
            int c = __ldg(&cxy[base + (size_t)top.s * slice]);
            int cx = c & 0xffff, cy = c >> 16;
            int coc = (cy < S) ? (cx | (cy << 11) | (top.s << 22)) : (x | (INV_Y << 11) | (u << 22));
            if (valid) {
                size_t o = base + (size_t)u * slice;
                m.aux[o] = dist;
                m.coc_aux[o] = coc;
            }
            if (u == top.t) {
                q--;
                if (q >= 0) top = unpack_entry(stack[q * 32]);
            }
        }
    }
}

}  // namespace

int gie_edt_prepare(gie_locmap *lm)
{
    const LocDev &m = lm->d;
    int WY = (m.Y + 31) / 32;
    GIE_CUDA_CHECK(cudaMalloc(&lm->ytab, (size_t)m.Z * WY * m.X * 8));
    GIE_CUDA_CHECK(cudaMalloc(&lm->g2, (size_t)m.N * 4));
    GIE_CUDA_CHECK(cudaMalloc(&lm->cxy, (size_t)m.N * 4));
    int L = m.X > m.Z ? m.X : m.Z;
    int n_items = max(m.Z * WY, m.Y * ((m.X + 31) / 32));
    int ctas = lm->num_sms * 4;
    int need = (n_items + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (need < ctas) ctas = need;   // small volumes: no idle persistent CTAs
    lm->edt_ctas = ctas;
    lm->stack_scratch_entries = (size_t)ctas * WARPS_PER_CTA * L * 32;
    GIE_CUDA_CHECK(cudaMalloc(&lm->stack_scratch, lm->stack_scratch_entries * 8));
    GIE_CUDA_CHECK(cudaMalloc(&lm->work_counters, 4 * sizeof(int)));
    return GIE_OK;
}

int gie_launch_batch_edt(gie_locmap *lm)
{
    const LocDev &m = lm->d;
    const int WY = (m.Y + 31) / 32, XG = (m.X + 31) / 32;
    const int L = m.X > m.Z ? m.X : m.Z;
    GIE_CUDA_CHECK(cudaMemsetAsync(lm->work_counters, 0, 4 * sizeof(int), lm->stream));
    {
        StageTimer t(lm, GIE_ST_EDT_PACK);
        dim3 grid((m.X + 127) / 128, m.Z);
        k_edt_ycols<<<grid, 128, 0, lm->stream>>>(m, lm->ytab, WY);
    }
    {
        StageTimer t(lm, GIE_ST_EDT_X);
        k_edt_xsweep<<<lm->edt_ctas, WARPS_PER_CTA * 32, 0, lm->stream>>>(m, lm->ytab, WY, lm->g2, lm->cxy, lm->stack_scratch,
                                                                          L, lm->work_counters + 0, m.Z * WY);
    }
    {
        StageTimer t(lm, GIE_ST_EDT_Z);
        k_edt_zsweep<<<lm->edt_ctas, WARPS_PER_CTA * 32, 0, lm->stream>>>(m, lm->g2, lm->cxy, lm->stack_scratch, L,
                                                                          lm->work_counters + 1, m.Y * XG, XG);
    }
    lm->launches += 3;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}
