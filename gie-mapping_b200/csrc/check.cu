// check.cu — on-device ground-truth check of the EDT against the occupancy map it was computed from.
//
// Replaces Gnd_truth_checker::cmp_dist (reference include/gt_checker.h:30-80) and the host loops that feed it
// (include/volumetric_mapper.h:181-356: publish_local_ptcld_2_rviz / publish_glb_2_rviz copy the whole local volume and the
// whole host mirror of the hash map into PCL clouds, build a FLANN KD-tree over the OCCUPIED voxels of the global map and
// query it once per EDT voxel, with the rosbag paused).  Same definition here, evaluated where the data lives:
//   occupied cloud = every hash voxel typed OCCUPIED (profile_*_rms use the GLOBAL map's obstacles, volumetric_mapper.h:262-272)
//   EDT cloud      = mode 0 (profile_loc_rms): every known voxel of the local volume with _edt_D          (:198-223)
//                    mode 1 (profile_glb_rms): every known hash voxel with a valid dist_sq, sqrt(dist_sq)   (:274-292)
//   error          = (nearest occupied distance - EDT distance) * voxel_width; counts of |error| > 1 mm in each direction,
//                    sum |e|, sum e^2, max |e|  (gt_checker.h:52-64)
//
// Mechanism: exact nearest neighbour by pruned brute force over voxel blocks.  k_occ_masks turns every allocated block
// into a 512-bit OCCUPIED mask and compacts the blocks that hold an obstacle.  k_check takes one query block per CTA pass
// (thread = voxel): pass 1 bounds the nearest-obstacle distance of ANY voxel of the query block from above with the
// farthest-corner distance to the nearest obstacle block; pass 2 visits only the obstacle blocks whose nearest-corner
// distance does not exceed that bound, with a per-voxel lower-bound test before the set bits of a block are walked.
#include "engine.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

struct CheckAcc {
    unsigned long long n, less, more, n_occ;
    double sum_abs, sum_sq;
    unsigned long long max_abs_bits;   // a non-negative double ordered as an integer
};

__global__ void __launch_bounds__(512) k_occ_masks(HashDev h, int nblocks, uint32_t *__restrict__ masks, int *__restrict__ occ_list,
                                                  int *__restrict__ occ_count, CheckAcc *acc)
{
    __shared__ int any;
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        if (threadIdx.x == 0) any = 0;
        __syncthreads();
        const bool occ = h.vox_type[(size_t)b * 512 + threadIdx.x] == GIE_VOX_OCCUPIED;   // engine order: (z&7)*64 + (y&7)*8 + (x&7)
        const unsigned bal = __ballot_sync(0xffffffffu, occ);
        if ((threadIdx.x & 31) == 0) {
            masks[(size_t)b * 16 + (threadIdx.x >> 5)] = bal;
            if (bal) { atomicOr(&any, 1); atomicAdd(&acc->n_occ, (unsigned long long)__popc(bal)); }
        }
        __syncthreads();
        if (threadIdx.x == 0 && any) occ_list[atomicAdd(occ_count, 1)] = b;
        __syncthreads();
    }
}

// squared distance bounds between two 8^3 voxel cubes whose block keys differ by dk (per axis)
__device__ __forceinline__ int cube_gap(int dk) { int a = abs(dk) * 8 - 7; return a > 0 ? a : 0; }      // nearest corners
__device__ __forceinline__ int cube_span(int dk) { return abs(dk) * 8 + 7; }                            // farthest corners

template <int MODE>
__global__ void __launch_bounds__(512) k_check(LocDev m, HashDev h, const int *__restrict__ qlist, const int *__restrict__ qcount, int nblocks,
                                               const uint32_t *__restrict__ masks, const int *__restrict__ occ_list,
                                               const int *__restrict__ occ_count, int32_t *__restrict__ truth_out, CheckAcc *acc)
{
    __shared__ int s_ub;
    __shared__ int s_cand[512];
    __shared__ int s_ncand;
    __shared__ uint32_t s_mask[16];
    __shared__ int3 s_okey;
    const int nocc = *occ_count;
    const int nq = MODE == 0 ? *qcount : nblocks;
    const int v = threadIdx.x;
    for (int qb = blockIdx.x; qb < nq; qb += gridDim.x) {
        int blk;
        if (MODE == 0) blk = __ldcg(&h.btab[__ldcg(&qlist[qb])]); else blk = qb;
        const int3 kq = h.block_keys[blk];
        const int3 g = make_int3(kq.x * 8 + (v & 7), kq.y * 8 + ((v >> 3) & 7), kq.z * 8 + (v >> 6));
        // is this voxel part of the EDT cloud, and what distance does the map claim for it?
        bool checked;
        float claimed = 0.f;
        int lid = -1;
        if (MODE == 0) {
            const int3 c = g - m.pvt;
            checked = gie_inside_loc(m, c);
            if (checked) { lid = gie_lidx(m, c); checked = m.glb_type[lid] != GIE_VOX_UNKNOWN; }
            if (checked) claimed = m.edt[lid];
        } else {
            const size_t vi = (size_t)blk * 512 + v;
            const int d = h.dist_sq[vi];
            checked = h.vox_type[vi] != GIE_VOX_UNKNOWN && !gie_invalid_dist_glb(d);
            if (checked) claimed = sqrtf((float)d);
        }
        if (threadIdx.x == 0) s_ub = 0x7fffffff;
        __syncthreads();
        if (!__syncthreads_or(checked)) continue;
        // pass 1: upper bound of the nearest-obstacle distance of any voxel of this block
        int ub = 0x7fffffff;
        for (int i = v; i < nocc; i += blockDim.x) {
            const int3 ko = h.block_keys[occ_list[i]];
            const int sx = cube_span(ko.x - kq.x), sy = cube_span(ko.y - kq.y), sz = cube_span(ko.z - kq.z);
            ub = min(ub, sx * sx + sy * sy + sz * sz);
        }
        ub = __reduce_min_sync(0xffffffffu, ub);
        if ((v & 31) == 0) atomicMin(&s_ub, ub);
        __syncthreads();
        ub = s_ub;
        // pass 2: exact search over the obstacle blocks that can hold a nearest obstacle
        int best = 0x7fffffff;
        for (int i0 = 0; i0 < nocc; i0 += blockDim.x) {
            __syncthreads();
            if (threadIdx.x == 0) s_ncand = 0;
            __syncthreads();
            const int i = i0 + v;
            if (i < nocc) {
                const int ob = occ_list[i];
                const int3 ko = h.block_keys[ob];
                const int gx = cube_gap(ko.x - kq.x), gy = cube_gap(ko.y - kq.y), gz = cube_gap(ko.z - kq.z);
                if (gx * gx + gy * gy + gz * gz <= ub) s_cand[atomicAdd(&s_ncand, 1)] = ob;
            }
            __syncthreads();
            const int nc = s_ncand;
            for (int j = 0; j < nc; j++) {
                const int ob = s_cand[j];
                __syncthreads();
                if (v < 16) s_mask[v] = masks[(size_t)ob * 16 + v];
                if (v == 16) s_okey = h.block_keys[ob];
                __syncthreads();
                if (!checked) continue;
                const int3 o0 = make_int3(s_okey.x * 8, s_okey.y * 8, s_okey.z * 8);
                // lower bound from this voxel to the obstacle block's cube
                const int lx = max(0, max(o0.x - g.x, g.x - (o0.x + 7))), ly = max(0, max(o0.y - g.y, g.y - (o0.y + 7))),
                          lz = max(0, max(o0.z - g.z, g.z - (o0.z + 7)));
                if (lx * lx + ly * ly + lz * lz >= best) continue;
#pragma unroll 1
                for (int w = 0; w < 16; w++) {
                    uint32_t bits = s_mask[w];
                    while (bits) {
                        const int bit = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int ov = w * 32 + bit;
                        const int dx = o0.x + (ov & 7) - g.x, dy = o0.y + ((ov >> 3) & 7) - g.y, dz = o0.z + (ov >> 6) - g.z;
                        best = min(best, dx * dx + dy * dy + dz * dz);
                    }
                }
            }
        }
        // gt_checker.h:46-64
        unsigned long long n = 0, less = 0, more = 0;
        double sa = 0.0, ss = 0.0, mx = 0.0;
        if (checked && best != 0x7fffffff) {
            if (MODE == 0 && truth_out) truth_out[lid] = best;
            const double knn = sqrt((double)best) * (double)m.w;
            const double edt = (double)(claimed * m.w);            // float product, as `edt_H[idx]*param.voxel_width`
            const double e = knn - edt;
            n = 1;
            if (e > 0.001) less = 1; else if (e < -0.001) more = 1;
            sa = fabs(e); ss = e * e; mx = fabs(e);
        }
        // block reduction, then one set of atomics per CTA pass
        for (int o = 16; o; o >>= 1) {
            n += __shfl_xor_sync(0xffffffffu, n, o); less += __shfl_xor_sync(0xffffffffu, less, o); more += __shfl_xor_sync(0xffffffffu, more, o);
            sa += __shfl_xor_sync(0xffffffffu, sa, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if ((v & 31) == 0 && n) {
            atomicAdd(&acc->n, n); atomicAdd(&acc->less, less); atomicAdd(&acc->more, more);
            atomicAdd(&acc->sum_abs, sa); atomicAdd(&acc->sum_sq, ss);
            atomicMax(&acc->max_abs_bits, (unsigned long long)__double_as_longlong(mx));
        }
    }
}

}  // namespace

int gie_wave_list_blocks(gie_hashmap *hm);   // wave.cu: refreshes blk_list / blk_count (allocated blocks that intersect the volume)

extern "C" int gie_hashmap_check_edt(gie_hashmap *hm, int mode, int32_t *truth_sq_host, gie_edt_check *out)
{
    if (!hm || !out || (mode != 0 && mode != 1) || (mode == 1 && truth_sq_host)) { gie_set_error("bad check_edt arguments"); return GIE_ERR_INVALID_ARG; }
    gie_locmap *lm = hm->lm;
    cudaStream_t s = lm->stream;
    int nblocks = 0, rc;
    if ((rc = gie_hashmap_num_blocks(hm, &nblocks)) != GIE_OK) return rc;
    memset(out, 0, sizeof(*out));
    out->rms = -1.0;   // cmp_dist's "no checking due to empty cloud" (gt_checker.h:34-40) until something is compared
    if (nblocks == 0) return GIE_OK;
    uint32_t *masks = nullptr;
    int *occ_list = nullptr, *occ_count = nullptr;
    int32_t *truth_dev = nullptr;
    CheckAcc *acc = nullptr;
    auto cleanup = [&]() { cudaFree(masks); cudaFree(occ_list); cudaFree(occ_count); cudaFree(truth_dev); cudaFree(acc); };
#define CHK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { gie_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); cleanup(); return GIE_ERR_CUDA; } } while (0)
    CHK(cudaMalloc(&masks, (size_t)nblocks * 16 * sizeof(uint32_t)));
    CHK(cudaMalloc(&occ_list, (size_t)nblocks * sizeof(int)));
    CHK(cudaMalloc(&occ_count, sizeof(int)));
    CHK(cudaMalloc(&acc, sizeof(CheckAcc)));
    CHK(cudaMemsetAsync(occ_count, 0, sizeof(int), s));
    CHK(cudaMemsetAsync(acc, 0, sizeof(CheckAcc), s));
    if (truth_sq_host) {
        CHK(cudaMalloc(&truth_dev, (size_t)lm->d.N * sizeof(int32_t)));
        CHK(cudaMemsetAsync(truth_dev, 0xff, (size_t)lm->d.N * sizeof(int32_t), s));   // -1 = not part of the EDT cloud
    }
    const int grid = lm->num_sms * 4;
    k_occ_masks<<<std::min(nblocks, grid), 512, 0, s>>>(hm->d, nblocks, masks, occ_list, occ_count, acc);
    if (mode == 0) {
        if ((rc = gie_wave_list_blocks(hm)) != GIE_OK) { cleanup(); return rc; }
        k_check<0><<<grid, 512, 0, s>>>(lm->d, hm->d, hm->blk_list, hm->blk_count, nblocks, masks, occ_list, occ_count, truth_dev, acc);
    } else {
        k_check<1><<<std::min(nblocks, grid), 512, 0, s>>>(lm->d, hm->d, nullptr, nullptr, nblocks, masks, occ_list, occ_count, nullptr, acc);
    }
    lm->launches += 2;
    CHK(cudaGetLastError());
    CheckAcc hacc;
    CHK(cudaMemcpyAsync(&hacc, acc, sizeof(hacc), cudaMemcpyDeviceToHost, s));
    if (truth_sq_host) CHK(cudaMemcpyAsync(truth_sq_host, truth_dev, (size_t)lm->d.N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CHK(cudaStreamSynchronize(s));
#undef CHK
    cleanup();
    out->n = (long long)hacc.n; out->n_occupied = (long long)hacc.n_occ;
    out->edt_less = (long long)hacc.less; out->edt_more = (long long)hacc.more;
    out->sum_abs = hacc.sum_abs; out->sum_sq = hacc.sum_sq;
    long long bits = (long long)hacc.max_abs_bits;
    memcpy(&out->max_abs, &bits, sizeof(double));
    out->rms = hacc.n ? sqrt(hacc.sum_sq / (double)hacc.n) : -1.0;   // cmp_dist returns -1 on an empty cloud (gt_checker.h:34-40)
    return GIE_OK;
}
