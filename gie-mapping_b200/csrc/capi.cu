// capi.cu — handle lifecycle and the extern "C" surface declared in include/gie_b200.h.
#include "engine.h"
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>

static thread_local std::string g_last_error;
void gie_set_error(const std::string &msg) { g_last_error = msg; }

namespace {

template <typename T>
__global__ void k_fill(T *p, size_t n, T v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
template <typename T>
int fill_async(T *p, size_t n, T v, cudaStream_t s)
{
    if (n == 0) return GIE_OK;
    int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    k_fill<T><<<blocks, 256, 0, s>>>(p, n, v);
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

// LocMap::convertCostMap (local_batch.h:382-391) packed on the device: SeenDist{d = _edt_D, s = 0, o = (bool)_glb_type}
__global__ void k_costmap(LocDev m, gie_seendist *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.N) return;
    gie_seendist sd;
    sd.d = m.edt[i]; sd.s = 0; sd.o = m.glb_type[i] != 0;
    out[i] = sd;
}

__global__ void k_warmup() {}

int ensure_stage(gie_locmap *lm, size_t bytes)
{
    if (lm->stage_bytes >= bytes) return GIE_OK;
    if (lm->stage_dev) { GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream)); GIE_CUDA_CHECK(cudaFree(lm->stage_dev)); lm->stage_dev = nullptr; }
    GIE_CUDA_CHECK(cudaMalloc(&lm->stage_dev, bytes));
    lm->stage_bytes = bytes;
    return GIE_OK;
}

int check_frame_args(gie_locmap *lm, gie_hashmap *hm)
{
    if (!lm || !hm || hm->lm != lm) { gie_set_error("null or mismatched handle"); return GIE_ERR_INVALID_ARG; }
    return GIE_OK;
}

// The device status word is sticky and is mirrored into pinned host memory at the end of every merge (tail of k_commit), so the
// per-frame entry points can refuse to go on after the block pool ran out WITHOUT a device synchronisation.  The reference
// throws "out of block memory" from its host-side allocator at once (blockalloc.h:56-58); here the failure surfaces on the
// next call after the frame that hit it, and gie_sync() reports it immediately.
int check_sticky_status(gie_hashmap *hm)
{
    const long long st = *(volatile long long *)&hm->stats_host[8];
    if (st & (GIE_DEV_ERR_OUT_OF_BLOCKS | GIE_DEV_ERR_HASH_FULL)) { gie_set_error("out of block memory (raise block_max)"); return GIE_ERR_OUT_OF_BLOCKS; }
    return GIE_OK;
}

struct ArrInfo { void *p; size_t bytes; };
ArrInfo arr_info(gie_locmap *lm, int which)
{
    const LocDev &m = lm->d;
    size_t n = (size_t)m.N;
    const size_t ns = (size_t)m.Z * m.ysn * m.X;   // batch-EDT arrays: the rows this map holds (== n for a whole map)
    if (lm->slab_only && which != GIE_ARR_AUX && which != GIE_ARR_COC_AUX && which != GIE_ARR_EDT_G2 && which != GIE_ARR_EDT_CXY &&
        which != GIE_ARR_EDT_NCOLS) return { nullptr, 0 };
    switch (which) {
        case GIE_ARR_RAY_COUNT: return { m.ray_count, n * 4 };
        case GIE_ARR_INST_TYPE: return { m.inst_type, n };
        case GIE_ARR_GLB_TYPE: return { m.glb_type, n };
        case GIE_ARR_EDT: return { m.edt, n * 4 };
        case GIE_ARR_AUX: return { m.aux, ns * 4 };
        case GIE_ARR_COC_AUX: return { m.coc_aux, ns * 4 };
        case GIE_ARR_PAIR: return { m.pair, n * 8 };
        case GIE_ARR_EDT_G2: return { lm->g2, ns * 4 };
        case GIE_ARR_EDT_CXY: return { lm->cxy, ns * 4 };
        case GIE_ARR_EDT_NCOLS: return { lm->edt_meta, (size_t)m.Z * 4 };
        default: return { nullptr, 0 };
    }
}

}  // namespace

extern "C" {

const char *gie_last_error(void) { return g_last_error.c_str(); }
const char *gie_version(void) { return "gie-b200 0.1 (sm_100a)"; }

int gie_warmup(void)
{
    k_warmup<<<64, 128>>>();
    GIE_CUDA_CHECK(cudaDeviceSynchronize());
    return GIE_OK;
}

int gie_locmap_create(gie_locmap **out, float voxel_size, int X, int Y, int Z, unsigned char thresh, float min_h,
                      float max_h, int cutoff_sq, int fast_mode)
{
    if (!out || X < 1 || Y < 1 || Z < 1 || !(voxel_size > 0.f)) { gie_set_error("bad LocMap arguments"); return GIE_ERR_INVALID_ARG; }
    // coc codec 11/11/10 bit in the wave-range frame (local_batch.h:12-17,51-59) and the 24-bit envelope height
    if (X > 1024 || Y > 1024 || Z > 1022) { gie_set_error("local map size too big (max 1024 x 1024 x 1022)"); return GIE_ERR_SIZE_UNSUPPORTED; }
    gie_locmap *lm = new gie_locmap();
    LocDev &m = lm->d;
    m.X = X; m.Y = Y; m.Z = Z; m.N = X * Y * Z; m.w = voxel_size; m.thresh = thresh; m.min_h = min_h; m.max_h = max_h;
    m.cutoff_sq = cutoff_sq; m.fast = fast_mode ? 1 : 0;
    m.max_width = X + Y + Z;
    m.max_loc_dist_sq = X * X + Y * Y + Z * Z;
    m.half = make_int3(X / 2, Y / 2, Z / 2);
    m.ys0 = 0; m.ysn = Y; m.n_slabs = 1; m.slab_rows = Y;
    GIE_CUDA_CHECK(cudaGetDevice(&lm->device));
    GIE_CUDA_CHECK(cudaDeviceGetAttribute(&lm->num_sms, cudaDevAttrMultiProcessorCount, lm->device));
    size_t n = (size_t)m.N;
    GIE_CUDA_CHECK(cudaMalloc(&m.ray_count, n * 4));
    GIE_CUDA_CHECK(cudaMalloc(&m.inst_type, n));
    GIE_CUDA_CHECK(cudaMalloc(&m.glb_type, n));
    GIE_CUDA_CHECK(cudaMalloc(&m.edt, n * 4));
    GIE_CUDA_CHECK(cudaMalloc(&m.aux, n * 4));
    GIE_CUDA_CHECK(cudaMalloc(&m.coc_aux, n * 4));
    GIE_CUDA_CHECK(cudaMalloc(&m.wave_layer, n * 4));
    GIE_CUDA_CHECK(cudaMalloc(&m.pair, n * 8));
    GIE_CUDA_CHECK(cudaMalloc(&m.nbr_flag, n));
    GIE_CUDA_CHECK(cudaMemset(m.ray_count, 0, n * 4));
    GIE_CUDA_CHECK(cudaMemset(m.inst_type, 0, n));
    GIE_CUDA_CHECK(cudaMemset(m.glb_type, 0, n));
    GIE_CUDA_CHECK(cudaMemset(m.edt, 0, n * 4));
    GIE_CUDA_CHECK(cudaMemset(m.aux, 0, n * 4));
    GIE_CUDA_CHECK(cudaMemset(m.coc_aux, 0, n * 4));
    GIE_CUDA_CHECK(cudaMemset(m.wave_layer, 0, n * 4));
    GIE_CUDA_CHECK(cudaMemset(m.pair, 0, n * 8));
    GIE_CUDA_CHECK(cudaMemset(m.nbr_flag, 0, n));
    int rc = gie_edt_prepare(lm);
    if (rc != GIE_OK) return rc;
    const float q[4] = { 1.f, 0.f, 0.f, 0.f }, t[3] = { 0.f, 0.f, 0.f };
    *out = lm;
    return gie_locmap_set_pose(lm, q, t);
}

int gie_locmap_destroy(gie_locmap *lm)
{
    if (!lm) return GIE_OK;
    cudaStreamSynchronize(lm->stream);
    LocDev &m = lm->d;
    cudaFree(m.ray_count); cudaFree(m.inst_type); cudaFree(m.glb_type); cudaFree(m.edt); cudaFree(m.aux);
    cudaFree(m.coc_aux); cudaFree(m.wave_layer); cudaFree(m.pair); cudaFree(m.nbr_flag);
    if (!lm->edt_inputs_aliased) { cudaFree(lm->ytab); cudaFree(lm->col_list); cudaFree(lm->edt_meta); }
    cudaFree(lm->slice_has); cudaFree(lm->g2); cudaFree(lm->cxy); cudaFree(lm->stack_scratch); cudaFree(lm->work_counters);
    for (int i = 0; i < lm->n_ipc_opened; i++) cudaIpcCloseMemHandle(lm->ipc_opened[i]);
    cudaFree(lm->stage_dev); cudaFree(lm->ray_scratch);
    for (int i = 0; i < GIE_ST_COUNT; i++) for (int j = 0; j < 2; j++) if (lm->ev[i][j]) cudaEventDestroy(lm->ev[i][j]);
    delete lm;
    return GIE_OK;
}

int gie_set_stream(gie_locmap *lm, void *cuda_stream)
{
    if (!lm) return GIE_ERR_INVALID_ARG;
    lm->stream = (cudaStream_t)cuda_stream;
    return GIE_OK;
}

int gie_make_projection(const float q[4], const float t[3], float L2G[12], float G2L[12])
{
    if (!q || !t || !L2G || !G2L) return GIE_ERR_INVALID_ARG;
    // cudaMat::SE3 quaternion constructor (include/cuda_toolkit/se3.cuh:47-75), float, evaluated as written
    volatile float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    float x = 2 * qx, y = 2 * qy, z = 2 * qz;
    float wx = x * qw, wy = y * qw, wz = z * qw;
    float xx = x * qx, xy = y * qx, xz = z * qx, yy = y * qy, yz = z * qy, zz = z * qz;
    float *d = L2G;
    d[0] = 1 - (yy + zz); d[1] = xy - wz; d[2] = xz + wy;
    d[4] = xy + wz; d[5] = 1 - (xx + zz); d[6] = yz - wx;
    d[8] = xz - wy; d[9] = yz + wx; d[10] = 1 - (xx + yy);
    d[3] = t[0]; d[7] = t[1]; d[11] = t[2];
    // SE3::inv (se3.cuh:89-105)
    float *r = G2L;
    r[0] = d[0]; r[1] = d[4]; r[2] = d[8];
    r[4] = d[1]; r[5] = d[5]; r[6] = d[9];
    r[8] = d[2]; r[9] = d[6]; r[10] = d[10];
    r[3] = -d[0] * d[3] - d[4] * d[7] - d[8] * d[11];
    r[7] = -d[1] * d[3] - d[5] * d[7] - d[9] * d[11];
    r[11] = -d[2] * d[3] - d[6] * d[7] - d[10] * d[11];
    return GIE_OK;
}

int gie_locmap_set_projection(gie_locmap *lm, const float L2G[12], const float G2L[12], const float origin[3])
{
    if (!lm || !L2G || !G2L || !origin) return GIE_ERR_INVALID_ARG;
    memcpy(lm->d.L2G, L2G, sizeof(float) * 12);
    memcpy(lm->d.G2L, G2L, sizeof(float) * 12);
    lm->d.origin = make_float3(origin[0], origin[1], origin[2]);
    return GIE_OK;
}

int gie_locmap_calculate_pivots(gie_locmap *lm, const float center[3])
{
    if (!lm || !center) return GIE_ERR_INVALID_ARG;
    LocDev &m = lm->d;
    // LocMap::calculate_pivot_origin / calculate_update_pivot (local_batch.h:128-166)
    int3 c = make_int3((int)floorf(center[0] / m.w + 0.5f), (int)floorf(center[1] / m.w + 0.5f), (int)floorf(center[2] / m.w + 0.5f));
    m.pvt = make_int3(c.x - m.X / 2, c.y - m.Y / 2, c.z - m.Z / 2);
    m.upvt = make_int3(c.x - GIE_WR_X / 2, c.y - GIE_WR_Y / 2, c.z - GIE_WR_Z / 2);
    lm->msg_origin = make_float3((float)m.pvt.x * m.w, (float)m.pvt.y * m.w, (float)m.pvt.z * m.w);
    if (lm->hm) return gie_hash_begin_frame(lm->hm);
    return GIE_OK;
}

int gie_locmap_set_pose(gie_locmap *lm, const float q[4], const float t[3])
{
    if (!lm || !q || !t) return GIE_ERR_INVALID_ARG;
    float L2G[12], G2L[12];
    gie_make_projection(q, t, L2G, G2L);
    gie_locmap_set_projection(lm, L2G, G2L, t);
    return gie_locmap_calculate_pivots(lm, t);
}

int gie_locmap_get_pivots(const gie_locmap *lm, int out6[6], float origin3[3])
{
    if (!lm) return GIE_ERR_INVALID_ARG;
    const LocDev &m = lm->d;
    if (out6) { out6[0] = m.pvt.x; out6[1] = m.pvt.y; out6[2] = m.pvt.z; out6[3] = m.upvt.x; out6[4] = m.upvt.y; out6[5] = m.upvt.z; }
    if (origin3) { origin3[0] = lm->msg_origin.x; origin3[1] = lm->msg_origin.y; origin3[2] = lm->msg_origin.z; }
    return GIE_OK;
}

int gie_locmap_device_ptr(gie_locmap *lm, int which, void **dev_ptr, size_t *bytes)
{
    if (!lm) return GIE_ERR_INVALID_ARG;
    ArrInfo a = arr_info(lm, which);
    if (!a.p) { gie_set_error("unknown array id"); return GIE_ERR_INVALID_ARG; }
    if (dev_ptr) *dev_ptr = a.p;
    if (bytes) *bytes = a.bytes;
    return GIE_OK;
}

int gie_locmap_download(gie_locmap *lm, int which, void *host_out)
{
    if (!lm || !host_out) return GIE_ERR_INVALID_ARG;
    const LocDev &m = lm->d;
    if (m.n_slabs > 1 && (which == GIE_ARR_AUX || which == GIE_ARR_COC_AUX)) {
        // assembled from the slabs (own or peer): slab g is [Z][slab_rows][X], the volume is [Z][Y][X]
        const size_t row = (size_t)m.X * 4, slab_pitch = row * m.slab_rows, vol_pitch = row * m.Y;
        for (int g = 0; g < m.n_slabs; g++) {
            const int32_t *src = which == GIE_ARR_AUX ? m.aux_s[g] : m.coc_s[g];
            GIE_CUDA_CHECK(cudaMemcpy2DAsync((char *)host_out + (size_t)g * slab_pitch, vol_pitch, src, slab_pitch, slab_pitch, m.Z,
                                             cudaMemcpyDeviceToHost, lm->stream));
        }
        GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream));
        return GIE_OK;
    }
    ArrInfo a = arr_info(lm, which);
    if (!a.p) { gie_set_error("unknown array id"); return GIE_ERR_INVALID_ARG; }
    GIE_CUDA_CHECK(cudaMemcpyAsync(host_out, a.p, a.bytes, cudaMemcpyDeviceToHost, lm->stream));
    GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream));
    return GIE_OK;
}

int gie_locmap_upload_glb_type(gie_locmap *lm, const signed char *src)
{
    if (!lm || !src) return GIE_ERR_INVALID_ARG;
    GIE_CUDA_CHECK(cudaMemcpyAsync(lm->d.glb_type, src, (size_t)lm->d.N, cudaMemcpyHostToDevice, lm->stream));
    GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream));
    lm->glb_type_foreign = true;
    return GIE_OK;
}

int gie_locmap_copy_ogm_to_host(gie_locmap *lm, signed char *dst) { return gie_locmap_download(lm, GIE_ARR_GLB_TYPE, dst); }
int gie_locmap_copy_edt_to_host(gie_locmap *lm, float *dst) { return gie_locmap_download(lm, GIE_ARR_EDT, dst); }

int gie_locmap_convert_costmap(gie_locmap *lm, gie_seendist *dst)
{
    if (!lm || !dst) return GIE_ERR_INVALID_ARG;
    size_t bytes = (size_t)lm->d.N * sizeof(gie_seendist);
    int rc = ensure_stage(lm, bytes);
    if (rc != GIE_OK) return rc;
    k_costmap<<<(lm->d.N + 255) / 256, 256, 0, lm->stream>>>(lm->d, (gie_seendist *)lm->stage_dev);
    lm->launches++;
    GIE_CUDA_CHECK(cudaMemcpyAsync(dst, lm->stage_dev, bytes, cudaMemcpyDeviceToHost, lm->stream));
    GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream));
    return GIE_OK;
}

int gie_hashmap_create(gie_hashmap **out, gie_locmap *lm, int bucket_max, int block_max)
{
    if (!out || !lm || block_max < 1) { gie_set_error("bad GlbHashMap arguments"); return GIE_ERR_INVALID_ARG; }
    gie_hashmap *hm = new gie_hashmap();
    hm->lm = lm;
    HashDev &h = hm->d;
    const LocDev &m = lm->d;
    size_t want = std::max<size_t>((size_t)block_max * 2, (size_t)std::max(bucket_max, 1) * 4);
    size_t cap = 1024;
    while (cap < want) cap <<= 1;
    hm->hash_cap = cap;
    h.cap_mask = (uint32_t)(cap - 1);
    h.block_max = block_max;
    cudaStream_t s = lm->stream;
    GIE_CUDA_CHECK(cudaMalloc(&h.keys, cap * 8));
    GIE_CUDA_CHECK(cudaMalloc(&h.vals, cap * 4));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.keys, 0xff, cap * 8, s));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.vals, 0xff, cap * 4, s));
    GIE_CUDA_CHECK(cudaMalloc(&h.block_count, sizeof(int)));
    GIE_CUDA_CHECK(cudaMalloc(&h.status, sizeof(int)));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.block_count, 0, sizeof(int), s));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.status, 0, sizeof(int), s));
    GIE_CUDA_CHECK(cudaMalloc(&h.block_keys, (size_t)block_max * sizeof(int3)));
    size_t nv = (size_t)block_max * 512;
    // every block is default-constructed up front, as the reference does (vhashing.h:519-555; voxmap_utils.cuh:29-44)
    GIE_CUDA_CHECK(cudaMalloc(&h.occ_val, nv));
    GIE_CUDA_CHECK(cudaMalloc(&h.vox_type, nv));
    GIE_CUDA_CHECK(cudaMalloc(&h.update_ct, nv * 4));
    GIE_CUDA_CHECK(cudaMalloc(&h.coc_glb, nv * 8));
    GIE_CUDA_CHECK(cudaMalloc(&h.dist_sq, nv * 4));
    GIE_CUDA_CHECK(cudaMalloc(&h.wave_layer, nv * 4));
    GIE_CUDA_CHECK(cudaMalloc(&h.pair, nv * 8));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.occ_val, 0, nv, s));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.vox_type, 0, nv, s));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.update_ct, 0, nv * 4, s));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.pair, 0, nv * 8, s));
    int rc;
    if ((rc = fill_async<unsigned long long>(h.coc_glb, nv, GIE_EMPTY_COC_PACKED, s)) != GIE_OK) return rc;
    if ((rc = fill_async<int32_t>(h.dist_sq, nv, GIE_EMPTY_VALUE, s)) != GIE_OK) return rc;
    if ((rc = fill_async<int32_t>(h.wave_layer, nv, -1, s)) != GIE_OK) return rc;
    // dense per-frame block table: local volume + a halo wide enough for the outside waves (cutoff) to stay inside it
    int reach = 0;
    while (reach * reach < m.cutoff_sq) reach++;
    hm->halo_blocks = m.fast ? 1 : (reach + 7) / 8 + 2;
    h.tab_dim = make_int3((m.X + 7) / 8 + 1 + 2 * hm->halo_blocks, (m.Y + 7) / 8 + 1 + 2 * hm->halo_blocks,
                          (m.Z + 7) / 8 + 1 + 2 * hm->halo_blocks);
    hm->tab_entries = (size_t)h.tab_dim.x * h.tab_dim.y * h.tab_dim.z;
    GIE_CUDA_CHECK(cudaMalloc(&h.btab, hm->tab_entries * 4));
    GIE_CUDA_CHECK(cudaMalloc(&h.touched, hm->tab_entries));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.touched, 0, hm->tab_entries, s));
    for (auto &bl : hm->blists) {
        GIE_CUDA_CHECK(cudaMalloc(&bl.list, hm->tab_entries * sizeof(int)));
        GIE_CUDA_CHECK(cudaMalloc(&bl.count, sizeof(int)));
        GIE_CUDA_CHECK(cudaMemsetAsync(bl.count, 0, sizeof(int), s));
    }
    GIE_CUDA_CHECK(cudaMalloc(&h.dirty, (size_t)block_max));
    GIE_CUDA_CHECK(cudaMemsetAsync(h.dirty, 0, (size_t)block_max, s));
    GIE_CUDA_CHECK(cudaMalloc(&hm->changed_list, (size_t)block_max * sizeof(int)));
    GIE_CUDA_CHECK(cudaMalloc(&hm->changed_count, sizeof(int)));
    GIE_CUDA_CHECK(cudaMallocHost(&hm->status_host, sizeof(int)));
    GIE_CUDA_CHECK(cudaMallocHost(&hm->stats_host, 16 * sizeof(long long)));
    *hm->status_host = 0;
    memset(hm->stats_host, 0, 16 * sizeof(long long));
    if ((rc = gie_wave_prepare(hm)) != GIE_OK) return rc;
    lm->hm = hm;
    if ((rc = gie_hash_begin_frame(hm)) != GIE_OK) return rc;
    GIE_CUDA_CHECK(cudaStreamSynchronize(s));
    *out = hm;
    return GIE_OK;
}

int gie_hashmap_destroy(gie_hashmap *hm)
{
    if (!hm) return GIE_OK;
    cudaStreamSynchronize(hm->lm->stream);
    HashDev &h = hm->d;
    cudaFree(h.keys); cudaFree(h.vals); cudaFree(h.block_count); cudaFree(h.status); cudaFree(h.block_keys);
    cudaFree(h.occ_val); cudaFree(h.vox_type); cudaFree(h.update_ct); cudaFree(h.coc_glb); cudaFree(h.dist_sq);
    cudaFree(h.wave_layer); cudaFree(h.pair); cudaFree(h.btab); cudaFree(h.touched); for (auto &bl : hm->blists) { cudaFree(bl.list); cudaFree(bl.count); } cudaFree(h.dirty); cudaFree(hm->changed_list); cudaFree(hm->changed_count); cudaFree(hm->obs_dev);
    for (int i = 0; i < 3; i++) { cudaFree(hm->qA[i]); cudaFree(hm->qB[i]); cudaFree(hm->qC[i]); }
    cudaFree(hm->cseed_key); cudaFree(hm->barrier);   // counters and blk_count live inside the barrier allocation
    cudaFree(hm->decA_dist); cudaFree(hm->decA_coc);
    cudaFree(hm->decA_pair); cudaFree(hm->decA_flags); cudaFree(hm->snap_id); cudaFree(hm->wave_trace); cudaFree(hm->blk_list); cudaFree(hm->blk_org);
    cudaFreeHost(hm->status_host); cudaFreeHost(hm->stats_host);
    if (hm->lm->hm == hm) hm->lm->hm = nullptr;
    delete hm;
    return GIE_OK;
}

// ---- volumes sharded over GPUs (no reference counterpart: the reference is single-GPU) ----------------------------------
int gie_locmap_create_slab(gie_locmap **out, int X, int Y, int Z, int row0, int rows)
{
    if (!out || X < 1 || Y < 1 || Z < 1 || row0 < 0 || rows < 1 || row0 + rows > Y || (row0 % 32) != 0) {
        gie_set_error("bad slab arguments (rows must start at a multiple of 32 inside the volume)");
        return GIE_ERR_INVALID_ARG;
    }
    if (X > 1024 || Y > 1024 || Z > 1022) { gie_set_error("local map size too big (max 1024 x 1024 x 1022)"); return GIE_ERR_SIZE_UNSUPPORTED; }
    gie_locmap *lm = new gie_locmap();
    LocDev &m = lm->d;
    m.X = X; m.Y = Y; m.Z = Z; m.N = X * Y * Z; m.w = 1.f;
    m.max_width = X + Y + Z;
    m.max_loc_dist_sq = X * X + Y * Y + Z * Z;
    m.ys0 = row0; m.ysn = rows; m.n_slabs = 1; m.slab_rows = Y;
    lm->slab_only = true;
    GIE_CUDA_CHECK(cudaGetDevice(&lm->device));
    GIE_CUDA_CHECK(cudaDeviceGetAttribute(&lm->num_sms, cudaDevAttrMultiProcessorCount, lm->device));
    const size_t n = (size_t)Z * rows * X;
    GIE_CUDA_CHECK(cudaMalloc(&m.aux, n * 4));
    GIE_CUDA_CHECK(cudaMalloc(&m.coc_aux, n * 4));
    GIE_CUDA_CHECK(cudaMemset(m.aux, 0, n * 4));
    GIE_CUDA_CHECK(cudaMemset(m.coc_aux, 0, n * 4));
    int rc = gie_edt_prepare(lm);
    if (rc != GIE_OK) return rc;
    *out = lm;
    return GIE_OK;
}

int gie_slab_alias_inputs(gie_locmap *slab, gie_locmap *owner)
{
    if (!slab || !owner || !slab->slab_only || slab->device != owner->device || slab->d.X != owner->d.X || slab->d.Y != owner->d.Y ||
        slab->d.Z != owner->d.Z) { gie_set_error("slab and owner maps do not match"); return GIE_ERR_INVALID_ARG; }
    if (!slab->edt_inputs_aliased) { cudaFree(slab->ytab); cudaFree(slab->col_list); cudaFree(slab->edt_meta); }
    slab->ytab = owner->ytab; slab->col_list = owner->col_list; slab->edt_meta = owner->edt_meta;
    slab->edt_inputs_aliased = true; slab->edt_compact = false;
    slab->stream = owner->stream;
    return GIE_OK;
}

int gie_slab_input_buffers(gie_locmap *slab, void **ytab_dev, size_t *ytab_bytes, void **col_list_dev, size_t *col_bytes, void **meta_dev,
                           size_t *meta_bytes)
{
    if (!slab) return GIE_ERR_INVALID_ARG;
    const LocDev &m = slab->d;
    if (ytab_dev) *ytab_dev = slab->ytab;
    if (ytab_bytes) *ytab_bytes = (size_t)m.Z * ((m.Y + 31) / 32) * m.X * 8;
    if (col_list_dev) *col_list_dev = slab->col_list;
    if (col_bytes) *col_bytes = (size_t)m.Z * m.X * 4;
    if (meta_dev) *meta_dev = slab->edt_meta;
    if (meta_bytes) *meta_bytes = (size_t)(2 * m.Z + 8) * 4;
    return GIE_OK;
}

int gie_slab_set_compact(gie_locmap *slab, int compact)
{
    if (!slab || !slab->slab_only) return GIE_ERR_INVALID_ARG;
    slab->edt_compact = compact != 0;
    return GIE_OK;
}

int gie_edt_pack(gie_locmap *lm, void *ytab_compact_dev, void *col_compact_dev)
{
    if (!lm || lm->slab_only) return GIE_ERR_INVALID_ARG;
    return gie_launch_edt_pack(lm, (unsigned long long *)ytab_compact_dev, (int *)col_compact_dev);
}

int gie_edt_slab_sweeps(gie_locmap *slab, int max_width)
{
    if (!slab || max_width < 0) return GIE_ERR_INVALID_ARG;
    return gie_launch_edt_slab(slab, max_width);
}

int gie_ipc_export(void *dev_ptr, unsigned char handle64[64])
{
    if (!dev_ptr || !handle64) return GIE_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    GIE_CUDA_CHECK(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, 64);
    return GIE_OK;
}

int gie_locmap_attach_slabs(gie_locmap *lm, int n_slabs, int slab_rows, void *const *aux_dev, void *const *coc_dev,
                            const unsigned char *aux_handles, const unsigned char *coc_handles)
{
    if (!lm || lm->slab_only || n_slabs < 2 || n_slabs > 8 || slab_rows < 32 || slab_rows % 32 || slab_rows * n_slabs != lm->d.Y) {
        gie_set_error("bad slab layout (2..8 slabs of equal height, a multiple of 32 rows)");
        return GIE_ERR_INVALID_ARG;
    }
    LocDev &m = lm->d;
    for (int g = 0; g < n_slabs; g++) {
        void *pa = aux_dev ? aux_dev[g] : nullptr, *pc = coc_dev ? coc_dev[g] : nullptr;
        if (!pa) {   // a peer process' slab: map it
            if (!aux_handles || !coc_handles) { gie_set_error("slab without pointer or IPC handle"); return GIE_ERR_INVALID_ARG; }
            cudaIpcMemHandle_t ha, hc;
            memcpy(&ha, aux_handles + 64 * g, 64); memcpy(&hc, coc_handles + 64 * g, 64);
            GIE_CUDA_CHECK(cudaIpcOpenMemHandle(&pa, ha, cudaIpcMemLazyEnablePeerAccess));
            GIE_CUDA_CHECK(cudaIpcOpenMemHandle(&pc, hc, cudaIpcMemLazyEnablePeerAccess));
            lm->ipc_opened[lm->n_ipc_opened++] = pa; lm->ipc_opened[lm->n_ipc_opened++] = pc;
        }
        m.aux_s[g] = (int32_t *)pa; m.coc_s[g] = (int32_t *)pc;
    }
    // the whole-volume copies of the batch-EDT arrays are not needed any more
    GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream));
    cudaFree(m.aux); cudaFree(m.coc_aux); cudaFree(lm->g2); cudaFree(lm->cxy);
    m.aux = nullptr; m.coc_aux = nullptr; lm->g2 = nullptr; lm->cxy = nullptr;
    m.n_slabs = n_slabs; m.slab_rows = slab_rows;
    return GIE_OK;
}

// ---- OGM ----------------------------------------------------------------------------------------------------------
int gie_ogm_pointcloud_dev(gie_locmap *lm, gie_hashmap *hm, const float *pts, int n, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (n < 0 || (n > 0 && !pts)) return GIE_ERR_INVALID_ARG;
    return gie_launch_ogm_pointcloud(lm, hm, pts, n, fmp, r2);
}
int gie_ogm_pointcloud_host(gie_locmap *lm, gie_hashmap *hm, const float *pts, int n, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (n < 0 || (n > 0 && !pts)) return GIE_ERR_INVALID_ARG;
    size_t bytes = (size_t)n * 12;
    if ((rc = ensure_stage(lm, std::max<size_t>(bytes, 16))) != GIE_OK) return rc;
    if (n) GIE_CUDA_CHECK(cudaMemcpyAsync(lm->stage_dev, pts, bytes, cudaMemcpyHostToDevice, lm->stream));
    return gie_launch_ogm_pointcloud(lm, hm, lm->stage_dev, n, fmp, r2);
}
int gie_ogm_scan2d_dev(gie_locmap *lm, gie_hashmap *hm, const float *scan, int scan_num, float tinc, float tmin, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (!scan || scan_num < 1 || tinc == 0.f) return GIE_ERR_INVALID_ARG;
    return gie_launch_ogm_scan2d(lm, hm, scan, scan_num, tinc, tmin, fmp, r2);
}
int gie_ogm_scan2d_host(gie_locmap *lm, gie_hashmap *hm, const float *scan, int scan_num, float tinc, float tmin, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (!scan || scan_num < 1 || tinc == 0.f) return GIE_ERR_INVALID_ARG;
    if ((rc = ensure_stage(lm, (size_t)scan_num * 4)) != GIE_OK) return rc;
    GIE_CUDA_CHECK(cudaMemcpyAsync(lm->stage_dev, scan, (size_t)scan_num * 4, cudaMemcpyHostToDevice, lm->stream));
    return gie_launch_ogm_scan2d(lm, hm, lm->stage_dev, scan_num, tinc, tmin, fmp, r2);
}
int gie_ogm_vlp16_dev(gie_locmap *lm, gie_hashmap *hm, const float *ranges, int scan_num, int ring_num, float tinc,
                      float tmin, float pinc, float pmin, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (!ranges || scan_num < 1 || ring_num < 1 || tinc == 0.f || pinc == 0.f) return GIE_ERR_INVALID_ARG;
    return gie_launch_ogm_vlp16(lm, hm, ranges, scan_num, ring_num, tinc, tmin, pinc, pmin, fmp, r2);
}
int gie_ogm_vlp16_host(gie_locmap *lm, gie_hashmap *hm, const float *ranges, int scan_num, int ring_num, float tinc,
                       float tmin, float pinc, float pmin, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (!ranges || scan_num < 1 || ring_num < 1 || tinc == 0.f || pinc == 0.f) return GIE_ERR_INVALID_ARG;
    size_t bytes = (size_t)scan_num * ring_num * 4;
    if ((rc = ensure_stage(lm, bytes)) != GIE_OK) return rc;
    GIE_CUDA_CHECK(cudaMemcpyAsync(lm->stage_dev, ranges, bytes, cudaMemcpyHostToDevice, lm->stream));
    return gie_launch_ogm_vlp16(lm, hm, lm->stage_dev, scan_num, ring_num, tinc, tmin, pinc, pmin, fmp, r2);
}
int gie_ogm_depth_dev(gie_locmap *lm, gie_hashmap *hm, const float *img, int rows, int cols, float cx, float cy, float fx,
                      float fy, int valid_nan, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (!img || rows < 1 || cols < 1) return GIE_ERR_INVALID_ARG;
    return gie_launch_ogm_depth(lm, hm, img, rows, cols, cx, cy, fx, fy, valid_nan, fmp, r2);
}
int gie_ogm_depth_host(gie_locmap *lm, gie_hashmap *hm, const float *img, int rows, int cols, float cx, float cy, float fx,
                       float fy, int valid_nan, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (!img || rows < 1 || cols < 1) return GIE_ERR_INVALID_ARG;
    size_t bytes = (size_t)rows * cols * 4;
    if ((rc = ensure_stage(lm, bytes)) != GIE_OK) return rc;
    GIE_CUDA_CHECK(cudaMemcpyAsync(lm->stage_dev, img, bytes, cudaMemcpyHostToDevice, lm->stream));
    return gie_launch_ogm_depth(lm, hm, lm->stage_dev, rows, cols, cx, cy, fx, fy, valid_nan, fmp, r2);
}

// ---- raw PointCloud2 front ends ----------------------------------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int gie_ogm_vlp16_pointcloud2_host(gie_locmap *lm, gie_hashmap *hm, const void *data, int n_points, int point_step, int off_x, int off_y,
                                   int off_ring, int scan_num, int ring_num, float tinc, float tmin, float pinc, float pmin, int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (n_points < 0 || (n_points > 0 && !data) || point_step < 1 || off_x < 0 || off_y < 0 || off_ring < 0 || off_x + 4 > point_step ||
        off_y + 4 > point_step || off_ring + 2 > point_step || scan_num < 1 || ring_num < 1 || tinc == 0.f || pinc == 0.f) {
        gie_set_error("bad PointCloud2 layout");
        return GIE_ERR_INVALID_ARG;
    }
    const size_t raw = align256((size_t)n_points * point_step), cells = (size_t)scan_num * ring_num;
    if ((rc = ensure_stage(lm, raw + align256(cells * 8) + cells * 4 + 256)) != GIE_OK) return rc;
    unsigned char *raw_dev = (unsigned char *)lm->stage_dev;
    unsigned long long *img = (unsigned long long *)(raw_dev + raw);
    float *ranges = (float *)(raw_dev + raw + align256(cells * 8));
    if (n_points) GIE_CUDA_CHECK(cudaMemcpyAsync(raw_dev, data, (size_t)n_points * point_step, cudaMemcpyHostToDevice, lm->stream));
    if ((rc = gie_launch_vlp16_bin(lm, raw_dev, n_points, point_step, off_x, off_y, off_ring, scan_num, ring_num, tinc, img, ranges)) != GIE_OK) return rc;
    return gie_launch_ogm_vlp16(lm, hm, ranges, scan_num, ring_num, tinc, tmin, pinc, pmin, fmp, r2);
}

int gie_vlp16_last_ranges(gie_locmap *lm, int n_points, int point_step, int scan_num, int ring_num, float *ranges_host)
{
    if (!lm || !ranges_host || !lm->stage_dev) return GIE_ERR_INVALID_ARG;
    const size_t raw = align256((size_t)n_points * point_step), cells = (size_t)scan_num * ring_num;
    if (lm->stage_bytes < raw + align256(cells * 8) + cells * 4) return GIE_ERR_INVALID_ARG;
    GIE_CUDA_CHECK(cudaMemcpyAsync(ranges_host, (char *)lm->stage_dev + raw + align256(cells * 8), cells * 4, cudaMemcpyDeviceToHost, lm->stream));
    GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream));
    return GIE_OK;
}

int gie_ogm_pointcloud2_host(gie_locmap *lm, gie_hashmap *hm, const void *data, int n_points, int point_step, int off_x, int max_points,
                             int fmp, int r2)
{
    int rc = check_frame_args(lm, hm);
    if (rc != GIE_OK) return rc;
    if (n_points < 0 || (n_points > 0 && !data) || point_step < 12 || off_x < 0 || off_x + 12 > point_step) {
        gie_set_error("bad PointCloud2 layout");
        return GIE_ERR_INVALID_ARG;
    }
    const int n = (max_points > 0 && n_points > max_points) ? max_points : n_points;   // cld_sz cap, pntcld_map_maker.cpp:55
    const size_t raw = align256((size_t)n * point_step);
    if ((rc = ensure_stage(lm, raw + (size_t)n * 12 + 256)) != GIE_OK) return rc;
    unsigned char *raw_dev = (unsigned char *)lm->stage_dev;
    float *pts = (float *)(raw_dev + raw);
    if (n) GIE_CUDA_CHECK(cudaMemcpyAsync(raw_dev, data, (size_t)n * point_step, cudaMemcpyHostToDevice, lm->stream));
    if ((rc = gie_launch_pc_repack(lm, raw_dev, n, point_step, off_x, pts)) != GIE_OK) return rc;
    return gie_launch_ogm_pointcloud(lm, hm, pts, n, fmp, r2);
}

// ---- per-frame stages -----------------------------------------------------------------------------------------------
int gie_hashmap_update_ogm(gie_hashmap *hm, int input_pntcld, int map_ct, int stream_glb_ogm, int n_obs, const float *obs_ll,
                           const float *obs_ur, const unsigned char *obs_activated)
{
    if (!hm || n_obs < 0 || (n_obs > 0 && (!obs_ll || !obs_ur || !obs_activated))) return GIE_ERR_INVALID_ARG;
    { int rc = check_sticky_status(hm); if (rc != GIE_OK) return rc; }
    // only activated boxes travel (Ext_Obs_Wrapper::bbx_H2D uploads all of them every frame, pre_map.cu:50-60);
    // slot 0 keeps its meaning as the fence even when it is off
    int n_dev = 0;
    bool any = false;
    for (int i = 0; i < n_obs; i++) any |= obs_activated[i] != 0;
    if (any) {
        std::vector<float> pack;
        for (int i = 0; i < n_obs; i++) {
            if (i > 0 && !obs_activated[i]) continue;
            const float row[7] = { obs_ll[3 * i], obs_ll[3 * i + 1], obs_ll[3 * i + 2], obs_ur[3 * i], obs_ur[3 * i + 1], obs_ur[3 * i + 2],
                                   obs_activated[i] ? 1.f : 0.f };
            pack.insert(pack.end(), row, row + 7);
        }
        n_dev = (int)pack.size() / 7;
        if (n_dev > hm->obs_cap) {
            GIE_CUDA_CHECK(cudaStreamSynchronize(hm->lm->stream));
            cudaFree(hm->obs_dev);
            GIE_CUDA_CHECK(cudaMalloc(&hm->obs_dev, (size_t)n_dev * 2 * 7 * sizeof(float)));
            hm->obs_cap = n_dev * 2;
        }
        // pageable source: the copy is staged before the call returns
        GIE_CUDA_CHECK(cudaMemcpyAsync(hm->obs_dev, pack.data(), pack.size() * sizeof(float), cudaMemcpyHostToDevice, hm->lm->stream));
    }
    return gie_launch_update_ogm(hm, input_pntcld, map_ct, stream_glb_ogm, n_dev);
}
int gie_edt_batch_update(gie_locmap *lm)
{
    if (!lm) return GIE_ERR_INVALID_ARG;
    if (lm->d.n_slabs > 1 || lm->slab_only) { gie_set_error("sharded volume: use gie_edt_pack + gie_edt_slab_sweeps"); return GIE_ERR_INVALID_ARG; }
    return gie_launch_batch_edt(lm);
}
int gie_edt_xy_sweeps(gie_locmap *lm)
{
    if (!lm) return GIE_ERR_INVALID_ARG;
    return gie_launch_edt_xy(lm);
}
int gie_edt_z_sweep(gie_locmap *lm, int max_width_override)
{
    if (!lm || max_width_override < 0) return GIE_ERR_INVALID_ARG;
    return gie_launch_edt_z(lm, max_width_override);
}
int gie_hashmap_merge_new_obsv(gie_hashmap *hm, int map_ct, int display_glb_edt)
{
    if (!hm) return GIE_ERR_INVALID_ARG;
    { int rc = check_sticky_status(hm); if (rc != GIE_OK) return rc; }
    return gie_launch_merge(hm, map_ct, display_glb_edt);
}

int gie_sync(gie_hashmap *hm)
{
    if (!hm) return GIE_ERR_INVALID_ARG;
    cudaStream_t s = hm->lm->stream;
    GIE_CUDA_CHECK(cudaMemcpyAsync(hm->status_host, hm->d.status, sizeof(int), cudaMemcpyDeviceToHost, s));
    GIE_CUDA_CHECK(cudaStreamSynchronize(s));
    int st = *hm->status_host;
    if (st & (GIE_DEV_ERR_OUT_OF_BLOCKS | GIE_DEV_ERR_HASH_FULL)) { gie_set_error("out of block memory (raise block_max)"); return GIE_ERR_OUT_OF_BLOCKS; }
    if (st & GIE_DEV_ERR_QUEUE_OVERFLOW) { gie_set_error("wavefront queue overflow"); return GIE_ERR_QUEUE_OVERFLOW; }
    return GIE_OK;
}

int gie_hashmap_num_blocks(gie_hashmap *hm, int *n)
{
    if (!hm || !n) return GIE_ERR_INVALID_ARG;
    GIE_CUDA_CHECK(cudaMemcpyAsync(n, hm->d.block_count, sizeof(int), cudaMemcpyDeviceToHost, hm->lm->stream));
    GIE_CUDA_CHECK(cudaStreamSynchronize(hm->lm->stream));
    if (*n > hm->d.block_max) *n = hm->d.block_max;
    return GIE_OK;
}

int gie_hashmap_export_blocks(gie_hashmap *hm, int32_t *keys_host, gie_glbvoxel *voxels_host, int max_blocks)
{
    if (!hm || !keys_host || !voxels_host) return GIE_ERR_INVALID_ARG;
    int n = 0, rc;
    if ((rc = gie_hashmap_num_blocks(hm, &n)) != GIE_OK) return rc;
    if (n > max_blocks) n = max_blocks;
    if (n == 0) return GIE_OK;
    cudaStream_t s = hm->lm->stream;
    gie_glbvoxel *tmp = nullptr;
    const int chunk = 4096;   // 80 MB of staging at a time
    GIE_CUDA_CHECK(cudaMalloc(&tmp, (size_t)std::min(n, chunk) * 512 * sizeof(gie_glbvoxel)));
    GIE_CUDA_CHECK(cudaMemcpyAsync(keys_host, hm->d.block_keys, (size_t)n * sizeof(int3), cudaMemcpyDeviceToHost, s));
    for (int b0 = 0; b0 < n; b0 += chunk) {
        int nb = std::min(chunk, n - b0);
        HashDev view = hm->d;
        size_t off = (size_t)b0 * 512;
        view.occ_val += off; view.vox_type += off; view.update_ct += off; view.coc_glb += off; view.dist_sq += off;
        view.wave_layer += off; view.pair += off;
        gie_hashmap shadow = *hm;
        shadow.d = view;
        if ((rc = gie_launch_export(&shadow, nb, tmp)) != GIE_OK) { cudaFree(tmp); return rc; }
        GIE_CUDA_CHECK(cudaMemcpyAsync(voxels_host + off, tmp, (size_t)nb * 512 * sizeof(gie_glbvoxel), cudaMemcpyDeviceToHost, s));
        GIE_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    cudaFree(tmp);
    return GIE_OK;
}

int gie_hashmap_device_view(gie_hashmap *hm, gie_device_view *out)
{
    if (!hm || !out) return GIE_ERR_INVALID_ARG;
    const HashDev &h = hm->d;
    out->keys = h.keys; out->vals = h.vals; out->cap_mask = h.cap_mask; out->block_max = h.block_max; out->block_count = h.block_count;
    out->occ_val = h.occ_val; out->vox_type = (const signed char *)h.vox_type; out->update_ct = h.update_ct; out->coc_glb = h.coc_glb;
    out->dist_sq = h.dist_sq; out->wave_layer = h.wave_layer; out->pair = h.pair;
    return GIE_OK;
}

int gie_hashmap_num_changed(gie_hashmap *hm, int *n)
{
    if (!hm || !n) return GIE_ERR_INVALID_ARG;
    int nb = 0, rc;
    if ((rc = gie_hashmap_num_blocks(hm, &nb)) != GIE_OK) return rc;
    if ((rc = gie_launch_list_changed(hm, nb, 0)) != GIE_OK) return rc;
    GIE_CUDA_CHECK(cudaMemcpyAsync(n, hm->changed_count, sizeof(int), cudaMemcpyDeviceToHost, hm->lm->stream));
    GIE_CUDA_CHECK(cudaStreamSynchronize(hm->lm->stream));
    return GIE_OK;
}

int gie_hashmap_stream_changed(gie_hashmap *hm, int32_t *keys_host, gie_glbvoxel *voxels_host, int max_blocks, int *n_out)
{
    if (!hm || !keys_host || !voxels_host || !n_out || max_blocks < 0) return GIE_ERR_INVALID_ARG;
    gie_locmap *lm = hm->lm;
    cudaStream_t s = lm->stream;
    int nb = 0, rc, n = 0;
    if ((rc = gie_hashmap_num_blocks(hm, &nb)) != GIE_OK) return rc;
    if ((rc = gie_launch_list_changed(hm, nb, 1)) != GIE_OK) return rc;
    GIE_CUDA_CHECK(cudaMemcpyAsync(&n, hm->changed_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    GIE_CUDA_CHECK(cudaStreamSynchronize(s));
    if (n > max_blocks) {   // the caller's buffers are too small: keep the flags of the blocks that do not fit
        // (list order is not sorted; re-flag the tail)
        std::vector<int> tail((size_t)(n - max_blocks));
        GIE_CUDA_CHECK(cudaMemcpy(tail.data(), hm->changed_list + max_blocks, tail.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (int b : tail) GIE_CUDA_CHECK(cudaMemset(hm->d.dirty + b, 1, 1));
        n = max_blocks;
    }
    *n_out = n;
    if (n == 0) return GIE_OK;
    // staging: keys (12 B) + voxels (20 KB) per block, gathered by one kernel and copied with one D2H each
    const size_t vbytes = (size_t)n * 512 * sizeof(gie_glbvoxel), kbytes = (size_t)n * 3 * sizeof(int32_t);
    if ((rc = ensure_stage(lm, vbytes + kbytes)) != GIE_OK) return rc;
    gie_glbvoxel *vdev = (gie_glbvoxel *)lm->stage_dev;
    int32_t *kdev = (int32_t *)((char *)lm->stage_dev + vbytes);
    if ((rc = gie_launch_gather_changed(hm, 0, n, kdev, vdev)) != GIE_OK) return rc;
    GIE_CUDA_CHECK(cudaMemcpyAsync(voxels_host, vdev, vbytes, cudaMemcpyDeviceToHost, s));
    GIE_CUDA_CHECK(cudaMemcpyAsync(keys_host, kdev, kbytes, cudaMemcpyDeviceToHost, s));
    GIE_CUDA_CHECK(cudaStreamSynchronize(s));
    return GIE_OK;
}

int gie_debug_wave_trace(gie_hashmap *hm, unsigned long long *out, int max_levels)
{
    if (!hm || !out || !hm->wave_trace) { gie_set_error("wave trace is off (set GIE_WAVE_TRACE=1 before creating the map)"); return GIE_ERR_INVALID_ARG; }
    GIE_CUDA_CHECK(cudaStreamSynchronize(hm->lm->stream));
    GIE_CUDA_CHECK(cudaMemcpy(out, hm->wave_trace, (size_t)4096 * 10 * 8, cudaMemcpyDeviceToHost));
    return GIE_OK;
}

int gie_hashmap_wave_stats(gie_hashmap *hm, int64_t out8[8])
{
    if (!hm || !out8) return GIE_ERR_INVALID_ARG;
    GIE_CUDA_CHECK(cudaStreamSynchronize(hm->lm->stream));
    for (int i = 0; i < 8; i++) out8[i] = hm->stats_host[i];
    return GIE_OK;
}

// ---- measurement ------------------------------------------------------------------------------------------------------
int gie_profile_enable(gie_locmap *lm, int on)
{
    if (!lm) return GIE_ERR_INVALID_ARG;
    if (on) for (int i = 0; i < GIE_ST_COUNT; i++) for (int j = 0; j < 2; j++) if (!lm->ev[i][j]) GIE_CUDA_CHECK(cudaEventCreate(&lm->ev[i][j]));
    lm->profile = on != 0;
    for (int i = 0; i < GIE_ST_COUNT; i++) lm->ev_valid[i] = false;
    return GIE_OK;
}
int gie_profile_last(gie_locmap *lm, float ms[GIE_ST_COUNT])
{
    if (!lm || !ms) return GIE_ERR_INVALID_ARG;
    GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream));
    for (int i = 0; i < GIE_ST_COUNT; i++) {
        ms[i] = 0.f;
        if (lm->profile && lm->ev_valid[i]) GIE_CUDA_CHECK(cudaEventElapsedTime(&ms[i], lm->ev[i][0], lm->ev[i][1]));
    }
    return GIE_OK;
}
int gie_launch_count(gie_locmap *lm, long long *n)
{
    if (!lm || !n) return GIE_ERR_INVALID_ARG;
    *n = lm->launches;
    return GIE_OK;
}

}  // extern "C"
