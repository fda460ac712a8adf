// ogm.cu — range-sensor integration into the local occupancy volume.
//
// Replaces (reference repo paths):
//   src/kernel/point_cloud/pntcld_raycast.cu:67-117 + ray_cast.h:57-144   point-cloud ray casting
//   src/kernel/hokuyo/hokuyo_fast.cu:9-91   + hokuyo_helper.h:17-33        2-D LiDAR projective
//   src/kernel/vlp16/vlp16_fast.cu:8-97     + vlp16_helper.h:35-64         VLP-16 projective
//   src/kernel/realsense/realsense_fast.cu:9-105 + camera_helper.h:11-23   depth camera projective
//
// This file is compiled with -fmad=false: every float expression below is evaluated exactly as written (IEEE
// single, round-to-nearest), which is what the parity tests compare against.
//
// Differences in mechanism (results identical):
//   * projective kernels run one thread per voxel with x fastest (coalesced) instead of thread=(y,z) looping over x;
//   * no 12 B/voxel block-key array: touched blocks are discovered and allocated by the merge kernel (hashmap.cu).
#include "engine.h"
#include <float.h>

namespace {

__device__ __forceinline__ float3 se3_apply(const float *d, float3 p)
{
    float3 r;
    r.x = d[0] * p.x + d[1] * p.y + d[2] * p.z;
    r.y = d[4] * p.x + d[5] * p.y + d[6] * p.z;
    r.z = d[8] * p.x + d[9] * p.y + d[10] * p.z;
    r.x = r.x + d[3]; r.y = r.y + d[7]; r.z = r.z + d[11];
    return r;
}
__device__ __forceinline__ int3 pos2coord(const LocDev &m, float3 p)
{
    return make_int3((int)floorf(p.x / m.w + 0.5f), (int)floorf(p.y / m.w + 0.5f), (int)floorf(p.z / m.w + 0.5f));
}
__device__ __forceinline__ float3 coord2pos(const LocDev &m, int3 c)
{
    return make_float3((float)c.x * m.w, (float)c.y * m.w, (float)c.z * m.w);
}

// registerLocObs (pntcld_raycast.cu:83-102)
__global__ void k_pc_register(LocDev m, HashDev h, const float *__restrict__ pts, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 p = make_float3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    float3 g = se3_apply(m.L2G, p);
    if (g.z >= m.min_h && g.z <= m.max_h) {
        int3 loc = pos2coord(m, g) - m.pvt;
        if (gie_inside_loc(m, loc)) {
            int id = gie_lidx(m, loc);
            m.inst_type[id] = GIE_VOX_OCCUPIED;
            atomicAdd(&m.ray_count[id], 1);
            gie_touch_block(h, loc + m.pvt);
        }
    }
}

// freeLocObs (pntcld_raycast.cu:67-80) + RAY::rayCastLoc (ray_cast.h:57-144)
//
// Every ray starts at the sensor, so the voxels around the origin are decremented by all ~64 k rays: tens of thousands of
// same-address atomics that serialise in one L2 slice.  Each CTA therefore accumulates the decrements that fall into a
// WIN^3 window around the origin voxel in shared memory and flushes the window once at the end (sums commute, the result
// is identical); decrements outside the window go straight to global memory.
constexpr int RAY_WIN = 16;
__global__ void __launch_bounds__(128) k_pc_free(LocDev m, HashDev h, const float *__restrict__ pts, int n, float max_length)
{
    __shared__ int win[RAY_WIN * RAY_WIN * RAY_WIN];
    for (int k = threadIdx.x; k < RAY_WIN * RAY_WIN * RAY_WIN; k += blockDim.x) win[k] = 0;
    __syncthreads();
    const float3 p0 = m.origin;
    const int3 p0i = pos2coord(m, p0);
    const int3 worg = p0i - make_int3(RAY_WIN / 2, RAY_WIN / 2, RAY_WIN / 2);   // global coords of window cell (0,0,0)
    // decrement of an in-volume voxel given its GLOBAL coords
    int last_ti = -1;
    auto dec = [&](int3 g, int id) {
        int ti = gie_tab_index(h, g);
        if (ti != last_ti) { h.touched[ti] = 1; last_ti = ti; }
        int3 q = g - worg;
        if ((unsigned)q.x < RAY_WIN && (unsigned)q.y < RAY_WIN && (unsigned)q.z < RAY_WIN) atomicAdd(&win[(q.z * RAY_WIN + q.y) * RAY_WIN + q.x], -1);
        else atomicAdd(&m.ray_count[id], -1);   // result unused -> RED
    };
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < n;
    int3 p1i = p0i;
    float3 p1 = p0;
    if (active) {
        float3 p = make_float3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        p1 = se3_apply(m.L2G, p);
        p1i = pos2coord(m, p1);
        // first `opr` on the origin voxel (ray_cast.h:70-72; clearRayLoc pntcld_raycast.cu:9-18)
        int3 loc = p0i - m.pvt;
        if (gie_inside_loc(m, loc)) {
            int id = gie_lidx(m, loc);
            if (m.inst_type[id] != GIE_VOX_OCCUPIED) dec(p0i, id);
        }
        if (eq3(p0i, p1i)) active = false;
    }
    if (active) {
        float dx = p1.x - p0.x, dy = p1.y - p0.y, dz = p1.z - p0.z;
        float len = sqrtf(dx * dx + dy * dy + dz * dz);
        dx = dx / len; dy = dy / len; dz = dz / len;
        int sx = dx > 0.0f ? 1 : (dx < 0.0f ? -1 : 0);
        int sy = dy > 0.0f ? 1 : (dy < 0.0f ? -1 : 0);
        int sz = dz > 0.0f ? 1 : (dz < 0.0f ? -1 : 0);
        float tmx = FLT_MAX, tmy = FLT_MAX, tmz = FLT_MAX, tdx = FLT_MAX, tdy = FLT_MAX, tdz = FLT_MAX;
        if (sx != 0) { float b = (float)p0i.x * m.w + (float)sx * m.w * 0.5f; tmx = (b - p0.x) / dx; tdx = m.w / fabsf(dx); }
        if (sy != 0) { float b = (float)p0i.y * m.w + (float)sy * m.w * 0.5f; tmy = (b - p0.y) / dy; tdy = m.w / fabsf(dy); }
        if (sz != 0) { float b = (float)p0i.z * m.w + (float)sz * m.w * 0.5f; tmz = (b - p0.z) / dz; tdz = m.w / fabsf(dz); }
        int3 cur = p0i;
        // The walk itself is pure arithmetic; what made a step slow was the dependent load of inst_type that decides whether
        // the ray stops.  inst_type is read-only in this kernel, so the DDA runs K steps ahead, the K loads are issued
        // together, and the decisions (stop at the first OCCUPIED voxel, otherwise decrement) are then applied in order —
        // same voxels, same order, same float operations as the one-step-at-a-time loop of ray_cast.h:104-143.
        constexpr int K = 8;
        for (;;) {
            int ids[K];
            int3 gs[K];
            int nsteps = 0;
            bool finished = false;
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (finished) { ids[j] = -1; gs[j] = cur; continue; }
                // comparison tree of ray_cast.h:107-114, reproduced literally
                if (tmx < tmy) {
                    if (tmx < tmz) { cur.x += sx; tmx += tdx; } else { cur.z += sz; tmz += tdz; }
                } else {
                    if (tmy < tmz) { cur.y += sy; tmy += tdy; } else { cur.z += sz; tmz += tdz; }
                }
                int3 loc = cur - m.pvt;
                ids[j] = gie_inside_loc(m, loc) ? gie_lidx(m, loc) : -1;
                gs[j] = cur;
                nsteps = j + 1;
                float d = fminf(fminf(tmx, tmy), tmz);
                finished = eq3(cur, p1i) || d > max_length || d > len;
            }
            int8_t t[K];
#pragma unroll
            for (int j = 0; j < K; j++) t[j] = (j < nsteps && ids[j] >= 0) ? m.inst_type[ids[j]] : (int8_t)GIE_VOX_UNKNOWN;
            bool hit = false;
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (hit || j >= nsteps || ids[j] < 0) continue;
                if (t[j] == GIE_VOX_OCCUPIED) hit = true;            // clearRayLoc returns false: the ray stops here
                else dec(gs[j], ids[j]);
            }
            if (hit || finished) break;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < RAY_WIN * RAY_WIN * RAY_WIN; k += blockDim.x) {
        int v = win[k];
        if (v == 0) continue;
        int3 loc = worg + make_int3(k % RAY_WIN, (k / RAY_WIN) % RAY_WIN, k / (RAY_WIN * RAY_WIN)) - m.pvt;
        atomicAdd(&m.ray_count[gie_lidx(m, loc)], v);   // only in-volume voxels were accumulated
    }
}

// robot sphere of getAllocKeys (pntcld_raycast.cu:33-41): count = -1 inside the sphere, after the ray casting
__global__ void k_pc_sphere(LocDev m, HashDev h, int r2, int r)
{
    int side = 2 * r + 1;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= side * side * side) return;
    int3 d = make_int3(i % side - r, (i / side) % side - r, i / (side * side) - r);
    if (d.x * d.x + d.y * d.y + d.z * d.z > r2) return;
    int3 c = d + m.half;
    if (gie_inside_loc(m, c)) { m.ray_count[gie_lidx(m, c)] = -1; gie_touch_block(h, c + m.pvt); }
}

__device__ __forceinline__ int pos_mod(int i, int n) { return (i % n + n) % n; }
__device__ __forceinline__ bool robot_sphere(const LocDev &m, int3 c, int r2)
{
    int3 d = c - m.half;
    return d.x * d.x + d.y * d.y + d.z * d.z <= r2;
}

enum { SENSOR_SCAN2D = 0, SENSOR_VLP16 = 1, SENSOR_DEPTH = 2 };
struct SensorParam {
    int scan_num, ring_num;
    float theta_inc, theta_min, phi_inc, phi_min;
    int rows, cols;
    float cx, cy, fx, fy;
    int valid_nan;
};

// setLocalOccupancy of the three projective sensors; one thread per voxel, x fastest
template <int SENSOR>
__global__ void __launch_bounds__(256) k_projective(LocDev m, HashDev h, const float *__restrict__ data, SensorParam sp, int fmp, int r2)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y, z = blockIdx.z;
    if (x >= m.X) return;
    int3 c = make_int3(x, y, z);
    int id = gie_lidx(m, c);
    if (fmp && robot_sphere(m, c, r2)) { m.inst_type[id] = GIE_VOX_FREE; gie_touch_block(h, c + m.pvt); return; }
    float3 gp = coord2pos(m, c + m.pvt);
    float3 l = se3_apply(m.G2L, gp);
    int8_t out = GIE_VOX_UNKNOWN;
    if (SENSOR == SENSOR_SCAN2D) {
        float theta = atan2f(l.y, l.x);
        int ti = (int)floorf((theta - sp.theta_min) / sp.theta_inc + 0.5f);
        ti = pos_mod(ti, sp.scan_num);
        float depth = (fabsf(l.z) < m.w) ? sqrtf(l.x * l.x + l.y * l.y) : -1.f;
        if (depth < 0 || ti < 0 || ti >= sp.scan_num) return;
        float real = __ldg(&data[ti]);
        if (isnan(real) || real <= 0.3f) return;
        if (depth < real - 0.3f) out = GIE_VOX_FREE;
        else if ((double)depth > (double)real + 0.3) out = GIE_VOX_UNKNOWN;
        else if (gp.z >= m.min_h && gp.z <= m.max_h) out = GIE_VOX_OCCUPIED;
    } else if (SENSOR == SENSOR_VLP16) {
        float theta = atan2f(l.y, l.x);
        int ti = (int)floorf((theta - sp.theta_min) / sp.theta_inc + 0.5f);
        ti = pos_mod(ti, sp.scan_num);
        float range_hor = sqrtf(l.y * l.y + l.x * l.x);
        float phi = atan2f(l.z, range_hor);
        int pi = (int)floorf((phi - sp.phi_min) / sp.phi_inc + 0.5f);
        if (pi < 0 || pi >= sp.ring_num) return;
        // vlp16_helper.h:57-62: distance of the point to its own ray is ~0, the gate always passes
        float depth = sqrtf(l.x * l.x + l.y * l.y);
        if (depth < 0 || ti < 0 || ti >= sp.scan_num) return;
        float real = __ldg(&data[pi * sp.scan_num + ti]);
        if (isnan(real) || real <= 0.3f) return;
        if (depth < real - 0.1f) { if (depth < real - 0.3f) out = GIE_VOX_FREE; }
        else if ((double)depth > (double)real + 0.1) out = GIE_VOX_UNKNOWN;
        else if (gp.z >= m.min_h && gp.z <= m.max_h) out = GIE_VOX_OCCUPIED;
    } else {
        float depth = l.x;
        if (depth <= 0.3f || depth > 6.0f) return;
        float fpx = floorf(-l.y * sp.fx / depth + sp.cx + 0.5f);
        float fpy = floorf(-l.z * sp.fy / depth + sp.cy + 0.5f);
        if (!(fpx >= 0.f && fpx < (float)sp.cols && fpy >= 0.f && fpy < (float)sp.rows)) return;
        float real = __ldg(&data[sp.cols * (int)fpy + (int)fpx]);
        if (real <= 0.21f) return;
        if (isnan(real)) { if (sp.valid_nan) real = 1000.f; else return; }
        if (depth < real - m.w) out = GIE_VOX_FREE;
        else if (depth > real + m.w) out = GIE_VOX_UNKNOWN;
        else if (gp.z >= m.min_h && gp.z <= m.max_h) out = GIE_VOX_OCCUPIED;
    }
    if (out != GIE_VOX_UNKNOWN) { m.inst_type[id] = out; gie_touch_block(h, c + m.pvt); }
}

template <int SENSOR>
int launch_projective(gie_locmap *lm, gie_hashmap *hm, const float *data, const SensorParam &sp, int fmp, int r2)
{
    StageTimer t(lm, GIE_ST_OGM);
    dim3 block(256), grid((lm->d.X + 255) / 256, lm->d.Y, lm->d.Z);
    if (lm->d.X <= 128) { block = dim3(128); grid.x = (lm->d.X + 127) / 128; }
    k_projective<SENSOR><<<grid, block, 0, lm->stream>>>(lm->d, hm->d, data, sp, fmp, r2);
    lm->launches++;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

// ---- sensor pre-processing on the device (the MapMakers' host loops) -------------------------------------------------
__device__ __forceinline__ float load_f32_unaligned(const unsigned char *p)
{
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    return __uint_as_float(v);
}

// Vlp16MapMaker::convertPyntCld (src/vlp16_map_maker.cpp:73-147): bin = (int)((atan2f(y, x) + (float)M_PI) / |theta_inc|),
// ranges[ring][bin] = sqrtf(x*x + y*y), points visited in message order so the LAST point of a bin wins.  Here every point
// does an atomicMax of (index + 1) << 32 | range bits; the largest index is the last point.
__global__ void k_vlp16_bin(const unsigned char *__restrict__ data, int n, int step, int off_x, int off_y, int off_ring,
                            int scan_num, int ring_num, float resolution, unsigned long long *__restrict__ img)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char *p = data + (size_t)i * step;
    float x = load_f32_unaligned(p + off_x), y = load_f32_unaligned(p + off_y);
    int r = (int)p[off_ring] | ((int)p[off_ring + 1] << 8);
    if (r >= ring_num) return;   // the reference would index past its scan lines here
    int bin = (int)((atan2f(y, x) + 3.14159274f) / resolution);
    if (bin >= 0 && bin < scan_num)
        atomicMax(&img[(size_t)r * scan_num + bin], ((unsigned long long)(uint32_t)(i + 1) << 32) | __float_as_uint(sqrtf(x * x + y * y)));
}
__global__ void k_vlp16_finish(const unsigned long long *__restrict__ img, int n, float *__restrict__ ranges)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long v = img[i];
    ranges[i] = v ? __uint_as_float((uint32_t)v) : INFINITY;   // scan lines start as INFINITY (:55-58)
}
// PntcldMapMaker::pntcld_process (src/pntcld_map_maker.cpp:49-61): the first cld_sz points' three consecutive floats at "x"
__global__ void k_pc_repack(const unsigned char *__restrict__ data, int n, int step, int off_x, float *__restrict__ pts)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char *p = data + (size_t)i * step + off_x;
    pts[3 * i] = load_f32_unaligned(p); pts[3 * i + 1] = load_f32_unaligned(p + 4); pts[3 * i + 2] = load_f32_unaligned(p + 8);
}

}  // namespace

int gie_launch_vlp16_bin(gie_locmap *lm, const unsigned char *raw_dev, int n, int step, int off_x, int off_y, int off_ring,
                         int scan_num, int ring_num, float theta_inc, unsigned long long *img_dev, float *ranges_dev)
{
    const int cells = scan_num * ring_num;
    GIE_CUDA_CHECK(cudaMemsetAsync(img_dev, 0, (size_t)cells * 8, lm->stream));
    if (n > 0) k_vlp16_bin<<<(n + 255) / 256, 256, 0, lm->stream>>>(raw_dev, n, step, off_x, off_y, off_ring, scan_num, ring_num, fabsf(theta_inc), img_dev);
    k_vlp16_finish<<<(cells + 255) / 256, 256, 0, lm->stream>>>(img_dev, cells, ranges_dev);
    lm->launches += n > 0 ? 2 : 1;
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}
int gie_launch_pc_repack(gie_locmap *lm, const unsigned char *raw_dev, int n, int step, int off_x, float *pts_dev)
{
    if (n > 0) { k_pc_repack<<<(n + 255) / 256, 256, 0, lm->stream>>>(raw_dev, n, step, off_x, pts_dev); lm->launches++; }
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_ogm_pointcloud(gie_locmap *lm, gie_hashmap *hm, const float *pts_dev, int n, int fmp, int r2)
{
    StageTimer t(lm, GIE_ST_OGM);
    if (n > 0) {
        int blocks = (n + 255) / 256;
        k_pc_register<<<blocks, 256, 0, lm->stream>>>(lm->d, hm->d, pts_dev, n);
        // pntcld_raycast.cu:79: 0.707f*loc_map._local_size.x*loc_map._voxel_width
        float max_len = 0.707f * (float)lm->d.X * lm->d.w;
        k_pc_free<<<(n + 127) / 128, 128, 0, lm->stream>>>(lm->d, hm->d, pts_dev, n, max_len);
        lm->launches += 2;
    }
    if (fmp) {
        int r = 0;
        while (r * r <= r2) r++;
        int side = 2 * r + 1, tot = side * side * side;
        k_pc_sphere<<<(tot + 255) / 256, 256, 0, lm->stream>>>(lm->d, hm->d, r2, r);
        lm->launches++;
    }
    GIE_CUDA_CHECK(cudaGetLastError());
    return GIE_OK;
}

int gie_launch_ogm_scan2d(gie_locmap *lm, gie_hashmap *hm, const float *scan, int scan_num, float tinc, float tmin, int fmp, int r2)
{
    SensorParam sp{};
    sp.scan_num = scan_num; sp.theta_inc = tinc; sp.theta_min = tmin;
    return launch_projective<SENSOR_SCAN2D>(lm, hm, scan, sp, fmp, r2);
}
int gie_launch_ogm_vlp16(gie_locmap *lm, gie_hashmap *hm, const float *ranges, int scan_num, int ring_num, float tinc,
                         float tmin, float pinc, float pmin, int fmp, int r2)
{
    SensorParam sp{};
    sp.scan_num = scan_num; sp.ring_num = ring_num; sp.theta_inc = tinc; sp.theta_min = tmin; sp.phi_inc = pinc; sp.phi_min = pmin;
    return launch_projective<SENSOR_VLP16>(lm, hm, ranges, sp, fmp, r2);
}
int gie_launch_ogm_depth(gie_locmap *lm, gie_hashmap *hm, const float *img, int rows, int cols, float cx, float cy,
                         float fx, float fy, int valid_nan, int fmp, int r2)
{
    SensorParam sp{};
    sp.rows = rows; sp.cols = cols; sp.cx = cx; sp.cy = cy; sp.fx = fx; sp.fy = fy; sp.valid_nan = valid_nan;
    return launch_projective<SENSOR_DEPTH>(lm, hm, img, sp, fmp, r2);
}
