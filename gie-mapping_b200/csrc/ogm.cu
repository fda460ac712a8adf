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
    gie_pdl_sync();
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
// One thread per ray, as in the reference, leaves a B200 with ~14 warps per SM, each a chain of a few hundred dependent
// steps (DDA step -> load of the voxel's type, which decides whether the ray stops -> decrement): 170 us at 6 active warps
// per SM.  The walk itself is pure arithmetic, so it is split off:
//   k_pc_walk  : thread = ray.  Runs the DDA without touching memory — the same float operations in the same order as
//                ray_cast.h:104-143, so every visited voxel is the reference's — and checkpoints its state every RAY_SEG steps.
//   k_pc_apply : thread = (ray, segment).  Replays the segment from its checkpoint twice: once loading the types of its voxels
//                (RAY_SEG independent loads) to find the first OCCUPIED one, once decrementing every in-volume voxel before it.
//   k_pc_undo  : thread = (ray, segment).  Segments that turn out to lie behind their ray's stop give back what they took.
// ~20x the threads, every load and atomic independent of the others; the visited set and the counts are unchanged.
// (Folding the type loads into the walk — the walk thread loads the types of a segment and looks at them one segment later —
// measured slower: 0.23 ms for the stage against 0.13, the loads lengthen the one dependent chain that bounds the kernel.)
// Every ray starts at the sensor, so the voxels around the origin are decremented by all rays: the CTAs of the first segment
// accumulate the decrements that fall into a WIN^3 window around the origin voxel in shared memory and flush the window once
// (sums commute, the result is identical).
constexpr int RAY_WIN = 16;
constexpr int RAY_SEG = 32;
struct RaySetup {
    float3 p0, p1;
    int3 p0i, p1i;
    float len, tdx, tdy, tdz, tmx, tmy, tmz;
    int sx, sy, sz;
};
// ray_cast.h:57-103, evaluated as written
__device__ __forceinline__ void ray_setup(const LocDev &m, const float *__restrict__ pts, int i, RaySetup &r)
{
    r.p0 = m.origin;
    r.p0i = pos2coord(m, r.p0);
    float3 p = make_float3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    r.p1 = se3_apply(m.L2G, p);
    r.p1i = pos2coord(m, r.p1);
    float dx = r.p1.x - r.p0.x, dy = r.p1.y - r.p0.y, dz = r.p1.z - r.p0.z;
    r.len = sqrtf(dx * dx + dy * dy + dz * dz);
    dx = dx / r.len; dy = dy / r.len; dz = dz / r.len;
    r.sx = dx > 0.0f ? 1 : (dx < 0.0f ? -1 : 0);
    r.sy = dy > 0.0f ? 1 : (dy < 0.0f ? -1 : 0);
    r.sz = dz > 0.0f ? 1 : (dz < 0.0f ? -1 : 0);
    r.tmx = FLT_MAX; r.tmy = FLT_MAX; r.tmz = FLT_MAX; r.tdx = FLT_MAX; r.tdy = FLT_MAX; r.tdz = FLT_MAX;
    if (r.sx != 0) { float b = (float)r.p0i.x * m.w + (float)r.sx * m.w * 0.5f; r.tmx = (b - r.p0.x) / dx; r.tdx = m.w / fabsf(dx); }
    if (r.sy != 0) { float b = (float)r.p0i.y * m.w + (float)r.sy * m.w * 0.5f; r.tmy = (b - r.p0.y) / dy; r.tdy = m.w / fabsf(dy); }
    if (r.sz != 0) { float b = (float)r.p0i.z * m.w + (float)r.sz * m.w * 0.5f; r.tmz = (b - r.p0.z) / dz; r.tdz = m.w / fabsf(dz); }
}
// What a DDA step needs of the set-up: k_pc_walk stores it per ray (16 bytes), the (ray, segment) threads of
// k_pc_apply load it instead of redoing the set-up's six IEEE divisions and the square root.
struct RayInc { float tdx, tdy, tdz; int sx, sy, sz; };
__device__ __forceinline__ RayInc ray_inc(const RaySetup &r) { return RayInc{ r.tdx, r.tdy, r.tdz, r.sx, r.sy, r.sz }; }
__device__ __forceinline__ float4 ray_inc_pack(const RayInc &r)
{
    return make_float4(r.tdx, r.tdy, r.tdz, __int_as_float((r.sx + 1) | ((r.sy + 1) << 2) | ((r.sz + 1) << 4)));
}
__device__ __forceinline__ RayInc ray_inc_unpack(float4 v)
{
    const int b = __float_as_int(v.w);
    return RayInc{ v.x, v.y, v.z, (b & 3) - 1, ((b >> 2) & 3) - 1, ((b >> 4) & 3) - 1 };
}
// one DDA step: the comparison tree of ray_cast.h:107-114 — (tmx < tmy) ? (tmx < tmz ? x : z) : (tmy < tmz ? y : z) — written
// without branches: the smaller of (x, y) by the first comparison meets z in the second.  Same comparisons on the same
// values, same single addition; the step is the dependent chain that bounds the walk, and the select form halves it.
__device__ __forceinline__ void ray_step(const RayInc &r, int3 &cur, float &tmx, float &tmy, float &tmz)
{
    const bool xm = tmx < tmy;
    const float a = xm ? tmx : tmy;
    const bool zw = !(a < tmz);
    const bool stepx = xm && !zw, stepy = !xm && !zw;
    if (stepx) { cur.x += r.sx; tmx += r.tdx; }
    if (stepy) { cur.y += r.sy; tmy += r.tdy; }
    if (zw) { cur.z += r.sz; tmz += r.tdz; }
}
struct RayCk { float tmx, tmy, tmz; int x, y, z; };   // state before step s * RAY_SEG

__global__ void __launch_bounds__(128) k_pc_walk(LocDev m, HashDev h, const float *__restrict__ pts, int n, float max_length, int max_segs,
                                                 RayCk *__restrict__ ck, float4 *__restrict__ incs, int *__restrict__ nsteps,
                                                 int *__restrict__ stop)
{
    gie_pdl_sync();
    __shared__ int origin_dec;
    if (threadIdx.x == 0) origin_dec = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    RaySetup r;
    if (i < n) {
        ray_setup(m, pts, i, r);
        // first `opr` on the origin voxel (ray_cast.h:70-72; clearRayLoc pntcld_raycast.cu:9-18): one decrement per ray
        const int3 loc = r.p0i - m.pvt;
        if (gie_inside_loc(m, loc) && m.inst_type[gie_lidx(m, loc)] != GIE_VOX_OCCUPIED) atomicAdd(&origin_dec, 1);
        int steps = 0;
        const RayInc inc = ray_inc(r);
        incs[i] = ray_inc_pack(inc);
        if (!eq3(r.p0i, r.p1i)) {
            int3 cur = r.p0i;
            float tmx = r.tmx, tmy = r.tmy, tmz = r.tmz;
            const int cap = max_segs * RAY_SEG;
            for (;;) {
                if ((steps & (RAY_SEG - 1)) == 0) ck[(size_t)(steps / RAY_SEG) * n + i] = RayCk{ tmx, tmy, tmz, cur.x, cur.y, cur.z };
                ray_step(inc, cur, tmx, tmy, tmz);
                steps++;
                const float d = fminf(fminf(tmx, tmy), tmz);
                if (eq3(cur, r.p1i) || d > max_length || d > r.len || steps >= cap) break;
            }
        }
        nsteps[i] = steps;
        stop[i] = steps;   // no OCCUPIED voxel met yet: the whole walk counts
    }
    __syncthreads();
    if (threadIdx.x == 0 && origin_dec) {
        const int3 p0i = pos2coord(m, m.origin);
        const int3 loc = p0i - m.pvt;
        atomicAdd(&m.ray_count[gie_lidx(m, loc)], -origin_dec);
        gie_touch_block(h, p0i);
    }
}

// thread = (ray, segment); CTAs are segment-major so that all threads of a CTA work on the same segment index.
// Pass 1 replays the segment and loads the types of its voxels (independent loads): the first OCCUPIED one is where the ray
// would stop IF no earlier segment stops it (clearRayLoc returns false, pntcld_raycast.cu:9-18).  Pass 2 replays again and
// decrements the voxels before that local stop — speculatively: a segment does not know yet whether an earlier one holds
// the ray's real stop.  k_pc_undo puts back what the segments behind the real stop took (integer sums commute, so
// ray_count ends bit-exact).  This replaces a scan kernel and an apply kernel that each replayed every segment; in the frame
// the stage takes 0.104 ms against 0.107 (the undo pass overlaps the next kernel's launch), although the two kernels alone,
// serialised under ncu, take 72 us against 56: rays that graze a nearer silhouette are stopped early and have several
// segments to give back.  (Listing those rays and undoing them warp-per-ray was slower still: 0.121 ms for the stage.)
__global__ void __launch_bounds__(128) k_pc_apply(LocDev m, HashDev h, int n, int max_segs, const RayCk *__restrict__ ck,
                                                  const float4 *__restrict__ incs, const int *__restrict__ nsteps, int *__restrict__ stop)
{
    gie_pdl_sync();
    __shared__ int win[RAY_WIN * RAY_WIN * RAY_WIN];
    const int seg = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool use_win = seg == 0;                                  // uniform in the CTA
    const int3 p0i = pos2coord(m, m.origin);
    const int3 worg = p0i - make_int3(RAY_WIN / 2, RAY_WIN / 2, RAY_WIN / 2);   // global coords of window cell (0,0,0)
    if (use_win) {
        for (int k = threadIdx.x; k < RAY_WIN * RAY_WIN * RAY_WIN; k += blockDim.x) win[k] = 0;
        __syncthreads();
    }
    const int first = seg * RAY_SEG;
    const int total = i < n ? __ldg(&nsteps[i]) : 0;
    if (first < total) {
        const RayInc r = ray_inc_unpack(__ldg(&incs[i]));
        const RayCk c = ck[(size_t)seg * n + i];   // segment-major: coalesced over the rays of a warp
        const int cnt = min(RAY_SEG, total - first);
        int hit = RAY_SEG;
        {
            int3 cur = make_int3(c.x, c.y, c.z);
            float tmx = c.tmx, tmy = c.tmy, tmz = c.tmz;
#pragma unroll
            for (int j = 0; j < RAY_SEG; j++) {
                if (j < cnt) {
                    ray_step(r, cur, tmx, tmy, tmz);
                    const int3 loc = cur - m.pvt;
                    if (gie_inside_loc(m, loc) && m.inst_type[gie_lidx(m, loc)] == GIE_VOX_OCCUPIED) hit = min(hit, j);
                }
            }
        }
        if (hit < cnt) atomicMin(&stop[i], first + hit);
        const int lim = min(cnt, hit);
        int3 cur = make_int3(c.x, c.y, c.z);
        float tmx = c.tmx, tmy = c.tmy, tmz = c.tmz;
        int last_ti = -1;
        for (int j = 0; j < lim; j++) {
            ray_step(r, cur, tmx, tmy, tmz);
            const int3 loc = cur - m.pvt;
            if (!gie_inside_loc(m, loc)) continue;                   // the walk continues outside the volume (ray_cast.h:116-121)
            const int ti = gie_tab_index(h, cur);
            if (ti != last_ti) { h.touched[ti] = 1; last_ti = ti; }
            const int3 q = cur - worg;
            if (use_win && (unsigned)q.x < RAY_WIN && (unsigned)q.y < RAY_WIN && (unsigned)q.z < RAY_WIN)
                atomicAdd(&win[(q.z * RAY_WIN + q.y) * RAY_WIN + q.x], -1);
            else atomicAdd(&m.ray_count[gie_lidx(m, loc)], -1);      // result unused -> RED
        }
    }
    if (use_win) {
        __syncthreads();
        for (int k = threadIdx.x; k < RAY_WIN * RAY_WIN * RAY_WIN; k += blockDim.x) {
            const int v = win[k];
            if (v == 0) continue;
            const int3 loc = worg + make_int3(k % RAY_WIN, (k / RAY_WIN) % RAY_WIN, k / (RAY_WIN * RAY_WIN)) - m.pvt;
            atomicAdd(&m.ray_count[gie_lidx(m, loc)], v);            // only in-volume voxels were accumulated
        }
    }
}

// the segments that lie entirely behind their ray's stop give back what they took (a block they touched stays flagged: the
// merge looks at it and finds nothing observed)
__global__ void __launch_bounds__(128) k_pc_undo(LocDev m, int n, int max_segs, const RayCk *__restrict__ ck, const float4 *__restrict__ incs,
                                                 const int *__restrict__ nsteps, const int *__restrict__ stop)
{
    gie_pdl_sync();
    const int seg = 1 + blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;   // segment 0 is never behind a stop
    if (i >= n) return;
    const int first = seg * RAY_SEG, total = __ldg(&nsteps[i]);
    if (first >= total || first <= __ldg(&stop[i])) return;          // stop == first: the segment's own first voxel, nothing was taken
    const RayInc r = ray_inc_unpack(__ldg(&incs[i]));
    const RayCk c = ck[(size_t)seg * n + i];
    const int cnt = min(RAY_SEG, total - first);
    int3 cur = make_int3(c.x, c.y, c.z);
    float tmx = c.tmx, tmy = c.tmy, tmz = c.tmz;
    for (int j = 0; j < cnt; j++) {                                  // what pass 2 of k_pc_apply did: up to the segment's own first OCCUPIED voxel
        ray_step(r, cur, tmx, tmy, tmz);
        const int3 loc = cur - m.pvt;
        if (!gie_inside_loc(m, loc)) continue;
        const int id = gie_lidx(m, loc);
        if (m.inst_type[id] == GIE_VOX_OCCUPIED) break;
        atomicAdd(&m.ray_count[id], 1);
    }
}

// robot sphere of getAllocKeys (pntcld_raycast.cu:33-41): count = -1 inside the sphere, after the ray casting
__global__ void k_pc_sphere(LocDev m, HashDev h, int r2, int r)
{
    gie_pdl_sync();
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
    gie_pdl_sync();
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
    gie_launch(k_projective<SENSOR>, dim3(grid), dim3(block), 0, lm->stream, lm->d, hm->d, data, sp, fmp, r2);
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
        gie_launch(k_pc_register, dim3(blocks), dim3(256), 0, lm->stream, lm->d, hm->d, pts_dev, n);
        // pntcld_raycast.cu:79: 0.707f*loc_map._local_size.x*loc_map._voxel_width
        float max_len = 0.707f * (float)lm->d.X * lm->d.w;
        // a walk crosses at most (|dx| + |dy| + |dz|) <= sqrt(3) voxel borders per voxel of length: bound on the segments of a ray
        const int max_steps = (int)(0.707f * (float)lm->d.X * 1.7320508f) + 8;
        const int max_segs = (max_steps + RAY_SEG - 1) / RAY_SEG;
        const size_t need = (size_t)n * max_segs * sizeof(RayCk) + (size_t)n * (2 * sizeof(int) + sizeof(float4)) + 512;
        if (lm->ray_scratch_bytes < need) {
            if (lm->ray_scratch) { GIE_CUDA_CHECK(cudaStreamSynchronize(lm->stream)); GIE_CUDA_CHECK(cudaFree(lm->ray_scratch)); lm->ray_scratch = nullptr; }
            GIE_CUDA_CHECK(cudaMalloc(&lm->ray_scratch, need));
            lm->ray_scratch_bytes = need;
        }
        RayCk *ck = (RayCk *)lm->ray_scratch;
        float4 *incs = (float4 *)((char *)lm->ray_scratch + (((size_t)n * max_segs * sizeof(RayCk) + 127) & ~(size_t)127));
        int *nsteps = (int *)(incs + n);
        int *stop = nsteps + n;
        const dim3 grid2((n + 127) / 128, max_segs);
        gie_launch(k_pc_walk, dim3((n + 127) / 128), dim3(128), 0, lm->stream, lm->d, hm->d, pts_dev, n, max_len, max_segs, ck, incs, nsteps, stop);
        gie_launch(k_pc_apply, dim3(grid2), dim3(128), 0, lm->stream, lm->d, hm->d, n, max_segs, ck, incs, nsteps, stop);
        if (max_segs > 1) gie_launch(k_pc_undo, dim3((n + 127) / 128, max_segs - 1), dim3(128), 0, lm->stream, lm->d, n, max_segs, ck, incs, nsteps, stop);
        lm->launches += 4;
    }
    if (fmp) {
        int r = 0;
        while (r * r <= r2) r++;
        int side = 2 * r + 1, tot = side * side * side;
        gie_launch(k_pc_sphere, dim3((tot + 255) / 256), dim3(256), 0, lm->stream, lm->d, hm->d, r2, r);
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
