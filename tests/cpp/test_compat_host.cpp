// Host-only checks of the reference-named C++ surface (include/gie_compat): nothing here touches the GPU, so it runs in the
// CPU test suite.  Values are checked against the formulas of the reference headers they stand in for (cited inline).
#include <cassert>
#include <cmath>
#include <cstdio>
#include <vector>

#include "map_structure/local_batch.h"
#include "par_wave/glb_hash_map.h"
#include "map_structure/pre_map.h"
#include "cutt/cutt.h"
#include "cuda_toolkit/occupancy/point_cloud/pntcld_param.h"
#include "cuda_toolkit/occupancy/hokuyo/scan_param.h"
#include "cuda_toolkit/occupancy/vlp16/multiscan_param.h"
#include "cuda_toolkit/occupancy/realsense/camera_param.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main()
{
    // LocMap bookkeeping (local_batch.h:35-60): derived sizes, index and coordinate helpers (:249-301, :393-407)
    LocMap m(0.2f, make_int3(128, 96, 32), 180, -10.f, 10.f, 100, true);
    CHECK(m._map_volume == 128 * 96 * 32);
    CHECK(m._max_width == 128 + 96 + 32);
    CHECK(m._max_loc_dist_sq == 128 * 128 + 96 * 96 + 32 * 32);
    CHECK(m._bdr_num == 2 * (128 * 96 + 96 * 32 + 128 * 32));
    CHECK(m._half_shift.x == 64 && m._half_shift.y == 48 && m._half_shift.z == 16);
    CHECK(m.seendist_size == m._map_volume * (int)sizeof(SeenDist) && sizeof(SeenDist) == 8);
    CHECK(m.id(3, 4, 5) == 5 * 128 * 96 + 4 * 128 + 3);
    CHECK(m.coord2idx_local(make_int3(3, 4, 5)) == m.id(3, 4, 5));
    int3 c = m.pos2coord(make_float3(0.29f, -0.31f, 1.0f));          // floor(p / w + 0.5)
    CHECK(c.x == 1 && c.y == -2 && c.z == 5);
    float3 p = m.coord2pos(make_int3(-3, 0, 7));
    CHECK(std::fabs(p.x + 0.6f) < 1e-6f && p.y == 0.f && std::fabs(p.z - 1.4f) < 1e-6f);
    CHECK(m.is_inside_local_volume(make_int3(127, 95, 31)) && !m.is_inside_local_volume(make_int3(128, 0, 0)) &&
          !m.is_inside_local_volume(make_int3(0, -1, 0)));

    // voxel-block addressing (voxmap_utils.cuh:93-109): floor division by 8, reference voxel order x*64 + y*8 + z
    int3 k = get_VB_key(make_int3(-1, 8, 17));
    CHECK(k.x == -1 && k.y == 1 && k.z == 2);
    CHECK(get_voxID_in_VB(make_int3(-1, 8, 17)) == 7 * 64 + 0 * 8 + 1);
    int3 r = reconstruct_vox_crd(make_int3(-8, 8, 16), 7 * 64 + 0 * 8 + 1);
    CHECK(r.x == -1 && r.y == 8 && r.z == 17);
    CHECK(sizeof(GlbVoxel) == 40 && sizeof(VoxelBlock) == 40 * 512);
    GlbVoxel v;
    CHECK(v.vox_type == VOXTYPE_UNKNOWN && v.dist_sq == EMPTY_VALUE && v.coc_glb.x == EMPTY_VALUE && v.wave_layer == -1);
    CHECK(invalid_blk_key(EMPTY_KEY) && !invalid_blk_key(make_int3(0, 0, 0)));
    BlockHasher hsh; CrdEqualTo eq; CrdLessThan lt;
    CHECK(hsh(make_int3(1, 2, 3)) == (((size_t)1 * 73856093u) ^ ((size_t)2 * 19349669u) ^ ((size_t)3 * 83492791u)));
    CHECK(eq(make_int3(1, 2, 3), make_int3(1, 2, 3)) && lt(make_int3(1, 2, 3), make_int3(1, 3, 0)) && !lt(make_int3(2, 0, 0), make_int3(1, 9, 9)));

    // external-obstacle boxes (pre_map.cu:80-101): box 0 is never activated here, boxes that intersect the volume are
    Ext_Obs_Wrapper obs(1);
    float3 ll1 = make_float3(1, 1, 0), ur1 = make_float3(2, 2, 1), ll2 = make_float3(50, 50, 0), ur2 = make_float3(51, 51, 1);
    obs.append_new_elem(ll1, ur1);
    obs.append_new_elem(ll2, ur2);
    float3 vol_ll = make_float3(-5, -5, -1), vol_ur = make_float3(5, 5, 3);
    obs.activate_AABB(vol_ll, vol_ur);
    CHECK(obs.ext_obs_num == 3 && obs.obs_activated[0] == 0 && obs.obs_activated[1] == 1 && obs.obs_activated[2] == 0);
    CHECK(obs.CheckAABBIntersection(ll1, ur1, vol_ll, vol_ur) && !obs.CheckAABBIntersection(ll2, ur2, vol_ll, vol_ur));

    // cuTT plans are handles only (volumetric_mapper.cpp:344-373)
    cuttHandle plan[3];
    int dim[3] = { 128, 96, 32 }, perm[3] = { 1, 0, 2 };
    CHECK(cuttPlan(&plan[0], 3, dim, perm, sizeof(int), nullptr) == CUTT_SUCCESS && cuttPlan(nullptr, 3, dim, perm, 4, nullptr) != CUTT_SUCCESS);
    CHECK(cuttDestroy(plan[0]) == CUTT_SUCCESS);

    // sensor parameter PODs keep the reference's constructor argument order
    ScanParam sp(1081, 30.f, 0.25f, -2.35f);
    CHECK(sp.scan_num == 1081 && sp.max_r == 30.f && sp.theta_inc == 0.25f && sp.theta_min == -2.35f);
    MulScanParam mp(440, 16, 10.f, 0.014f, -3.14f, 0.035f, -0.26f);
    CHECK(mp.scan_num == 440 && mp.ring_num == 16 && mp.max_r == 10.f && mp.phi_inc == 0.035f && mp.phi_min == -0.26f);
    CamParam cp(480, 640, 320.5f, 240.5f, 554.26f, 554.26f, true);
    CHECK(cp.rows == 480 && cp.cols == 640 && cp.cx == 320.5f && cp.fy == 554.26f && cp.valid_NaN);
    PntcldParam pp(65536);
    CHECK(pp.cld_sz == 65536 && pp.valid_pnt_count == 0);

    // projection from a pose (projection.h:15-33; se3.cuh:47-75,89-105): pure host math inside the C ABI library
    Projection proj = make_projection(0.70710678f, 0.f, 0.f, 0.70710678f, 1.f, 2.f, 3.f);   // 90 deg about z
    CHECK(std::fabs(proj.L2G.data[0]) < 1e-6f && std::fabs(proj.L2G.data[1] + 1.f) < 1e-6f && std::fabs(proj.L2G.data[4] - 1.f) < 1e-6f);
    CHECK(proj.L2G.data[3] == 1.f && proj.L2G.data[7] == 2.f && proj.L2G.data[11] == 3.f && proj.origin.z == 3.f);
    // G2L * L2G = identity on a point
    float q[3] = { 0.3f, -1.2f, 2.5f }, g[3], b[3];
    for (int i = 0; i < 3; i++) g[i] = proj.L2G.data[4 * i] * q[0] + proj.L2G.data[4 * i + 1] * q[1] + proj.L2G.data[4 * i + 2] * q[2] + proj.L2G.data[4 * i + 3];
    for (int i = 0; i < 3; i++) b[i] = proj.G2L.data[4 * i] * g[0] + proj.G2L.data[4 * i + 1] * g[1] + proj.G2L.data[4 * i + 2] * g[2] + proj.G2L.data[4 * i + 3];
    CHECK(std::fabs(b[0] - q[0]) < 1e-5f && std::fabs(b[1] - q[1]) < 1e-5f && std::fabs(b[2] - q[2]) < 1e-5f);
    std::printf("compat host checks OK\n");
    return 0;
}
