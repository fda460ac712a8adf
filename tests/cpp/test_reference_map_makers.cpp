// Links the reference's OWN, unmodified src/{hokuyo,realsense,pntcld,vlp16}_map_maker.cpp (compiled against include/gie_compat
// with -DGIE_COMPAT_REFERENCE_MAPMAKERS and the ROS message stand-ins of tests/cpp/ros_stubs) into one program with the C ABI
// library.  Without a GPU it only checks the host-side behaviour that needs no device: construction, the ROS-typed
// parameter extraction, and that every symbol resolves.  With --run (GPU box) it integrates one frame per sensor through the
// reference's map makers and prints a checksum of the resulting occupancy for the parity test.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "cuda_toolkit/occupancy/hokuyo/hokuyo_map_maker.h"
#include "cuda_toolkit/occupancy/realsense/realsense_map_maker.h"
#include "cuda_toolkit/occupancy/point_cloud/pntcld_map_maker.h"
#include "cuda_toolkit/occupancy/vlp16/vlp16_map_maker.h"
#include "par_wave/glb_hash_map.h"

static unsigned long long checksum(LocMap &lm)
{
    lm.copy_ogm_2_host();
    unsigned long long h = 1469598103934665603ULL;
    const int n = lm._local_size.x * lm._local_size.y * lm._local_size.z;
    for (int i = 0; i < n; i++) { h ^= (unsigned char)lm.glb_type_H[i]; h *= 1099511628211ULL; }
    return h;
}

int main(int argc, char **argv)
{
    const bool run = argc > 1 && !strcmp(argv[1], "--run");
    {
        HokuyoMapMaker h; RealsenseMapMaker r; PntcldMapMaker p; Vlp16MapMaker v;
        if (h.is_initialized() || r.is_initialized() || p.is_initialized() || v.is_initialized()) return 1;
    }
    if (!run) { printf("reference map makers link OK\n"); return 0; }

    const int3 size = make_int3(64, 64, 32);
    Projection proj;
    {
        const float q[4] = { 1.f, 0.f, 0.f, 0.f }, t[3] = { 0.3f, -0.2f, 1.5f };
        gie_make_projection(q, t, proj.L2G.data, proj.G2L.data);
        proj.origin = make_float3(t[0], t[1], t[2]);
    }
    // --- 2-D scan through the reference's HokuyoMapMaker
    {
        LocMap lm(0.2f, size, 180, -10.f, 10.f, 100, true);
        lm.create_gpu_map();
        GlbHashMap hm(lm._bdr_num, lm._local_size, 4000, 12000);
        hm.setLocMap(&lm);
        auto scan = std::make_shared<sensor_msgs::LaserScan>();
        scan->ranges.assign(1081, 3.0f); scan->range_max = 30.f; scan->angle_increment = 0.25f * 3.14159265f / 180.f; scan->angle_min = -135.f * 3.14159265f / 180.f;
        HokuyoMapMaker mk;
        mk.setLocMap(&lm);
        mk.initialize(sensor_msgs::LaserScan::ConstPtr(scan));
        lm.calculate_pivot_origin(proj.origin); lm.calculate_update_pivot(proj.origin);
        mk.updateLocalOGM(proj, scan, thrust::raw_pointer_cast(hm.VB_keys_loc_D.data()), 1, false, 0);
        hm.updateHashOGM(false, 1, false, nullptr);
        printf("scan2d %llu\n", checksum(lm));
    }
    // --- point cloud through the reference's PntcldMapMaker (x, y, z float32 at offsets 0, 4, 8, point_step 16)
    {
        LocMap lm(0.1f, size, 180, -10.f, 10.f, 64, false);
        lm.create_gpu_map();
        GlbHashMap hm(lm._bdr_num, lm._local_size, 4000, 12000);
        hm.setLocMap(&lm);
        auto pc = std::make_shared<sensor_msgs::PointCloud2>();
        pc->height = 1; pc->width = 2000; pc->point_step = 16;
        pc->fields = { { "x", 0, sensor_msgs::PointField::FLOAT32, 1 }, { "y", 4, sensor_msgs::PointField::FLOAT32, 1 },
                       { "z", 8, sensor_msgs::PointField::FLOAT32, 1 } };
        pc->data.resize((size_t)pc->width * 16);
        for (unsigned i = 0; i < pc->width; i++) {
            volatile float fx = 0.002f * (float)i, fz = 0.05f * (float)(i % 7);   // single roundings, no contraction
            float p[4] = { -2.0f + fx, 1.5f, 0.25f + fz, 0.f };
            memcpy(&pc->data[(size_t)i * 16], p, 16);
        }
        PntcldMapMaker mk;
        mk.setLocMap(&lm);
        mk.initialize(sensor_msgs::PointCloud2::ConstPtr(pc));
        lm.calculate_pivot_origin(proj.origin); lm.calculate_update_pivot(proj.origin);
        mk.updateLocalOGM(proj, pc, thrust::raw_pointer_cast(hm.VB_keys_loc_D.data()), 1, false, 0);
        hm.updateHashOGM(true, 1, false, nullptr);
        printf("pointcloud %llu\n", checksum(lm));
    }
    printf("reference map makers run OK\n");
    return 0;
}
