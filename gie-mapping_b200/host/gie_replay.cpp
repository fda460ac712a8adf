// gie_replay — ROS-free C++ host driver: replays VOLMAPNODE's constructor and publishMap call order
// (reference src/volumetric_mapper.cpp:70-126 and :138-224) on recorded sensor frames, written against the reference's
// own operator surface (LocMap, GlbHashMap, EDT_OCC::batchEDTUpdate, *::localOGMKernels, Ext_Obs_Wrapper, warmupCuda,
// cuttPlan) as provided by include/gie_compat/ over the C ABI of libgie_b200.so.  Plain C++17, built with g++.
//
//   gie_replay <frames.bin> <out.bin|-> [--time] [--stream] [--costmap] [--mapmakers]
//
// --mapmakers routes the sensor data through the node-level front ends (PntcldMapMaker / HokuyoMapMaker / RealsenseMapMaker
// ::updateLocalOGM on HOST payloads, as VOLMAPNODE does at volumetric_mapper.cpp:158-176) instead of calling
// XXX::localOGMKernels on a device buffer.
//
// frames.bin (little endian; writer: gie-mapping_b200/replay_io.py):
//   header  int32[18] {magic 'GIE1', sensor, X, Y, Z, occupancy_threshold, cutoff_grids_sq, fast_mode, bucket_max,
//                      block_max, for_motion_planner, robot_r2_grids, nframes, rows, cols, scan_num, ring_num, valid_NaN}
//           float[11] {voxel_width, ogm_min_h, ogm_max_h, theta_inc, theta_min, phi_inc, phi_min, cx, cy, fx, fy}
//   frame   float[4] q(w,x,y,z), float[3] t, int32 n, float[n] payload (points xyz | scan | ranges | depth image)
// out.bin: per frame glb_type i8[N], aux i32[N], coc_aux i32[N], pair u64[N], edt f32[N]; with --costmap additionally
//   SeenDist[N]; with --stream, after the last frame: int32 nblocks, then per block int32 key[3] + GlbVoxel[512] from the
//   host mirror that streamPipeline maintains (VB_keys_H / VB_values_H).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime_api.h>

#include "map_structure/local_batch.h"
#include "cuda_toolkit/projection.h"
#include "cuda_toolkit/edt/edt_interfaces.h"
#include "par_wave/glb_hash_map.h"
#include "map_structure/pre_map.h"
#include "kernel/point_cloud/pntcld_interfaces.h"
#include "kernel/hokuyo/hokuyo_interfaces.h"
#include "kernel/vlp16/vlp16_interface.h"
#include "kernel/realsense/realsense_interfaces.h"
#include "cuda_toolkit/occupancy/point_cloud/pntcld_map_maker.h"
#include "cuda_toolkit/occupancy/hokuyo/hokuyo_map_maker.h"
#include "cuda_toolkit/occupancy/realsense/realsense_map_maker.h"
#include "cuda_toolkit/occupancy/vlp16/vlp16_map_maker.h"

namespace {
struct Header {
    int magic, sensor, X, Y, Z, thresh, cutoff_sq, fast, bucket_max, block_max, fmp, r2, nframes;
    int rows, cols, scan_num, ring_num, valid_nan;
    float w, min_h, max_h, theta_inc, theta_min, phi_inc, phi_min, cx, cy, fx, fy;
};
struct Frame { float pose[7]; std::vector<float> payload; };

// VOLMAPNODE::setupRotationPlan (volumetric_mapper.cpp:344-373): kept so the call order is complete; plans are no-ops here
void setup_rotation_plan(const LocMap &m, cuttHandle plan[3])
{
    int dx = m._local_size.x, dy = m._local_size.y, dz = m._local_size.z;
    int dims[3][3] = { { dx, dy, dz }, { dy, dx, dz }, { dy, dz, dx } };
    int perm[3][3] = { { 1, 0, 2 }, { 0, 2, 1 }, { 2, 0, 1 } };
    for (int i = 0; i < 3; i++) cuttCheck(cuttPlan(&plan[i], 3, dims[i], perm[i], sizeof(int), nullptr));
}
template <typename T>
void dump(FILE *fo, LocMap &m, int which, std::vector<T> &buf)
{
    GIE_CHECK(gie_locmap_download(m.handle(), which, buf.data()));
    fwrite(buf.data(), sizeof(T), buf.size(), fo);
}
}  // namespace

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s frames.bin out.bin|- [--time] [--stream] [--costmap]\n", argv[0]); return 2; }
    bool timing = false, stream = false, costmap = false, mapmakers = false;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--time")) timing = true;
        else if (!strcmp(argv[i], "--stream")) stream = true;
        else if (!strcmp(argv[i], "--costmap")) costmap = true;
        else if (!strcmp(argv[i], "--mapmakers")) mapmakers = true;
    }
    FILE *fi = fopen(argv[1], "rb");
    Header h;
    if (!fi || fread(&h, sizeof(h), 1, fi) != 1 || h.magic != 0x47494531) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    std::vector<Frame> frames(h.nframes);
    size_t max_payload = 4;
    for (auto &f : frames) {
        int n = 0;
        if (fread(f.pose, 4, 7, fi) != 7 || fread(&n, 4, 1, fi) != 1) { fprintf(stderr, "short file\n"); return 2; }
        f.payload.resize(n);
        if (n && fread(f.payload.data(), 4, n, fi) != (size_t)n) { fprintf(stderr, "short payload\n"); return 2; }
        if ((size_t)n > max_payload) max_payload = n;
    }
    fclose(fi);
    FILE *fo = strcmp(argv[2], "-") ? fopen(argv[2], "wb") : nullptr;

    try {
        // ---- VOLMAPNODE::VOLMAPNODE (volumetric_mapper.cpp:70-126) -------------------------------------------------------
        LocMap *loc_map = new LocMap(h.w, make_int3(h.X, h.Y, h.Z), (unsigned char)h.thresh, h.min_h, h.max_h, h.cutoff_sq, h.fast != 0);
        loc_map->create_gpu_map();
        cuttHandle rotation_plan[3];
        setup_rotation_plan(*loc_map, rotation_plan);
        GlbHashMap *hash_map = new GlbHashMap(loc_map->_bdr_num, loc_map->_local_size, h.bucket_max, h.block_max);
        hash_map->setLocMap(loc_map);
        Ext_Obs_Wrapper *ext_obs = new Ext_Obs_Wrapper(1);
        PntcldMapMaker pnt_map_maker;
        HokuyoMapMaker hok_map_maker;
        RealsenseMapMaker rea_map_maker;
        pnt_map_maker.setLocMap(loc_map); hok_map_maker.setLocMap(loc_map); rea_map_maker.setLocMap(loc_map);
        pnt_map_maker.initialize(PntcldParam((int)(max_payload / 3)));
        hok_map_maker.initialize(ScanParam(h.scan_num, 30.f, h.theta_inc, h.theta_min));
        rea_map_maker.initialize(CamParam(h.rows, h.cols, h.cx, h.cy, h.fx, h.fy, h.valid_nan != 0));
        warmupCuda();

        float *sensor_dev = nullptr;   // the MapMakers' device buffer (e.g. PntcldMapMaker::_gpu_cld)
        if (cudaMalloc((void **)&sensor_dev, max_payload * sizeof(float)) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
        const size_t N = (size_t)loc_map->_map_volume;
        std::vector<signed char> b8(N); std::vector<int> b32(N); std::vector<unsigned long long> b64(N); std::vector<float> bf(N);
        const bool display_glb_edt = stream, display_glb_ogm = false;
        double tot_ogm = 0, tot_edt = 0;
        int time = 0;

        for (auto &f : frames) {
            // ---- VOLMAPNODE::publishMap (volumetric_mapper.cpp:138-224) ---------------------------------------------------
            time++;
            auto t0 = std::chrono::steady_clock::now();
            Projection proj = make_projection(f.pose[0], f.pose[1], f.pose[2], f.pose[3], f.pose[4], f.pose[5], f.pose[6]);
            loc_map->calculate_pivot_origin(proj.origin);
            loc_map->calculate_update_pivot(proj.origin);
            int3 *keys = hash_map->VB_keys_loc_D.data();
            const int n = (int)f.payload.size();
            const bool via_maker = mapmakers && h.sensor != 2;   // the VLP-16 maker wants the raw cloud, the frame file holds range images
            if (n && !via_maker) cudaMemcpy(sensor_dev, f.payload.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
            if (via_maker && h.sensor == 0) {
                pnt_map_maker.updateLocalOGM(proj, (const uint8_t *)f.payload.data(), n / 3, 12, 0, keys, time, h.fmp != 0, h.r2);
            } else if (via_maker && h.sensor == 1) {
                hok_map_maker.updateLocalOGM(proj, f.payload.data(), keys, time, h.fmp != 0, h.r2);
            } else if (via_maker && h.sensor == 3) {
                rea_map_maker.updateLocalOGM(proj, f.payload.data(), keys, time, h.fmp != 0, h.r2);
            } else if (h.sensor == 0) {
                PntcldParam pp(n / 3);
                pp.valid_pnt_count = n / 3;
                PNTCLD_RAYCAST::localOGMKernels(loc_map, (float3 *)sensor_dev, proj, pp, keys, time, h.fmp != 0, h.r2);
            } else if (h.sensor == 1) {
                HOKUYO_FAST::localOGMKernels(loc_map, sensor_dev, proj, ScanParam(h.scan_num, 30.f, h.theta_inc, h.theta_min), keys, h.fmp != 0, h.r2);
            } else if (h.sensor == 2) {
                VLP_FAST::localOGMKernels(loc_map, sensor_dev, proj,
                                          MulScanParam(h.scan_num, h.ring_num, 10.f, h.theta_inc, h.theta_min, h.phi_inc, h.phi_min), keys, h.fmp != 0, h.r2);
            } else {
                REALSENSE_FAST::localOGMKernels(loc_map, sensor_dev, proj, CamParam(h.rows, h.cols, h.cx, h.cy, h.fx, h.fy, h.valid_nan != 0), keys,
                                                h.fmp != 0, h.r2);
            }
            float3 ll = loc_map->_msg_origin, ur = make_float3(ll.x + h.X * h.w, ll.y + h.Y * h.w, ll.z + h.Z * h.w);
            ext_obs->activate_AABB(ll, ur);   // VOLMAPNODE::update_ext_map (:498-508)
            hash_map->updateHashOGM(h.sensor == 0, time, display_glb_ogm && !display_glb_edt, ext_obs);
            if (timing) cudaDeviceSynchronize();   // GPU_DEV_SYNC "only for profiling" (:186)
            auto t1 = std::chrono::steady_clock::now();
            EDT_OCC::batchEDTUpdate(loc_map, rotation_plan, time);
            hash_map->mergeNewObsv(time, display_glb_edt);
            if (display_glb_edt || display_glb_ogm) hash_map->streamPipeline();
            if (h.fmp || costmap) loc_map->convertCostMap();
            hash_map->sync();
            auto t2 = std::chrono::steady_clock::now();
            double ogm_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
            double edt_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
            tot_ogm += ogm_ms; tot_edt += edt_ms;
            if (timing) printf("frame %d ogm_ms %.4f edt_ms %.4f\n", time - 1, ogm_ms, edt_ms);
            if (fo) {
                dump(fo, *loc_map, GIE_ARR_GLB_TYPE, b8);
                dump(fo, *loc_map, GIE_ARR_AUX, b32);
                dump(fo, *loc_map, GIE_ARR_COC_AUX, b32);
                dump(fo, *loc_map, GIE_ARR_PAIR, b64);
                dump(fo, *loc_map, GIE_ARR_EDT, bf);
                if (costmap) fwrite(loc_map->seendist_out, sizeof(SeenDist), N, fo);
            }
        }
        if (fo && stream) {
            int nb = hash_map->VB_cnt_H;
            fwrite(&nb, 4, 1, fo);
            for (int i = 0; i < nb; i++) {
                fwrite(&hash_map->VB_keys_H[i], sizeof(int3), 1, fo);
                fwrite(&hash_map->VB_values_H[i], sizeof(VoxelBlock), 1, fo);
            }
        }
        printf("total frames %d ogm_ms %.4f edt_ms %.4f\n", h.nframes, tot_ogm, tot_edt);
        cudaFree(sensor_dev);
        delete ext_obs;
        delete hash_map;
        delete loc_map;
    } catch (const gie::Error &e) {
        fprintf(stderr, "gie_replay: %s\n", e.what());
        return 1;
    }
    if (fo) fclose(fo);
    return 0;
}
