// Declaration-only form of the four MapMaker classes, for builds that keep the reference's OWN src/*_map_maker.cpp
// (define GIE_COMPAT_REFERENCE_MAPMAKERS): class names, method signatures and data members as in the reference's
// include/cuda_toolkit/occupancy/{hokuyo,realsense,point_cloud,vlp16}/*_map_maker.h, so that those four .cpp files compile
// UNCHANGED against this tree — their device staging buffers (GPU_MALLOC / GPU_MEMCPY_H2D, cuda_macro.h) and host loops stay,
// and the *::localOGMKernels calls beneath them land in the C ABI (kernel/ogm_interfaces.h).
// tests/test_host_cpu.py::test_reference_map_makers_compile_unchanged builds exactly that.
#pragma once
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/LaserScan.h>
#include <sensor_msgs/PointCloud2.h>
#include "cuda_toolkit/projection.h"
#include "cuda_toolkit/occupancy/sensor_params.h"
#include "map_structure/local_batch.h"

class HokuyoMapMaker {   // hokuyo_map_maker.h:9-30
public:
    HokuyoMapMaker();
    ~HokuyoMapMaker();
    void initialize(const ScanParam &p);
    void initialize(const sensor_msgs::LaserScan::ConstPtr &msg);
    bool is_initialized() { return _initialized; }
    void setLocMap(LocMap *lMap);
    void updateLocalOGM(const Projection &proj, const sensor_msgs::LaserScan::ConstPtr &scan, int3 *VB_keys_loc_D, const int time,
                        bool for_motion_planner, int rbt_r2_grids);
private:
    ScanParam _scan_param;
    int _scan_byte_sz;
    SCAN_DEPTH_TPYE *_gpu_scan;
    bool _initialized = false;
    LocMap *_lMap;
};

class RealsenseMapMaker {   // realsense_map_maker.h:10-27
public:
    RealsenseMapMaker();
    ~RealsenseMapMaker();
    void initialize(const CamParam &p);
    void initialize(const sensor_msgs::CameraInfo::ConstPtr &msg, bool valid_NaN);
    void setLocMap(LocMap *lMap);
    void updateLocalOGM(const Projection &proj, const sensor_msgs::Image::ConstPtr &dep_img, int3 *VB_keys_loc_D, const int time,
                        bool for_motion_planner, int rbt_r2_grids);
    bool is_initialized() { return _initialized; }
private:
    CamParam _cam_param;
    LocMap *_lMap;
    int _img_byte_sz;
    REALSENSE_DEPTH_TPYE *_gpu_dep_img;
    bool _initialized = false;
};

class PntcldMapMaker {   // pntcld_map_maker.h:9-31
public:
    PntcldMapMaker();
    ~PntcldMapMaker();
    void initialize(const PntcldParam &p);
    void initialize(const sensor_msgs::PointCloud2::ConstPtr &msg);
    void setLocMap(LocMap *lMap);
    void updateLocalOGM(const Projection &proj, const sensor_msgs::PointCloud2::ConstPtr &msg, int3 *VB_keys_loc_D, const int time,
                        bool for_motion_planner, int rbt_r2_grids);
    bool is_initialized() { return _initialized; }
    void pntcld_process(const sensor_msgs::PointCloud2ConstPtr &msg);
private:
    PntcldParam _pnt_param;
    LocMap *_lMap;
    int _cld_byte_sz;
    PNT_TYPE *_gpu_cld;
    PNT_TYPE *_cpu_cld;
    bool _initialized = false;
};

class Vlp16MapMaker {   // vlp16_map_maker.h:11-33
public:
    Vlp16MapMaker();
    ~Vlp16MapMaker();
    void initialize(const MulScanParam &p);
    bool is_initialized() { return _initialized; }
    void setLocMap(LocMap *lMap);
    void updateLocalOGM(const Projection &proj, const sensor_msgs::PointCloud2ConstPtr &pyntcld, int3 *VB_keys_loc_D, const int time,
                        bool for_motion_planner, int rbt_r2_grids);
    void convertPyntCld(const sensor_msgs::PointCloud2ConstPtr &msg);
private:
    MulScanParam _mul_scan_param;
    int _range_byte_sz;
    SCAN_DEPTH_TPYE *_gpu_mulscan;
    bool _initialized = false;
    sensor_msgs::LaserScan scanlines[16];
    const int rayid_toup[16] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 };
    LocMap *_lMap;
};
