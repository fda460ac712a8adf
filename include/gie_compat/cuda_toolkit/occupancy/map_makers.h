// Stand-ins for the reference's four ROS-typed sensor front ends (include/cuda_toolkit/occupancy/*/xxx_map_maker.h,
// src/*_map_maker.cpp): same class and method names; every updateLocalOGM has a ROS-free overload on the raw message payload
// and, with -DGIE_COMPAT_WITH_ROS, the reference's overload on the message pointer.  No device staging buffers live here: the
// host-buffer entry points of the C ABI copy the payload once and, for the two point-cloud front ends, do the reference's
// host-side conversion loops on the device (gie_ogm_vlp16_pointcloud2_host / gie_ogm_pointcloud2_host).
#pragma once
#include <cstdint>
#include "cuda_toolkit/projection.h"
#include "cuda_toolkit/occupancy/sensor_params.h"
#include "kernel/ogm_interfaces.h"
#include "map_structure/local_batch.h"
#ifdef GIE_COMPAT_WITH_ROS
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/LaserScan.h>
#include <sensor_msgs/PointCloud2.h>
#endif

// src/hokuyo_map_maker.cpp
class HokuyoMapMaker {
public:
    void initialize(const ScanParam &p) { _scan_param = p; _initialized = true; }
    bool is_initialized() { return _initialized; }
    void setLocMap(LocMap *lMap) { _lMap = lMap; }
    // ranges = sensor_msgs/LaserScan::ranges (float[scan_num])
    void updateLocalOGM(const Projection &proj, const float *ranges, int3 * /*VB_keys_loc_D*/, const int /*time*/, bool for_motion_planner,
                        int rbt_r2_grids)
    {
        gie::use_projection(_lMap, proj);
        GIE_CHECK(gie_ogm_scan2d_host(_lMap->handle(), _lMap->_hash, ranges, _scan_param.scan_num, _scan_param.theta_inc, _scan_param.theta_min,
                                      for_motion_planner, rbt_r2_grids));
    }
#ifdef GIE_COMPAT_WITH_ROS
    void initialize(const sensor_msgs::LaserScan::ConstPtr &msg)   // hokuyo_map_maker.cpp:28-37
    {
        initialize(ScanParam((int)msg->ranges.size(), msg->range_max, msg->angle_increment, msg->angle_min));
    }
    void updateLocalOGM(const Projection &proj, const sensor_msgs::LaserScan::ConstPtr &scan, int3 *keys, const int time, bool fmp, int r2)
    {
        updateLocalOGM(proj, &scan->ranges.at(0), keys, time, fmp, r2);
    }
#endif
private:
    ScanParam _scan_param;
    bool _initialized = false;
    LocMap *_lMap = nullptr;
};

// src/realsense_map_maker.cpp
class RealsenseMapMaker {
public:
    void initialize(const CamParam &p) { _cam_param = p; _initialized = true; }
    bool is_initialized() { return _initialized; }
    void setLocMap(LocMap *lMap) { _lMap = lMap; }
    // depth = sensor_msgs/Image::data reinterpreted as float[rows * cols] metres (32FC1), as the reference does (:47-49)
    void updateLocalOGM(const Projection &proj, const float *depth, int3 *, const int, bool for_motion_planner, int rbt_r2_grids)
    {
        gie::use_projection(_lMap, proj);
        GIE_CHECK(gie_ogm_depth_host(_lMap->handle(), _lMap->_hash, depth, _cam_param.rows, _cam_param.cols, _cam_param.cx, _cam_param.cy,
                                     _cam_param.fx, _cam_param.fy, _cam_param.valid_NaN, for_motion_planner, rbt_r2_grids));
    }
#ifdef GIE_COMPAT_WITH_ROS
    void initialize(const sensor_msgs::CameraInfo::ConstPtr &msg, bool valid_NaN)   // realsense_map_maker.cpp:28-39
    {
        initialize(CamParam((int)msg->height, (int)msg->width, (float)msg->K[2], (float)msg->K[5], (float)msg->K[0], (float)msg->K[4], valid_NaN));
    }
    void updateLocalOGM(const Projection &proj, const sensor_msgs::Image::ConstPtr &img, int3 *keys, const int time, bool fmp, int r2)
    {
        updateLocalOGM(proj, (const float *)&img->data[0], keys, time, fmp, r2);
    }
#endif
private:
    CamParam _cam_param;
    bool _initialized = false;
    LocMap *_lMap = nullptr;
};

// src/pntcld_map_maker.cpp
class PntcldMapMaker {
public:
    void initialize(const PntcldParam &p) { _pnt_param = p; _initialized = true; }
    bool is_initialized() { return _initialized; }
    void setLocMap(LocMap *lMap) { _lMap = lMap; }
    // data = sensor_msgs/PointCloud2::data, point_step bytes per point, float32 x,y,z consecutive at off_x; at most cld_sz
    // points are used (pntcld_process, :49-61)
    void updateLocalOGM(const Projection &proj, const uint8_t *data, int n_points, int point_step, int off_x, int3 *, const int, bool for_motion_planner,
                        int rbt_r2_grids)
    {
        gie::use_projection(_lMap, proj);
        GIE_CHECK(gie_ogm_pointcloud2_host(_lMap->handle(), _lMap->_hash, data, n_points, point_step, off_x, _pnt_param.cld_sz, for_motion_planner,
                                           rbt_r2_grids));
        _pnt_param.valid_pnt_count = (_pnt_param.cld_sz > 0 && n_points > _pnt_param.cld_sz) ? _pnt_param.cld_sz : n_points;
    }
#ifdef GIE_COMPAT_WITH_ROS
    void initialize(const sensor_msgs::PointCloud2::ConstPtr &msg) { initialize(PntcldParam((int)(msg->width * msg->height))); }   // :36-47
    void updateLocalOGM(const Projection &proj, const sensor_msgs::PointCloud2::ConstPtr &msg, int3 *keys, const int time, bool fmp, int r2)
    {
        int off_x = 0;
        for (const auto &f : msg->fields) if (f.name == "x") off_x = (int)f.offset;
        updateLocalOGM(proj, msg->data.data(), (int)(msg->width * msg->height), (int)msg->point_step, off_x, keys, time, fmp, r2);
    }
#endif
private:
    PntcldParam _pnt_param;
    bool _initialized = false;
    LocMap *_lMap = nullptr;
};

// src/vlp16_map_maker.cpp
class Vlp16MapMaker {
public:
    void initialize(const MulScanParam &p) { _mul_scan_param = p; _initialized = true; }
    bool is_initialized() { return _initialized; }
    void setLocMap(LocMap *lMap) { _lMap = lMap; }
    // data = PointCloud2::data with float32 x / y and uint16 ring fields; binned into the 16 x 440 range image on the device
    // (convertPyntCld, :73-147) instead of the host loop + 16 per-ring copies
    void updateLocalOGM(const Projection &proj, const uint8_t *data, int n_points, int point_step, int off_x, int off_y, int off_ring, int3 *,
                        const int, bool for_motion_planner, int rbt_r2_grids)
    {
        gie::use_projection(_lMap, proj);
        GIE_CHECK(gie_ogm_vlp16_pointcloud2_host(_lMap->handle(), _lMap->_hash, data, n_points, point_step, off_x, off_y, off_ring,
                                                 _mul_scan_param.scan_num, _mul_scan_param.ring_num, _mul_scan_param.theta_inc,
                                                 _mul_scan_param.theta_min, _mul_scan_param.phi_inc, _mul_scan_param.phi_min, for_motion_planner,
                                                 rbt_r2_grids));
    }
#ifdef GIE_COMPAT_WITH_ROS
    void updateLocalOGM(const Projection &proj, const sensor_msgs::PointCloud2ConstPtr &msg, int3 *keys, const int time, bool fmp, int r2)
    {
        // field lookup with the reference's datatype checks (convertPyntCld, vlp16_map_maker.cpp:81-109): FLOAT32 = 7, UINT16 = 4
        int ox = -1, oy = -1, oi = -1, orr = -1;
        for (const auto &f : msg->fields) {
            if (f.datatype == 7) {
                if (f.name == "x") ox = (int)f.offset; else if (f.name == "y") oy = (int)f.offset; else if (f.name == "intensity") oi = (int)f.offset;
            } else if (f.datatype == 4 && f.name == "ring") orr = (int)f.offset;
        }
        // The reference bins only when x, y and ring exist AND x sits at offset 0, y at 4 and the intensity / ring offsets are
        // multiples of 4 (:111-121; it indexes the point as a float array); otherwise its scan lines stay all-INFINITY and the OGM
        // kernel still runs on them (:52-70), which marks the whole field of view FREE.  Same here: zero points give an
        // all-INFINITY range image.
        const bool binned = ox >= 0 && oy >= 0 && orr >= 0 && ox == 0 && oy == 4 && (oi % 4 == 0) && (orr % 4 == 0);
        const int n = binned ? (int)(msg->width * msg->height) : 0;
        updateLocalOGM(proj, msg->data.data(), n, (int)msg->point_step, binned ? ox : 0, binned ? oy : 0, binned ? orr : 0, keys, time, fmp, r2);
    }
#endif
private:
    MulScanParam _mul_scan_param;
    bool _initialized = false;
    LocMap *_lMap = nullptr;
};
