// The four kernel-level OGM entry points (reference: src/kernel/point_cloud/pntcld_interfaces.h:9-13,
// src/kernel/hokuyo/hokuyo_interfaces.h:9-13, src/kernel/vlp16/vlp16_interface.h:12-16,
// src/kernel/realsense/realsense_interfaces.h:9-13).  Same namespaces, names and argument lists; the sensor data pointer
// is a DEVICE pointer as in the reference.  `VB_keys_loc_D` is accepted and ignored.
//
// The reference's src/*_map_maker.cpp include "kernel/<sensor>/<sensor>_interfaces.h" with quotes, which the compiler resolves
// NEXT TO THE INCLUDING FILE first, i.e. to the reference's own declaration-only headers under src/kernel/.  A build that
// keeps those .cpp files therefore needs these four functions as linkable symbols: compile gie_compat_kernels.cpp (this
// directory's parent), which includes this header with GIE_COMPAT_EMIT_KERNEL_SYMBOLS defined.
#pragma once
#ifdef GIE_COMPAT_EMIT_KERNEL_SYMBOLS
#define GIE_KERNEL_ENTRY
#else
#define GIE_KERNEL_ENTRY inline
#endif
#include "cuda_toolkit/projection.h"
#include "cuda_toolkit/occupancy/sensor_params.h"
#include "map_structure/local_batch.h"

namespace gie {
inline void use_projection(LocMap *m, const Projection &p)
{
    const float o[3] = { p.origin.x, p.origin.y, p.origin.z };
    GIE_CHECK(gie_locmap_set_projection(m->handle(), p.L2G.data, p.G2L.data, o));
}
}  // namespace gie

namespace PNTCLD_RAYCAST {
GIE_KERNEL_ENTRY void localOGMKernels(LocMap *loc_map, float3 *pnt_cld, Projection proj, PntcldParam param, int3 * /*VB_keys_loc_D*/,
                            int /*time*/, bool for_motion_planner, int rbt_r2_grids)
{
    gie::use_projection(loc_map, proj);
    GIE_CHECK(gie_ogm_pointcloud_dev(loc_map->handle(), loc_map->_hash, &pnt_cld->x, param.valid_pnt_count, for_motion_planner, rbt_r2_grids));
}
}
namespace HOKUYO_FAST {
GIE_KERNEL_ENTRY void localOGMKernels(LocMap *loc_map, SCAN_DEPTH_TPYE *detph_data, Projection proj, ScanParam param, int3 *, bool for_motion_planner,
                            int rbt_r2_grids)
{
    gie::use_projection(loc_map, proj);
    GIE_CHECK(gie_ogm_scan2d_dev(loc_map->handle(), loc_map->_hash, detph_data, param.scan_num, param.theta_inc, param.theta_min,
                                 for_motion_planner, rbt_r2_grids));
}
}
namespace VLP_FAST {
GIE_KERNEL_ENTRY void localOGMKernels(LocMap *loc_map, SCAN_DEPTH_TPYE *detph_data, Projection proj, MulScanParam param, int3 *, bool for_motion_planner,
                            int rbt_r2_grids)
{
    gie::use_projection(loc_map, proj);
    GIE_CHECK(gie_ogm_vlp16_dev(loc_map->handle(), loc_map->_hash, detph_data, param.scan_num, param.ring_num, param.theta_inc,
                                param.theta_min, param.phi_inc, param.phi_min, for_motion_planner, rbt_r2_grids));
}
}
namespace REALSENSE_FAST {
GIE_KERNEL_ENTRY void localOGMKernels(LocMap *loc_map, REALSENSE_DEPTH_TPYE *detph_data, Projection proj, CamParam param, int3 *, bool for_motion_planner,
                            int rbt_r2_grids)
{
    gie::use_projection(loc_map, proj);
    GIE_CHECK(gie_ogm_depth_dev(loc_map->handle(), loc_map->_hash, detph_data, param.rows, param.cols, param.cx, param.cy, param.fx,
                                param.fy, param.valid_NaN, for_motion_planner, rbt_r2_grids));
}
}
