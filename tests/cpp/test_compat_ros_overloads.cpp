// Compile-only: the ROS-typed overloads of the compat MapMakers and trans2proj against minimal message / tf stand-ins.
#include "cuda_toolkit/occupancy/point_cloud/pntcld_map_maker.h"
#include "cuda_toolkit/occupancy/vlp16/vlp16_map_maker.h"
#include "cuda_toolkit/occupancy/hokuyo/hokuyo_map_maker.h"
#include "cuda_toolkit/occupancy/realsense/realsense_map_maker.h"
#include "par_wave/glb_hash_map.h"

void node_like(LocMap *lm, GlbHashMap *hm, const tf::Transform &trans, const sensor_msgs::LaserScan::ConstPtr &scan,
               const sensor_msgs::Image::ConstPtr &img, const sensor_msgs::CameraInfo::ConstPtr &info,
               const sensor_msgs::PointCloud2::ConstPtr &cloud)
{
    Projection proj = trans2proj(trans);
    int3 *keys = thrust::raw_pointer_cast(hm->VB_keys_loc_D.data());
    HokuyoMapMaker hok; hok.setLocMap(lm); hok.initialize(scan); hok.updateLocalOGM(proj, scan, keys, 1, false, 0);
    RealsenseMapMaker rea; rea.setLocMap(lm); rea.initialize(info, true); rea.updateLocalOGM(proj, img, keys, 1, false, 0);
    PntcldMapMaker pnt; pnt.setLocMap(lm); pnt.initialize(cloud); pnt.updateLocalOGM(proj, cloud, keys, 1, false, 0);
    Vlp16MapMaker vlp; vlp.setLocMap(lm); vlp.initialize(MulScanParam(440, 16, 10.f, 0.0143f, -3.1416f, 0.0349f, -0.2618f));
    vlp.updateLocalOGM(proj, cloud, keys, 1, false, 0);
}
int main() { return 0; }
