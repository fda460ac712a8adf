#pragma once
#include <pcl/point_cloud.h>
namespace pcl { template <class P> struct MomentOfInertiaEstimation {
    void setInputCloud(const typename PointCloud<P>::Ptr &) {}
    void compute() {}
    bool getAABB(P &, P &) { return true; }
}; }
