#pragma once
#include <pcl/point_cloud.h>
namespace pcl { template <class P> struct KdTreeFLANN {
    void setInputCloud(const typename PointCloud<P>::Ptr &) {}
    int nearestKSearch(const P &, int, std::vector<int> &, std::vector<float> &) { return 0; }
}; }
