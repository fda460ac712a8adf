#pragma once
#include <pcl/point_cloud.h>
namespace pcl { namespace search { template <class P> struct KdTree {
    typedef std::shared_ptr<KdTree<P>> Ptr;
    void setInputCloud(const typename PointCloud<P>::Ptr &) {}
    int radiusSearch(int, double, std::vector<int> &, std::vector<float> &) { return 0; }
}; } }
