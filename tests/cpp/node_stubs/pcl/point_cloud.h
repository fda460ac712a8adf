#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include <pcl/point_types.h>
#include <Eigen/Dense>   // PCL drags Eigen in; include/parameters.h relies on that
namespace pcl {
struct PCLHeader { std::uint32_t seq = 0; std::uint64_t stamp = 0; std::string frame_id; };
struct PointIndices { PCLHeader header; std::vector<int> indices; };
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    PCLHeader header;
    std::vector<P> points;
    std::uint32_t width = 0, height = 0;
    void push_back(const P &p) { points.push_back(p); }
    size_t size() const { return points.size(); }
    void clear() { points.clear(); }
    P &operator[](size_t i) { return points[i]; }
    const P &operator[](size_t i) const { return points[i]; }
};
template <class A, class B> void copyPointCloud(const PointCloud<A> &in, PointCloud<B> &out)
{
    out.header = in.header; out.points.resize(in.points.size());
    for (size_t i = 0; i < in.points.size(); i++) { out.points[i].x = in.points[i].x; out.points[i].y = in.points[i].y; out.points[i].z = in.points[i].z; }
}
}
