#pragma once
#include <pcl/point_cloud.h>
#include <ros/time.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl_conversions { inline void toPCL(const ros::Time &, std::uint64_t &) {} }
namespace pcl { template <class P> void fromROSMsg(const sensor_msgs::PointCloud2 &, PointCloud<P> &) {} }
