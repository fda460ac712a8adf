// Minimal stand-in for the ROS message header, only so that the ROS-typed code paths compile in the CPU test suite (member
// names and types as in sensor_msgs/LaserScan.msg).
#pragma once
#include <memory>
#include <vector>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct LaserScan {
    std_msgs::Header header;
    float angle_min, angle_max, angle_increment, time_increment, scan_time, range_min, range_max;
    std::vector<float> ranges, intensities;
    typedef std::shared_ptr<LaserScan const> ConstPtr;
};
}
