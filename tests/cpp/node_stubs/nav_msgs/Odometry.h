#pragma once
#include <memory>
#include <geometry_msgs/TransformStamped.h>
namespace nav_msgs { struct Odometry { std_msgs::Header header; geometry_msgs::PoseWithCovariance pose; typedef std::shared_ptr<Odometry const> ConstPtr; }; }
