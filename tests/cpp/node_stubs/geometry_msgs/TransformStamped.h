#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; };
struct TransformStamped { std_msgs::Header header; Transform transform; typedef std::shared_ptr<TransformStamped const> ConstPtr; };
}
