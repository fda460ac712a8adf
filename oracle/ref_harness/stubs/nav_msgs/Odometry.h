// Stub so that the reference's include/cuda_toolkit/projection.h:4 compiles without ROS.  Test infrastructure only.
#pragma once
namespace nav_msgs { struct Odometry {}; }
