#pragma once
#include <cstdint>
#include <memory>
namespace sensor_msgs {
struct CameraInfo {
    uint32_t height, width;
    double K[9];
    typedef std::shared_ptr<CameraInfo const> ConstPtr;
};
}
