#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace sensor_msgs {
struct Image {
    uint32_t height, width, step;
    std::string encoding;
    uint8_t is_bigendian;
    std::vector<uint8_t> data;
    typedef std::shared_ptr<Image const> ConstPtr;
};
}
