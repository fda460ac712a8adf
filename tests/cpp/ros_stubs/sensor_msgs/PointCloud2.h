#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace sensor_msgs {
struct PointField { std::string name; uint32_t offset; uint8_t datatype; uint32_t count; };
struct PointCloud2 {
    uint32_t height, width, point_step, row_step;
    std::vector<PointField> fields;
    std::vector<uint8_t> data;
    typedef std::shared_ptr<PointCloud2 const> ConstPtr;
};
typedef std::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
}
