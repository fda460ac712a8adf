#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct PointField {
    enum { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
    std::string name; uint32_t offset; uint8_t datatype; uint32_t count;
};
struct PointCloud2 {
    std_msgs::Header header;
    uint32_t height, width, point_step, row_step;
    std::vector<PointField> fields;
    std::vector<uint8_t> data;
    typedef std::shared_ptr<PointCloud2 const> ConstPtr;
};
typedef std::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
}
