// msg/CostMap.msg of the reference as a plain struct
#pragma once
#include <cstdint>
#include <vector>
namespace GIE {
struct CostMap {
    enum { TYPE_EDT = 0 };
    float x_origin, y_origin, z_origin, width;
    int x_size, y_size, z_size;
    int type;
    std::vector<uint8_t> payload8;
};
}
