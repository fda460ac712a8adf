#pragma once
#include <cstdint>
#include <string>
namespace std_msgs { struct Header { uint32_t seq = 0; double stamp = 0; std::string frame_id; }; }
