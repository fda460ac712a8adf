#pragma once
#include <string>
namespace std_srvs { struct SetBool { struct { bool data; } request; struct { bool success; std::string message; } response; }; }
