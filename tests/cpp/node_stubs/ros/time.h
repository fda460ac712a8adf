#pragma once
namespace ros {
struct Time { double sec = 0; static Time now() { return Time(); } double toSec() const { return sec; } };
struct Duration { double sec; Duration(double s = 0) : sec(s) {} };
}
