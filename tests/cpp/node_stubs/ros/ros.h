// Compile-only stand-in for <ros/ros.h> (+ the boost names it drags in), just wide enough for the reference's
// include/volumetric_mapper.h, include/parameters.h and src/volumetric_mapper.cpp.  Test infrastructure only.
#pragma once
#include <cmath>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#include <ros/time.h>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
using std::bind;
}
using namespace std::placeholders;   // _1, _2 of boost::bind
namespace ros {
struct TimerEvent {};
struct Timer { void stop() {} void start() {} };
struct Publisher { template <class M> void publish(const M &) const {} };
struct Subscriber {};
class NodeHandle {
public:
    template <class T, class D> bool param(const std::string &, T &v, const D &d) const { v = (T)d; return false; }
    template <class T> bool getParam(const std::string &, T &) const { return false; }
    template <class M> Publisher advertise(const std::string &, int) { return Publisher(); }
    template <class M, class C> Subscriber subscribe(const std::string &, int, void (C::*)(const std::shared_ptr<M const> &), C *) { return Subscriber(); }
    template <class C> Timer createTimer(Duration, void (C::*)(const TimerEvent &), C *) { return Timer(); }
};
inline void init(int &, char **, const std::string &) {}
inline void spin() {}
namespace service { template <class S> bool call(const std::string &, S &) { return true; } }
}
