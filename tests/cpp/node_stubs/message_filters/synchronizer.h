#pragma once
#include <message_filters/subscriber.h>
namespace message_filters {
template <class Policy> struct Synchronizer {
    template <class A, class B> Synchronizer(const Policy &, A &, B &) {}
    void setMaxIntervalDuration(ros::Duration) {}
    template <class F> void registerCallback(const F &) {}
};
}
