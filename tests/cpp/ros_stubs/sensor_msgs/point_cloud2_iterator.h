// Stand-in for sensor_msgs/point_cloud2_iterator.h: PointCloud2ConstIterator<T> over the field named in the constructor;
// it[k] is the k-th T after that field inside the current point, as in ROS.
#pragma once
#include <cstring>
#include <stdexcept>
#include <sensor_msgs/PointCloud2.h>
namespace sensor_msgs {
template <typename T>
class PointCloud2ConstIterator {
public:
    PointCloud2ConstIterator(const PointCloud2 &msg, const std::string &field) : step_(msg.point_step)
    {
        size_t off = 0;
        bool found = false;
        for (const auto &f : msg.fields) if (f.name == field) { off = f.offset; found = true; }
        if (!found) throw std::runtime_error("field " + field + " does not exist");
        cur_ = msg.data.data() + off;
        end_ = cur_ + (size_t)msg.width * msg.height * msg.point_step;
    }
    const T &operator[](size_t i) const { return *(reinterpret_cast<const T *>(cur_) + i); }
    const T &operator*() const { return *reinterpret_cast<const T *>(cur_); }
    PointCloud2ConstIterator &operator++() { cur_ += step_; return *this; }
    bool operator!=(const PointCloud2ConstIterator &o) const { return cur_ != o.cur_; }
    PointCloud2ConstIterator end() const { PointCloud2ConstIterator e(*this); e.cur_ = end_; return e; }
private:
    const unsigned char *cur_ = nullptr, *end_ = nullptr;
    size_t step_;
};
}
