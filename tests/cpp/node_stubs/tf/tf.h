#pragma once
#include <string>
#include <ros/time.h>
#include <geometry_msgs/TransformStamped.h>
namespace tf {
struct Quaternion {
    double q[4];   // x y z w
    Quaternion() : q{0, 0, 0, 1} {}
    Quaternion(double x, double y, double z, double w) : q{x, y, z, w} {}
    double x() const { return q[0]; } double y() const { return q[1]; } double z() const { return q[2]; } double w() const { return q[3]; }
};
struct Vector3 {
    double m_floats[4];
    Vector3() : m_floats{0, 0, 0, 0} {}
    Vector3(double x, double y, double z) : m_floats{x, y, z, 0} {}
    double x() const { return m_floats[0]; } double y() const { return m_floats[1]; } double z() const { return m_floats[2]; }
};
struct Transform {
    Quaternion rot; Vector3 org;
    Transform() {}
    Transform(const Quaternion &r, const Vector3 &o) : rot(r), org(o) {}
    Quaternion getRotation() const { return rot; }
    Vector3 getOrigin() const { return org; }
    void setRotation(const Quaternion &r) { rot = r; }
    void setOrigin(const Vector3 &o) { org = o; }
};
struct StampedTransform : Transform {
    StampedTransform(const Transform &t, const ros::Time &, const std::string &, const std::string &) : Transform(t) {}
};
inline void quaternionMsgToTF(const geometry_msgs::Quaternion &m, Quaternion &q) { q = Quaternion(m.x, m.y, m.z, m.w); }
}
