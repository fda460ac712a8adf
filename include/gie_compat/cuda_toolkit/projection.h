// Stand-in for include/cuda_toolkit/projection.h.  `Projection` keeps the member names (L2G, G2L, origin); the SE3
// members expose the row-major 3x4 [R|t] as `data[12]` like cudaMat::SE3<float> (se3.cuh:196-199).
#pragma once
#include "cuda_toolkit/cuda_macro.h"

namespace cudaMat {
template <typename T> struct SE3 { T data[12]; };
}
struct Projection {
    cudaMat::SE3<float> L2G;
    cudaMat::SE3<float> G2L;
    float3 origin;
};

// ROS-free core of trans2proj (projection.h:15-33): body->world quaternion (w,x,y,z) and translation.
inline Projection make_projection(float qw, float qx, float qy, float qz, float tx, float ty, float tz)
{
    Projection p;
    const float q[4] = { qw, qx, qy, qz }, t[3] = { tx, ty, tz };
    GIE_CHECK(gie_make_projection(q, t, p.L2G.data, p.G2L.data));
    p.origin = make_float3(tx, ty, tz);
    return p;
}

#ifdef GIE_COMPAT_WITH_TF   // define when <tf/tf.h> is available (a ROS build)
#include <tf/tf.h>
inline Projection trans2proj(const tf::Transform &trans)
{
    tf::Quaternion r = trans.getRotation();
    tf::Vector3 o = trans.getOrigin();
    return make_projection((float)r.w(), (float)r.x(), (float)r.y(), (float)r.z(), (float)o.x(), (float)o.y(), (float)o.z());
}
#endif
