// Sensor parameter PODs of the four OGM front ends (reference: include/cuda_toolkit/occupancy/*/{pntcld,scan,multiscan,
// camera}_param.h).  Member names and constructor argument order follow the reference so call sites compile unchanged.
#pragma once
#include <vector_types.h>
struct PntcldParam {
    int cld_sz = 0, valid_pnt_count = 0;
    PntcldParam() = default;
    explicit PntcldParam(int n) : cld_sz(n) {}
};
struct ScanParam {
    float max_r = 0.f, theta_inc = 0.f, theta_min = 0.f;
    int scan_num = 0;
    ScanParam() = default;
    ScanParam(int n, float max_range, float d_theta, float theta0) : max_r(max_range), theta_inc(d_theta), theta_min(theta0), scan_num(n) {}
};
struct MulScanParam {
    float max_r = 0.f, theta_inc = 0.f, theta_min = 0.f, phi_inc = 0.f, phi_min = 0.f;
    int scan_num = 0, ring_num = 0;
    MulScanParam() = default;
    MulScanParam(int n, int rings, float max_range, float d_theta, float theta0, float d_phi, float phi0)
        : max_r(max_range), theta_inc(d_theta), theta_min(theta0), phi_inc(d_phi), phi_min(phi0), scan_num(n), ring_num(rings) {}
};
struct CamParam {
    int rows = 0, cols = 0;
    float cx = 0.f, cy = 0.f, fx = 0.f, fy = 0.f;
    bool valid_NaN = false;
    CamParam() = default;
    CamParam(int r, int c, float cx_, float cy_, float fx_, float fy_, bool nan_ok) : rows(r), cols(c), cx(cx_), cy(cy_), fx(fx_), fy(fy_), valid_NaN(nan_ok) {}
};
typedef float SCAN_DEPTH_TPYE;        // spelling as in the reference
typedef float REALSENSE_DEPTH_TPYE;
typedef float LASER_RANGE_TPYE;
typedef float3 PNT_TYPE;
