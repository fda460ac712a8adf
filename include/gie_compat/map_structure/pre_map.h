// Stand-in for include/map_structure/pre_map.h (+ src/kernel/pre_map/pre_map.cu): external-obstacle AABBs / virtual fence.
// Host vectors only; GlbHashMap::updateHashOGM hands the activated boxes to the engine each frame.
#pragma once
#include <vector>
#include "cuda_toolkit/projection.h"
#include "par_wave/voxmap_utils.cuh"
#include "warmup.h"

class Ext_Obs_Wrapper {
public:
    std::vector<float3> rt_obsbbx_ll, rt_obsbbx_ur;
    std::vector<unsigned char> obs_activated;   // [0] = outer fence: voxels OUTSIDE box 0 are obstacles when set
    int ext_obs_num;

    explicit Ext_Obs_Wrapper(int obs_num) { change_obs_num(obs_num); }
    void change_obs_num(int obs_num)
    {
        ext_obs_num = obs_num;
        rt_obsbbx_ll.resize(obs_num, make_float3(0, 0, 0));
        rt_obsbbx_ur.resize(obs_num, make_float3(0, 0, 0));
        obs_activated.resize(obs_num, 0);
    }
    void assign_obs_premap(std::vector<float3> &ll, std::vector<float3> &ur) { rt_obsbbx_ll = ll; rt_obsbbx_ur = ur; }
    void append_new_elem(float3 &ll, float3 &ur) { rt_obsbbx_ll.push_back(ll); rt_obsbbx_ur.push_back(ur); }
    bool CheckAABBIntersection(float3 &a_ll, float3 &a_ur, float3 &b_ll, float3 &b_ur)
    {
        return a_ll.x <= b_ur.x && a_ur.x >= b_ll.x && a_ll.y <= b_ur.y && a_ur.y >= b_ll.y && a_ll.z <= b_ur.z && a_ur.z >= b_ll.z;
    }
    void bbx_H2D() { if (ext_obs_num != (int)rt_obsbbx_ll.size()) { ext_obs_num = (int)rt_obsbbx_ll.size(); obs_activated.resize(ext_obs_num, 0); } }
    // pre_map.cu:80-101: box 0 never activated here; box i>0 activated when it intersects the local volume
    void activate_AABB(float3 &loc_map_ll, float3 &loc_map_ur)
    {
        bbx_H2D();
        if (ext_obs_num > 0) obs_activated[0] = 0;
        for (int i = 1; i < ext_obs_num; i++)
            obs_activated[i] = CheckAABBIntersection(loc_map_ll, loc_map_ur, rt_obsbbx_ll[i], rt_obsbbx_ur[i]) ? 1 : 0;
    }
};
