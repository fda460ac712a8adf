// Stand-in for include/map_structure/local_batch.h: class LocMap with the constructor, methods and public data members
// VOLMAPNODE uses (src/volumetric_mapper.cpp:73-78,154-155,208-209,375-389; include/volumetric_mapper.h:185-186).
// The dense device arrays are owned by the engine (gie_locmap); raw device views are available through dev_ptr().
#pragma once
#include <cmath>
#include "cuda_toolkit/cuda_macro.h"

#define VOXTYPE_OCCUPIED GIE_VOX_OCCUPIED
#define VOXTYPE_FREE GIE_VOX_FREE
#define VOXTYPE_UNKNOWN GIE_VOX_UNKNOWN
#define VOXTYPE_FNT GIE_VOX_FNT

typedef gie_seendist SeenDist;   // {float d; bool s; bool o;} — CostMap payload element (msg/CostMap.msg)

struct gie_hashmap;
class LocMap {
public:
    LocMap(const float voxel_size, const int3 local_size, const unsigned char occupancy_threshold, const float ogm_min_h,
           const float ogm_max_h, const int cutoff_grids_sq, const bool fast_mode)
        : _voxel_width(voxel_size), _local_size(local_size), _occu_thresh(occupancy_threshold), _update_max_h(ogm_max_h),
          _update_min_h(ogm_min_h), _cutoff_grids_sq(cutoff_grids_sq), _fast_mode(fast_mode)
    {
        const int X = local_size.x, Y = local_size.y, Z = local_size.z;
        _map_volume = X * Y * Z;
        _max_width = X + Y + Z;
        _max_loc_dist_sq = X * X + Y * Y + Z * Z;
        _bdr_num = 2 * (X * Y + Y * Z + X * Z);
        _half_shift = make_int3(X / 2, Y / 2, Z / 2);
        seendist_size = _map_volume * (int)sizeof(SeenDist);
        _pvt = _update_pvt = make_int3(0, 0, 0);
        _msg_origin = make_float3(0.f, 0.f, 0.f);
    }
    ~LocMap() { if (_h) delete_gpu_map(); }
    LocMap(const LocMap &) = delete;
    LocMap &operator=(const LocMap &) = delete;

    void create_gpu_map()
    {
        GIE_CHECK(gie_locmap_create(&_h, _voxel_width, _local_size.x, _local_size.y, _local_size.z, _occu_thresh,
                                    _update_min_h, _update_max_h, _cutoff_grids_sq, _fast_mode ? 1 : 0));
        glb_type_H = new char[_map_volume];
        edt_H = new float[_map_volume];
        seendist_out = new SeenDist[_map_volume];
    }
    void delete_gpu_map()
    {
        gie_locmap_destroy(_h);
        _h = nullptr;
        delete[] glb_type_H; delete[] edt_H; delete[] seendist_out;
        glb_type_H = nullptr; edt_H = nullptr; seendist_out = nullptr;
    }

    int3 calculate_pivot_origin(float3 map_center)
    {
        _center = map_center;
        push_pivots();
        return _pvt;
    }
    void calculate_update_pivot(float3 map_center)
    {
        // both pivots derive from the same centre in the reference's only call site (volumetric_mapper.cpp:154-155)
        if (map_center.x != _center.x || map_center.y != _center.y || map_center.z != _center.z) {
            _center = map_center;
            push_pivots();
        }
    }

    void copy_ogm_2_host() { GIE_CHECK(gie_locmap_copy_ogm_to_host(_h, (signed char *)glb_type_H)); }
    void copy_edt_2_host() { GIE_CHECK(gie_locmap_copy_edt_to_host(_h, edt_H)); }
    // packs SeenDist on the device and copies once (the reference does two D2H copies and a host loop, :382-391)
    void convertCostMap() { GIE_CHECK(gie_locmap_convert_costmap(_h, seendist_out)); }

    int id(int x, int y, int z, int /*phase*/ = 0) const { return z * _local_size.x * _local_size.y + y * _local_size.x + x; }
    int coord2idx_local(const int3 &c) const { return id(c.x, c.y, c.z); }
    float3 coord2pos(const int3 &c) const { return make_float3(c.x * _voxel_width, c.y * _voxel_width, c.z * _voxel_width); }
    int3 pos2coord(const float3 &p) const
    {
        return make_int3((int)floorf(p.x / _voxel_width + 0.5f), (int)floorf(p.y / _voxel_width + 0.5f),
                         (int)floorf(p.z / _voxel_width + 0.5f));
    }
    int3 loc2glb(const int3 &c) const { return make_int3(c.x + _pvt.x, c.y + _pvt.y, c.z + _pvt.z); }
    int3 glb2loc(const int3 &c) const { return make_int3(c.x - _pvt.x, c.y - _pvt.y, c.z - _pvt.z); }
    bool is_inside_local_volume(const int3 &c) const
    {
        return c.x >= 0 && c.x < _local_size.x && c.y >= 0 && c.y < _local_size.y && c.z >= 0 && c.z < _local_size.z;
    }

    // engine access
    gie_locmap *handle() const { return _h; }
    void *dev_ptr(int which) const { void *p = nullptr; GIE_CHECK(gie_locmap_device_ptr(_h, which, &p, nullptr)); return p; }
    void set_stream(void *cuda_stream) { GIE_CHECK(gie_set_stream(_h, cuda_stream)); }

    float _voxel_width;
    int3 _local_size;
    unsigned char _occu_thresh;
    float _update_max_h, _update_min_h;
    int3 _pvt;
    int _max_width, _max_loc_dist_sq, _bdr_num, _map_volume;
    char *glb_type_H = nullptr;
    float *edt_H = nullptr;
    SeenDist *seendist_out = nullptr;
    int3 _half_shift;
    int seendist_size;
    float3 _msg_origin;
    int _cutoff_grids_sq;
    bool _fast_mode;
    int3 _update_pvt;
    gie_hashmap *_hash = nullptr;   // set by GlbHashMap::setLocMap

private:
    void push_pivots()
    {
        const float c[3] = { _center.x, _center.y, _center.z };
        GIE_CHECK(gie_locmap_calculate_pivots(_h, c));
        int p[6]; float o[3];
        GIE_CHECK(gie_locmap_get_pivots(_h, p, o));
        _pvt = make_int3(p[0], p[1], p[2]);
        _update_pvt = make_int3(p[3], p[4], p[5]);
        _msg_origin = make_float3(o[0], o[1], o[2]);
    }
    gie_locmap *_h = nullptr;
    float3 _center = make_float3(NAN, NAN, NAN);
};
