// Stand-in for include/par_wave/glb_hash_map.h: struct GlbHashMap with the methods and public data VOLMAPNODE uses
// (src/volumetric_mapper.cpp:82-83,157-198; README.md:163-170 for the host mirror of the global map).
#pragma once
#include <unordered_map>
#include <vector>
#include <thrust/detail/raw_pointer_cast.h>   // the node wraps VB_keys_loc_D.data() in thrust::raw_pointer_cast
#include "par_wave/voxmap_utils.cuh"
#include "map_structure/pre_map.h"

// `thrust::raw_pointer_cast(_hash_map->VB_keys_loc_D.data())` in the node keeps compiling: data() is a raw device pointer.
// The engine records touched blocks itself, so this is a token allocation, not the reference's 12 B/voxel key array.
struct GieKeyArrayStub {
    int3 *ptr = nullptr;
    int3 *data() const { return ptr; }
};

struct GlbHashMap {
public:
    GlbHashMap(int /*bdr_size*/, int3 /*loc_dim*/, int bucket_max, int block_max) : _bucket_max(bucket_max), _block_max(block_max) {}
    ~GlbHashMap() { if (_h) gie_hashmap_destroy(_h); }
    GlbHashMap(const GlbHashMap &) = delete;
    GlbHashMap &operator=(const GlbHashMap &) = delete;

    void setLocMap(LocMap *lMap)
    {
        _lMap = lMap;
        GIE_CHECK(gie_hashmap_create(&_h, lMap->handle(), _bucket_max, _block_max));
        lMap->_hash = _h;
        void *p = nullptr;
        GIE_CHECK(gie_locmap_device_ptr(lMap->handle(), GIE_ARR_RAY_COUNT, &p, nullptr));
        VB_keys_loc_D.ptr = (int3 *)p;   // any valid device address; never dereferenced by the engine
    }
    void allocHashTB() {}   // folded into updateHashOGM (blocks are allocated by the merge kernel)

    void updateHashOGM(bool input_pynt, const int map_ct, bool stream_glb_ogm, Ext_Obs_Wrapper *ext_obsv)
    {
        int n = ext_obsv ? ext_obsv->ext_obs_num : 0;
        GIE_CHECK(gie_hashmap_update_ogm(_h, input_pynt, map_ct, stream_glb_ogm, n,
                                         n ? &ext_obsv->rt_obsbbx_ll[0].x : nullptr, n ? &ext_obsv->rt_obsbbx_ur[0].x : nullptr,
                                         n ? ext_obsv->obs_activated.data() : nullptr));
    }
    void mergeNewObsv(const int map_ct, const bool display_glb_edt) { GIE_CHECK(gie_hashmap_merge_new_obsv(_h, map_ct, display_glb_edt)); }

    // GPU -> CPU streaming of the blocks that changed since the last call: one gather kernel + one copy
    void streamPipeline()
    {
        int n = 0;
        GIE_CHECK(gie_hashmap_num_changed(_h, &n));
        if (n == 0) return;
        _stage_keys.resize((size_t)n);
        _stage_vals.resize((size_t)n);
        GIE_CHECK(gie_hashmap_stream_changed(_h, (int32_t *)_stage_keys.data(), (gie_glbvoxel *)_stage_vals.data(), n, &n));
        for (int i = 0; i < n; i++) {
            auto it = hash_table_H_std.find(_stage_keys[i]);
            int idx;
            if (it == hash_table_H_std.end()) {
                idx = VB_cnt_H++;
                hash_table_H_std.emplace(_stage_keys[i], idx);
                if ((size_t)idx >= VB_values_H.size()) { VB_values_H.resize((size_t)idx * 2 + 64); VB_keys_H.resize((size_t)idx * 2 + 64, EMPTY_KEY); }
                VB_keys_H[idx] = _stage_keys[i];
            } else idx = it->second;
            VB_values_H[idx] = _stage_vals[i];
        }
    }
    void sync() { GIE_CHECK(gie_sync(_h)); }   // surfaces sticky device errors (out of blocks, queue overflow)
    gie_hashmap *handle() const { return _h; }

    std::vector<VoxelBlock> VB_values_H;
    std::vector<int3> VB_keys_H;
    std::unordered_map<int3, int, BlockHasher, CrdEqualTo> hash_table_H_std;
    int VB_cnt_H = 0;
    GieKeyArrayStub VB_keys_loc_D;
    LocMap *_lMap = nullptr;

private:
    gie_hashmap *_h = nullptr;
    int _bucket_max, _block_max;
    std::vector<int3> _stage_keys;
    std::vector<VoxelBlock> _stage_vals;
};
