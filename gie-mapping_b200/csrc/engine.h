// engine.h — host-side handle structs and the launch functions each .cu file provides.
#pragma once
#include "common.cuh"
#include "../../include/gie_b200.h"
#include <cstdlib>
#include <string>

struct XsLaunch { int rpi = 32, threads = 0, NB = 0, CAP = 0, BW = 0; size_t smem = 0; };   // launch shape of a banded sweep (edt.cu)

struct gie_locmap {
    LocDev d{};                 // device view, passed by value to kernels (as the reference passes LocMap)
    cudaStream_t stream = nullptr;
    int device = 0;
    int num_sms = 148;
    float3 msg_origin{};
    // batch-EDT scratch
    unsigned long long *ytab = nullptr;   // [Z][ceil(Y/32)][X] (mask word, lo_prev, hi_next)
    int32_t *g2 = nullptr;                // [Z][Y][X] plane dist_sq after the x sweep
    int32_t *cxy = nullptr;               // [Z][Y][X] cocx | cocy << 16
    int *col_list = nullptr;              // [Z][X] columns of each slice that hold an obstacle (ascending)
    int *edt_meta = nullptr;              // n_cols[Z], slice_list[Z], n_slices
    int *slice_has = nullptr;             // [Z] 1 = the OGM merge set a y-pass bit in this slice (valid while ytab is in step with the merge)
    unsigned long long *stack_scratch = nullptr;
    size_t stack_scratch_entries = 0;
    int edt_ctas = 0;                     // persistent grid of the z sweep
    int xs_ctas = 0;                      // persistent grid of the x sweep
    XsLaunch xs;
    bool edt_compact = false;             // ytab / col_list hold only the obstacle-bearing slices, in slice_list order (slab map fed by a peer)
    bool slab_only = false;               // a slab map of a sharded volume: batch-EDT arrays for rows [ys0, ys0 + ysn) only
    bool edt_inputs_aliased = false;      // ytab / col_list / edt_meta belong to another map on this device
    void *ipc_opened[16]{};               // peer slab arrays mapped through CUDA IPC (closed on destroy)
    int n_ipc_opened = 0;
    XsLaunch zs;                          // banded z sweep (dense regime)
    bool zs_banded = false;
    int zs_ctas = 0;
    int *work_counters = nullptr;         // device: [4]
    // ray-cast scratch: per-ray checkpoints, step counts and stop indices (ogm.cu)
    void *ray_scratch = nullptr;
    size_t ray_scratch_bytes = 0;
    // staging for *_host entry points
    float *stage_dev = nullptr;
    size_t stage_bytes = 0;
    // profiling
    bool profile = false;
    cudaEvent_t ev[GIE_ST_COUNT][2]{};
    bool ev_valid[GIE_ST_COUNT]{};
    long long launches = 0;
    bool glb_type_foreign = false;        // glb_type was overwritten from outside (test hook): the next merge clears all of it
    long long ytab_serial = -1;           // OGM merge whose block list describes the y-pass bit words that are set (-1: unknown, clear all)
    gie_hashmap *hm = nullptr;
};

struct gie_hashmap {
    HashDev d{};
    gie_locmap *lm = nullptr;
    size_t hash_cap = 0;
    int halo_blocks = 1;
    size_t tab_entries = 0;
    // wavefront queues
    unsigned long long *qA[3]{};  // outside queues: packed global coords
    unsigned long long *qB[3]{};
    unsigned long long *qC[3]{};  // inside queue: (pusher coc id << 32) | packed local coords
    unsigned long long *cseed_key = nullptr;  // deferred C-seed pair updates (same slots as qC[0])
    int queue_cap = 0;
    int *counters = nullptr;      // device ints, see wave.cu
    unsigned int *barrier = nullptr;
    size_t barrier_words = 0;
    // decision scratch for wave A (parallel to the current queue)
    int32_t *decA_dist = nullptr;
    unsigned long long *decA_coc = nullptr;
    unsigned long long *decA_pair = nullptr;
    int32_t *decA_flags = nullptr;
    uint32_t *snap_id = nullptr;  // per-queue-slot snapshot for waves B/C
    int wave_ctas = 0;
    // Table indices of the touched or allocated blocks that intersect the volume, per OGM merge, with the pivots they were
    // listed under.  Two buffers: blists[bl_cur] belongs to the latest merge, the other to the one before — whatever the
    // previous frame wrote into glb_type (and into the y-pass bit words) lies inside THOSE blocks, so that is what gets cleared.
    struct BlockList { int *list = nullptr; int *count = nullptr; int3 pvt{}, tab_org{}; bool valid = false; };
    BlockList blists[2];
    int bl_cur = 0;
    long long merge_serial = 0;   // OGM merges done so far
    int *blk_list = nullptr;      // table indices of the allocated blocks that intersect the local volume (per merge)
    int4 *blk_org = nullptr;      // per listed block: local coordinates of its first voxel, pool index
    int *blk_count = nullptr;
    int merge_epoch = 0;          // merges done so far; the seed mark of m.wave_layer (memset to 0 at creation)
    int wave_cluster = 1;         // CTAs per thread-block cluster of the wave kernel
    unsigned long long *wave_trace = nullptr;   // diagnostics, allocated when GIE_WAVE_TRACE is set
    // external-obstacle boxes of the current frame: [n][7] = ll.xyz, ur.xyz, activated
    float *obs_dev = nullptr;
    int obs_cap = 0;
    // streaming scratch: compacted list of changed blocks
    int *changed_list = nullptr;
    int *changed_count = nullptr;
    int *status_host = nullptr;   // pinned
    long long *stats_host = nullptr;  // pinned [16]: wave statistics [0..7], sticky device status [8]
};

void gie_set_error(const std::string &msg);
#define GIE_CUDA_CHECK(expr)                                                                           \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            gie_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                         \
            return GIE_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

// Per-frame kernels are launched with programmatic dependent launch: the next kernel's CTAs are placed while the previous
// kernel drains, and wait (griddepcontrol.wait, the first statement of every such kernel — before any early return, so that the
// dependency chain stays transitive) until it has completed and its memory operations are visible.  20 kernel boundaries per
// frame otherwise cost a drain + launch each.  GIE_NO_PDL=1 falls back to plain stream order (the device-side instructions
// are no-ops then).
inline bool gie_pdl_enabled()
{
    static const bool on = getenv("GIE_NO_PDL") == nullptr;
    return on;
}
template <typename... KArgs, typename... Args>
inline void gie_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = gie_pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in the caller's cudaGetLastError check
}

struct StageTimer {
    gie_locmap *lm; int st;
    StageTimer(gie_locmap *l, int s) : lm(l), st(s) { if (lm->profile) { cudaEventRecord(lm->ev[st][0], lm->stream); } }
    ~StageTimer() { if (lm->profile) { cudaEventRecord(lm->ev[st][1], lm->stream); lm->ev_valid[st] = true; } }
};

// ogm.cu
int gie_launch_ogm_pointcloud(gie_locmap *lm, gie_hashmap *hm, const float *pts_dev, int n, int fmp, int r2);
int gie_launch_ogm_scan2d(gie_locmap *lm, gie_hashmap *hm, const float *scan, int scan_num, float tinc, float tmin, int fmp, int r2);
int gie_launch_ogm_vlp16(gie_locmap *lm, gie_hashmap *hm, const float *ranges, int scan_num, int ring_num, float tinc,
                         float tmin, float pinc, float pmin, int fmp, int r2);
int gie_launch_ogm_depth(gie_locmap *lm, gie_hashmap *hm, const float *img, int rows, int cols, float cx, float cy,
                         float fx, float fy, int valid_nan, int fmp, int r2);
int gie_launch_vlp16_bin(gie_locmap *lm, const unsigned char *raw_dev, int n, int step, int off_x, int off_y, int off_ring,
                         int scan_num, int ring_num, float theta_inc, unsigned long long *img_dev, float *ranges_dev);
int gie_launch_pc_repack(gie_locmap *lm, const unsigned char *raw_dev, int n, int step, int off_x, float *pts_dev);
// hashmap.cu
int gie_hash_begin_frame(gie_hashmap *hm);                       // sets the table origin, clears touched flags
int gie_launch_update_ogm(gie_hashmap *hm, int input_pntcld, int map_ct, int stream_glb_ogm, int n_obs);
int gie_launch_export(gie_hashmap *hm, int nblocks, gie_glbvoxel *out_dev);
int gie_launch_list_changed(gie_hashmap *hm, int nblocks, int clear);   // -> changed_list / changed_count
int gie_launch_gather_changed(gie_hashmap *hm, int first, int n, int32_t *keys_dev, gie_glbvoxel *out_dev);
// edt.cu
int gie_edt_prepare(gie_locmap *lm);
int gie_launch_batch_edt(gie_locmap *lm);
int gie_launch_edt_xy(gie_locmap *lm);
int gie_launch_edt_z(gie_locmap *lm, int max_width_override);
int gie_launch_edt_pack(gie_locmap *lm, unsigned long long *ytab_compact, int *col_compact);
int gie_launch_edt_slab(gie_locmap *lm, int max_width);
// wave.cu
int gie_wave_prepare(gie_hashmap *hm);
int gie_launch_merge(gie_hashmap *hm, int map_ct, int display_glb_edt);
