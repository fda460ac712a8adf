/*
 * gie_b200.h — C ABI of the B200-native OGM + incremental-EDT engine.
 *
 * This is the drop-in boundary for ONE hot path of JINXER000/GIE-mapping: what
 * VOLMAPNODE::publishMap (src/volumetric_mapper.cpp:138-224) calls per frame.
 * Every entry point names the reference interface it replaces (file:line in the
 * reference repo).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns GIE_OK (0) or a negative gie_status; the reference
 *     exit(1)s / asserts instead (include/cuda_toolkit/cuda_macro.h:21-31).
 *     gie_last_error() returns a thread-local message for the last failure.
 *   - "_dev" pointers are device pointers on the engine's device, "_host" are
 *     host pointers (pinned or pageable).  Output buffers are caller-owned.
 *   - all work is enqueued on the stream given to gie_set_stream (default: the
 *     legacy default stream, as in the reference); calls that copy to the host
 *     synchronise that stream before returning.
 *   - not re-entrant per handle (the reference is single-threaded per node).
 */
#ifndef GIE_B200_H
#define GIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gie_status {
    GIE_OK = 0,
    GIE_ERR_INVALID_ARG = -1,
    GIE_ERR_CUDA = -2,
    GIE_ERR_OUT_OF_BLOCKS = -3,   /* reference: throw "out of block memory" (include/vox_hash/blockalloc.h:56-58) */
    GIE_ERR_QUEUE_OVERFLOW = -4,  /* reference: assert in parWave (src/kernel/par_wave/wave_helper.h:26-30,82-88) */
    GIE_ERR_SIZE_UNSUPPORTED = -5 /* reference: "Local map size too big!!!" (include/map_structure/local_batch.h:54-58) */
} gie_status;

/* voxel types: include/map_structure/local_batch.h:7-10 */
#define GIE_VOX_UNKNOWN 0
#define GIE_VOX_FREE 1
#define GIE_VOX_OCCUPIED 2
#define GIE_VOX_FNT 3
#define GIE_EMPTY_VALUE 999999 /* include/par_wave/voxmap_utils.cuh:8 */

typedef struct gie_locmap gie_locmap;   /* replaces class LocMap   (include/map_structure/local_batch.h:32-569) */
typedef struct gie_hashmap gie_hashmap; /* replaces struct GlbHashMap (include/par_wave/glb_hash_map.h:11-65)   */

/* SeenDist, include/map_structure/local_batch.h:19-24 (CostMap payload, msg/CostMap.msg) */
typedef struct gie_seendist {
    float d;
    unsigned char s;
    unsigned char o;
} gie_seendist;

/* GlbVoxel in the reference's memory layout (include/par_wave/voxmap_utils.cuh:29-44), 40 bytes.
 * Used only by gie_hashmap_export_blocks; inside the engine blocks are field-major. */
typedef struct gie_glbvoxel {
    unsigned char occ_val;
    signed char vox_type;
    int32_t update_ct;
    int32_t coc_glb[3];
    int32_t dist_sq;
    int32_t wave_layer;
    uint64_t dist_id_pair; /* Dist_id union (local_batch.h:26-30): low word sq_dist[0] = dist_sq, high word
                              parent_loc_id[1] = wave-range coc id (11/11/10 bit) */
} gie_glbvoxel;

const char *gie_last_error(void);
const char *gie_version(void);

/* ---- LocMap ------------------------------------------------------------------------------------- */
/* LocMap::LocMap + create_gpu_map (local_batch.h:35-89).  Requires X,Y <= 1024, Z <= 512 (coc codec 11/11/10 bit). */
int gie_locmap_create(gie_locmap **out, float voxel_size, int size_x, int size_y, int size_z,
                      unsigned char occupancy_threshold, float ogm_min_h, float ogm_max_h, int cutoff_grids_sq,
                      int fast_mode);
/* LocMap::delete_gpu_map (local_batch.h:91-110) */
int gie_locmap_destroy(gie_locmap *lm);
/* cudaStream_t as void*; NULL = legacy default stream */
int gie_set_stream(gie_locmap *lm, void *cuda_stream);
/* trans2proj (include/cuda_toolkit/projection.h:15-33) + LocMap::calculate_pivot_origin / calculate_update_pivot
 * (local_batch.h:128-166), as called at volumetric_mapper.cpp:144-155.  q = (w,x,y,z) body->world, t = translation. */
int gie_locmap_set_pose(gie_locmap *lm, const float q_wxyz[4], const float t_xyz[3]);
/* The same three steps as separate calls, for callers that keep the reference's call order (volumetric_mapper.cpp:144-155,
 * where proj.origin.z may be overridden by ugv_height before the pivots are computed):
 *   gie_make_projection        = trans2proj: cudaMat::SE3(qw,qx,qy,qz,tx,ty,tz) and its inverse (se3.cuh:47-75,89-105),
 *                                row-major 3x4 [R|t]
 *   gie_locmap_set_projection  = the `Projection proj` the reference passes by value to every localOGMKernels
 *   gie_locmap_calculate_pivots= LocMap::calculate_pivot_origin + calculate_update_pivot (local_batch.h:128-166) */
int gie_make_projection(const float q_wxyz[4], const float t_xyz[3], float L2G[12], float G2L[12]);
int gie_locmap_set_projection(gie_locmap *lm, const float L2G[12], const float G2L[12], const float origin[3]);
int gie_locmap_calculate_pivots(gie_locmap *lm, const float map_center[3]);
/* out6 = {_pvt.xyz, _update_pvt.xyz}; origin3 = _msg_origin */
int gie_locmap_get_pivots(const gie_locmap *lm, int out6[6], float origin3[3]);
/* LocMap::copy_ogm_2_host / copy_edt_2_host / convertCostMap (local_batch.h:370-391) */
int gie_locmap_copy_ogm_to_host(gie_locmap *lm, signed char *glb_type_host);
int gie_locmap_copy_edt_to_host(gie_locmap *lm, float *edt_host);
int gie_locmap_convert_costmap(gie_locmap *lm, gie_seendist *seendist_host);
/* raw device views for GPU consumers (LocMap::_glb_type, _edt_D, _aux, _coc_idx_aux, _dist_id_pair; local_batch.h:540-562) */
enum gie_array { GIE_ARR_RAY_COUNT = 0, GIE_ARR_INST_TYPE = 1, GIE_ARR_GLB_TYPE = 2, GIE_ARR_EDT = 3, GIE_ARR_AUX = 4,
                 GIE_ARR_COC_AUX = 5, GIE_ARR_PAIR = 6,
                 /* batch-EDT intermediates, for the multi-GPU re-partition (gie_edt_xy_sweeps / gie_edt_z_sweep) */
                 GIE_ARR_EDT_G2 = 7,    /* int32 [Z][Y][X]: in-slice squared distance after the y and x sweeps */
                 GIE_ARR_EDT_CXY = 8,   /* int32 [Z][Y][X]: closest obstacle of the slice, x | y << 16 */
                 GIE_ARR_EDT_NCOLS = 9  /* int32 [Z]: obstacle-bearing columns per slice (0 = the slice is skipped) */ };
int gie_locmap_device_ptr(gie_locmap *lm, int which, void **dev_ptr, size_t *bytes);
int gie_locmap_download(gie_locmap *lm, int which, void *host_out);
/* test hook: overwrite _glb_type (N bytes) so the batch EDT can be driven directly */
int gie_locmap_upload_glb_type(gie_locmap *lm, const signed char *glb_type_host);

/* ---- GlbHashMap --------------------------------------------------------------------------------- */
/* GlbHashMap::GlbHashMap + setLocMap (src/kernel/par_wave/glb_hash_map.cu:9-56) */
int gie_hashmap_create(gie_hashmap **out, gie_locmap *lm, int bucket_max, int block_max);
int gie_hashmap_destroy(gie_hashmap *hm);

/* ---- OGM: *_FAST::localOGMKernels / *MapMaker::updateLocalOGM ------------------------------------
 * The reference passes `int3* VB_keys_loc_D` (12 B/voxel key array owned by GlbHashMap); here the hash map
 * handle is passed instead and the engine records touched blocks in a bitmap. */
/* PNTCLD_RAYCAST::localOGMKernels (src/kernel/point_cloud/pntcld_raycast.cu:105-117); pts = float3[n], sensor frame */
int gie_ogm_pointcloud_dev(gie_locmap *lm, gie_hashmap *hm, const float *pts_dev, int n, int for_motion_planner,
                           int rbt_r2_grids);
/* PntcldMapMaker::updateLocalOGM (src/pntcld_map_maker.cpp:63-73): H2D copy + kernels */
int gie_ogm_pointcloud_host(gie_locmap *lm, gie_hashmap *hm, const float *pts_host, int n, int for_motion_planner,
                            int rbt_r2_grids);
/* HOKUYO_FAST::localOGMKernels (src/kernel/hokuyo/hokuyo_fast.cu:83-91); ScanParam: include/.../hokuyo/scan_param.h */
int gie_ogm_scan2d_dev(gie_locmap *lm, gie_hashmap *hm, const float *scan_dev, int scan_num, float theta_inc,
                       float theta_min, int for_motion_planner, int rbt_r2_grids);
int gie_ogm_scan2d_host(gie_locmap *lm, gie_hashmap *hm, const float *scan_host, int scan_num, float theta_inc,
                        float theta_min, int for_motion_planner, int rbt_r2_grids);
/* VLP_FAST::localOGMKernels (src/kernel/vlp16/vlp16_fast.cu:89-97); ranges = float[ring_num*scan_num] horizontal range */
int gie_ogm_vlp16_dev(gie_locmap *lm, gie_hashmap *hm, const float *ranges_dev, int scan_num, int ring_num,
                      float theta_inc, float theta_min, float phi_inc, float phi_min, int for_motion_planner,
                      int rbt_r2_grids);
int gie_ogm_vlp16_host(gie_locmap *lm, gie_hashmap *hm, const float *ranges_host, int scan_num, int ring_num,
                       float theta_inc, float theta_min, float phi_inc, float phi_min, int for_motion_planner,
                       int rbt_r2_grids);
/* REALSENSE_FAST::localOGMKernels (src/kernel/realsense/realsense_fast.cu:97-105); depth = float[rows*cols] metres */
int gie_ogm_depth_dev(gie_locmap *lm, gie_hashmap *hm, const float *depth_dev, int rows, int cols, float cx, float cy,
                      float fx, float fy, int valid_nan, int for_motion_planner, int rbt_r2_grids);
int gie_ogm_depth_host(gie_locmap *lm, gie_hashmap *hm, const float *depth_host, int rows, int cols, float cx,
                       float cy, float fx, float fy, int valid_nan, int for_motion_planner, int rbt_r2_grids);

/* Raw sensor_msgs/PointCloud2 front ends: the MapMakers' host-side conversion loops moved to the device (one upload of the
 * message bytes, no per-ring copies).  data_host = msg->data, point_step = msg->point_step, off_* = field offsets.
 *   Vlp16MapMaker::convertPyntCld + updateLocalOGM (src/vlp16_map_maker.cpp:52-147): float32 x,y and uint16 ring are binned
 *     into the [ring_num][scan_num] horizontal-range image, bin = (int)((atan2f(y,x) + pi) / |theta_inc|), the last point of a
 *     bin in message order wins, empty bins are INFINITY; then VLP_FAST::localOGMKernels.  (The reference calls glibc's atan2f
 *     on the host; here it is CUDA's — a point within an ulp of a bin edge can land in the neighbouring bin.)
 *   PntcldMapMaker::pntcld_process + updateLocalOGM (src/pntcld_map_maker.cpp:49-73): the three consecutive floats at off_x of
 *     the first max_points points (cld_sz; 0 = no cap) become the float3 cloud; then PNTCLD_RAYCAST::localOGMKernels.
 * gie_vlp16_last_ranges downloads the range image of the last vlp16_pointcloud2 call (same n_points / point_step). */
int gie_ogm_vlp16_pointcloud2_host(gie_locmap *lm, gie_hashmap *hm, const void *data_host, int n_points, int point_step, int off_x,
                                   int off_y, int off_ring, int scan_num, int ring_num, float theta_inc, float theta_min,
                                   float phi_inc, float phi_min, int for_motion_planner, int rbt_r2_grids);
int gie_vlp16_last_ranges(gie_locmap *lm, int n_points, int point_step, int scan_num, int ring_num, float *ranges_host);
int gie_ogm_pointcloud2_host(gie_locmap *lm, gie_hashmap *hm, const void *data_host, int n_points, int point_step, int off_x,
                             int max_points, int for_motion_planner, int rbt_r2_grids);

/* GlbHashMap::updateHashOGM (glb_hash_map.cu:115-143) incl. allocHashTB (:58-113).
 *   stream_glb_ogm : record blocks whose voxel type changed, for gie_hashmap_stream_changed (unify_helper.cuh:103-113)
 *   n_obs, obs_*   : Ext_Obs_Wrapper's boxes (include/map_structure/pre_map.h:12-28), host arrays float[3*n_obs] lower-left /
 *                    upper-right corners in metres and unsigned char[n_obs] activation flags; box 0 is the outer fence
 *                    (voxels OUTSIDE it become obstacles), boxes 1.. are obstacles inside (unify_helper.cuh:68-86,149-162).
 *                    n_obs = 0 or no activated box = the reference's shipped default (pre_map.cu:85). */
int gie_hashmap_update_ogm(gie_hashmap *hm, int input_pntcld, int map_ct, int stream_glb_ogm, int n_obs,
                           const float *obs_ll_host, const float *obs_ur_host, const unsigned char *obs_activated_host);
/* EDT_OCC::batchEDTUpdate (src/kernel/edt/local_edt.cu:7-28); cuTT plans are not needed */
int gie_edt_batch_update(gie_locmap *lm);
/* The two halves of gie_edt_batch_update as separate calls, for volumes sharded across GPUs (no reference counterpart: the
 * reference is single-GPU).  gie_edt_xy_sweeps runs EDTphase1+2 on this map's slices (local to a z-slab) and leaves
 * GIE_ARR_EDT_G2 / _CXY / _NCOLS; gie_edt_z_sweep runs EDTphase3 over whatever those three arrays hold — after the caller
 * re-partitioned them from z-slabs to y-slabs (gie-mapping_b200/sharded.py does it with NCCL all-to-all) — and writes
 * GIE_ARR_AUX / GIE_ARR_COC_AUX.  max_width_override = X+Y+Z of the WHOLE volume (the "sees nothing" sentinel,
 * local_batch.h:44), 0 = this map's own. */
int gie_edt_xy_sweeps(gie_locmap *lm);
int gie_edt_z_sweep(gie_locmap *lm, int max_width_override);
/* ---- one volume sharded over several GPUs (no reference counterpart) -------------------------------------------
 * The dense half of the frame — the x and z sweeps of the batch EDT, which write 8 B per voxel of the volume — is cut into
 * slabs of rows y, one per GPU; the sparse half (ray cast, hash merge, wavefronts) stays with the map that owns the hashed
 * global map and reads the batch-EDT result out of the slabs, its own or a peer GPU's (CUDA IPC mapping over NVLink).
 *   gie_locmap_create_slab  : a map that holds only the batch-EDT arrays of rows [row0, row0 + rows) (row0 % 32 == 0)
 *   gie_slab_alias_inputs   : same device as the owner map: read its ytab / column lists directly
 *   gie_slab_input_buffers  : other device: the buffers to receive them into (e.g. with an NCCL broadcast);
 *   gie_slab_set_compact    :   compact = only the planes of the obstacle-bearing slices were sent, in slice order
 *   gie_edt_pack            : owner map: the y pass of the whole volume (EDTphase1 as bit words + links, column and slice lists);
 *                             optionally gathers the planes of the obstacle-bearing slices into contiguous send buffers
 *   gie_edt_slab_sweeps     : EDTphase2 + EDTphase3 on the slab's rows; max_width = X+Y+Z of the whole volume (0 = own)
 *   gie_ipc_export          : CUDA IPC handle (64 bytes) of a slab's output array, for the owner process
 *   gie_locmap_attach_slabs : owner map: use these slabs as the batch-EDT result (device pointers for slabs of this process, IPC
 *                             handles [64 * n_slabs] for the others, pointer NULL); frees its own whole-volume copies */
int gie_locmap_create_slab(gie_locmap **out, int size_x, int size_y, int size_z, int row0, int rows);
int gie_slab_alias_inputs(gie_locmap *slab, gie_locmap *owner);
int gie_slab_input_buffers(gie_locmap *slab, void **ytab_dev, size_t *ytab_bytes, void **col_list_dev, size_t *col_bytes, void **meta_dev,
                           size_t *meta_bytes);
int gie_slab_set_compact(gie_locmap *slab, int compact);
int gie_edt_pack(gie_locmap *lm, void *ytab_compact_dev, void *col_compact_dev);
int gie_edt_slab_sweeps(gie_locmap *slab, int max_width);
int gie_ipc_export(void *dev_ptr, unsigned char handle64[64]);
int gie_locmap_attach_slabs(gie_locmap *lm, int n_slabs, int slab_rows, void *const *aux_dev, void *const *coc_dev,
                            const unsigned char *aux_handles, const unsigned char *coc_handles);

/* GlbHashMap::mergeNewObsv (glb_hash_map.cu:146-207).  display_glb_edt: record blocks whose distances changed
 * (wave_core.cuh:128-134,250-256; unify_helper.cuh:510-520) for gie_hashmap_stream_changed. */
int gie_hashmap_merge_new_obsv(gie_hashmap *hm, int map_ct, int display_glb_edt);
/* GlbHashMap::streamPipeline + streamD2H (glb_hash_map.cu:209-247): the blocks recorded as changed since the last call,
 * gathered on the device into the reference's GlbVoxel layout and copied with ONE device->host transfer (the reference
 * issues one blocking 20 KB copy per block).  keys_host = int[3*max_blocks], voxels_host = gie_glbvoxel[512*max_blocks];
 * *n_out blocks are returned and un-flagged; blocks that did not fit stay flagged.  gie_hashmap_num_changed = how many
 * are pending. */
int gie_hashmap_num_changed(gie_hashmap *hm, int *n);
int gie_hashmap_stream_changed(gie_hashmap *hm, int32_t *keys_host, gie_glbvoxel *voxels_host, int max_blocks, int *n_out);
/* waits for the stream and returns the sticky device-side status (queue overflow, out of blocks) */
int gie_sync(gie_hashmap *hm);

/* number of allocated voxel blocks; export in the reference's GlbVoxel layout (README.md:163-170 contract):
 * keys_host = int[3*n], voxels_host = gie_glbvoxel[512*n] with voxel index (x&7)*64+(y&7)*8+(z&7) */
int gie_hashmap_num_blocks(gie_hashmap *hm, int *n);
int gie_hashmap_export_blocks(gie_hashmap *hm, int32_t *keys_host, gie_glbvoxel *voxels_host, int max_blocks);
/* Device-side view of the global map for GPU consumers (README.md:163-170: "Each voxel can be retrieved by using device
 * function get_VB_key() and get_voxID_in_VB()").  Raw device pointers of the engine's open-addressing block table and of its
 * FIELD-MAJOR voxel pools (index = block * 512 + (z&7)*64 + (y&7)*8 + (x&7)); include/gie_compat/par_wave/gie_device_view.cuh
 * holds the matching __device__ accessors (gie_dv_find_block, gie_dv_voxel -> GlbVoxel).  Valid until the map is destroyed;
 * reads must be ordered after the engine's work on its stream. */
typedef struct gie_device_view {
    const unsigned long long *keys;   /* packed block key (3 x 21 bit) or ~0 */
    const int32_t *vals;              /* block index of the key slot */
    uint32_t cap_mask;
    int32_t block_max;
    const int32_t *block_count;
    const unsigned char *occ_val;
    const signed char *vox_type;
    const int32_t *update_ct;
    const unsigned long long *coc_glb; /* 21 bits per axis, biased by 2^20 */
    const int32_t *dist_sq;
    const int32_t *wave_layer;
    const unsigned long long *pair;    /* (dist_sq << 32) | wave-range coc id */
} gie_device_view;
int gie_hashmap_device_view(gie_hashmap *hm, gie_device_view *out);

/* frontier sizes / BFS levels of the last merge: {fA, fB, fC, levelsA, levelsB, levelsC, fB_after_A, fC_after_B} */
int gie_hashmap_wave_stats(gie_hashmap *hm, int64_t out8[8]);

/* ---- self-validation ---------------------------------------------------------------------------- */
/* Gnd_truth_checker::cmp_dist (include/gt_checker.h:30-80) evaluated on the device instead of through PCL clouds and a FLANN
 * KD-tree on the host (include/volumetric_mapper.h:181-356).  For every voxel of the EDT cloud the exact distance to the
 * nearest OCCUPIED voxel of the GLOBAL map is found (pruned brute force over voxel blocks) and compared with the distance
 * the map holds: error = (nearest - edt) * voxel_width in metres.
 *   mode 0 = profile_loc_rms: the known voxels of the local volume, edt = _edt_D
 *   mode 1 = profile_glb_rms: every known hash voxel with a valid dist_sq, edt = sqrt(dist_sq)
 *   truth_sq_host (mode 0 only, may be NULL): int32[X*Y*Z], squared nearest-obstacle distance in voxels per local voxel,
 *                 -1 where the voxel is not part of the EDT cloud
 * rms = sqrt(sum_sq / n) is cmp_dist's return value (-1 when either cloud is empty); edt_less / edt_more count the voxels
 * whose EDT is more than 1 mm below / above the nearest-obstacle distance (l_cnt / m_cnt, gt_checker.h:55-58). */
typedef struct gie_edt_check {
    long long n;            /* voxels compared */
    long long n_occupied;   /* OCCUPIED voxels in the global map */
    long long edt_less, edt_more;
    double sum_abs, sum_sq, max_abs, rms;
} gie_edt_check;
int gie_hashmap_check_edt(gie_hashmap *hm, int mode, int32_t *truth_sq_host, gie_edt_check *out);

/* ---- measurement -------------------------------------------------------------------------------- */
/* When enabled, CUDA events bracket each stage on the engine stream. */
enum gie_stage { GIE_ST_OGM = 0, GIE_ST_HASH_MERGE = 1, GIE_ST_EDT_PACK = 2, GIE_ST_EDT_X = 3, GIE_ST_EDT_Z = 4,
                 GIE_ST_MARK_FRONTIER = 5, GIE_ST_WAVES = 6, GIE_ST_COMMIT = 7, GIE_ST_COUNT = 8 };
int gie_profile_enable(gie_locmap *lm, int on);
/* milliseconds of the last frame's stages (synchronises the stream) */
int gie_profile_last(gie_locmap *lm, float ms_out[GIE_ST_COUNT]);
/* number of kernels this library launched since creation */
int gie_launch_count(gie_locmap *lm, long long *n);

/* diagnostics: per wave-C BFS level {frontier size, 5 globaltimer stamps in ns}; only when the map was created with the
 * environment variable GIE_WAVE_TRACE set.  out = uint64[6 * max_levels] */
int gie_debug_wave_trace(gie_hashmap *hm, unsigned long long *out, int max_levels);

/* warmupCuda (include/warmup.h:9) */
int gie_warmup(void);

#ifdef __cplusplus
}
#endif
#endif /* GIE_B200_H */
