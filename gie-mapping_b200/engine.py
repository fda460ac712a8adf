"""ctypes host mirror of the reference's operator surface for the per-frame hot path.

Names follow the reference (paths in the reference repo):
  LocMap                      include/map_structure/local_batch.h:32-569
  GlbHashMap                  include/par_wave/glb_hash_map.h:11-65
  GlbHashMap.updateHashOGM    src/kernel/par_wave/glb_hash_map.cu:115-143
  GlbHashMap.mergeNewObsv     src/kernel/par_wave/glb_hash_map.cu:146-207
  batchEDTUpdate              src/kernel/edt/local_edt.cu:7-28
  *localOGMKernels            src/kernel/{point_cloud,hokuyo,vlp16,realsense}/*
  Mapper.publishMap           src/volumetric_mapper.cpp:138-224 (call order only; no ROS)

Every compute call goes through the C ABI of libgie_b200.so.  There is no fallback path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ARR_RAY_COUNT, ARR_INST_TYPE, ARR_GLB_TYPE, ARR_EDT, ARR_AUX, ARR_COC_AUX, ARR_PAIR = range(7)
ARR_EDT_G2, ARR_EDT_CXY, ARR_EDT_NCOLS = 7, 8, 9
_ARR_DTYPE = {ARR_RAY_COUNT: np.int32, ARR_INST_TYPE: np.int8, ARR_GLB_TYPE: np.int8, ARR_EDT: np.float32,
              ARR_AUX: np.int32, ARR_COC_AUX: np.int32, ARR_PAIR: np.uint64}
STAGE_NAMES = ["ogm", "hash_merge", "edt_pack", "edt_x", "edt_z", "mark_frontier", "waves", "commit"]

GLBVOXEL_DTYPE = np.dtype([("occ_val", np.uint8), ("vox_type", np.int8), ("_pad", np.int16), ("update_ct", np.int32),
                           ("coc_glb", np.int32, 3), ("dist_sq", np.int32), ("wave_layer", np.int32),
                           ("_pad2", np.int32), ("dist_id_pair", np.uint64)])
SEENDIST_DTYPE = np.dtype([("d", np.float32), ("s", np.uint8), ("o", np.uint8), ("_pad", np.uint16)])


class GieError(RuntimeError):
    pass


class EdtCheck(C.Structure):
    """gie_edt_check (include/gie_b200.h)."""
    _fields_ = [("n", C.c_longlong), ("n_occupied", C.c_longlong), ("edt_less", C.c_longlong), ("edt_more", C.c_longlong),
                ("sum_abs", C.c_double), ("sum_sq", C.c_double), ("max_abs", C.c_double), ("rms", C.c_double)]


def library_path():
    return os.path.join(_HERE, "libgie_b200.so")


def load_library():
    """Load libgie_b200.so from the package directory.  Raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise GieError(f"{path} is missing: build it with `make -C gie-mapping_b200/csrc` or __graft_entry__.build(); "
                       "this engine has no CPU fallback")
    lib = C.CDLL(path)
    lib.gie_last_error.restype = C.c_char_p
    lib.gie_version.restype = C.c_char_p
    p, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "gie_locmap_create": [C.POINTER(p), f, i, i, i, C.c_ubyte, f, f, i, i],
        "gie_locmap_destroy": [p], "gie_set_stream": [p, p], "gie_locmap_set_pose": [p, p, p],
        "gie_locmap_get_pivots": [p, p, p], "gie_locmap_copy_ogm_to_host": [p, p], "gie_locmap_copy_edt_to_host": [p, p],
        "gie_locmap_convert_costmap": [p, p], "gie_locmap_device_ptr": [p, i, C.POINTER(p), C.POINTER(C.c_size_t)],
        "gie_locmap_download": [p, i, p], "gie_locmap_upload_glb_type": [p, p],
        "gie_hashmap_create": [C.POINTER(p), p, i, i], "gie_hashmap_destroy": [p],
        "gie_ogm_pointcloud_dev": [p, p, p, i, i, i], "gie_ogm_pointcloud_host": [p, p, p, i, i, i],
        "gie_ogm_scan2d_dev": [p, p, p, i, f, f, i, i], "gie_ogm_scan2d_host": [p, p, p, i, f, f, i, i],
        "gie_ogm_vlp16_dev": [p, p, p, i, i, f, f, f, f, i, i], "gie_ogm_vlp16_host": [p, p, p, i, i, f, f, f, f, i, i],
        "gie_ogm_depth_dev": [p, p, p, i, i, f, f, f, f, i, i, i], "gie_ogm_depth_host": [p, p, p, i, i, f, f, f, f, i, i, i],
        "gie_hashmap_update_ogm": [p, i, i, i, i, p, p, p], "gie_edt_batch_update": [p], "gie_hashmap_merge_new_obsv": [p, i, i],
        "gie_hashmap_num_changed": [p, C.POINTER(i)], "gie_hashmap_stream_changed": [p, p, p, i, C.POINTER(i)],
        "gie_ogm_vlp16_pointcloud2_host": [p, p, p, i, i, i, i, i, i, i, f, f, f, f, i, i],
        "gie_vlp16_last_ranges": [p, i, i, i, i, p], "gie_ogm_pointcloud2_host": [p, p, p, i, i, i, i, i, i],
        "gie_edt_xy_sweeps": [p], "gie_edt_z_sweep": [p, i],
        "gie_make_projection": [p, p, p, p], "gie_locmap_set_projection": [p, p, p, p], "gie_locmap_calculate_pivots": [p, p],
        "gie_sync": [p], "gie_hashmap_num_blocks": [p, C.POINTER(i)], "gie_hashmap_export_blocks": [p, p, p, i],
        "gie_hashmap_wave_stats": [p, p], "gie_profile_enable": [p, i], "gie_profile_last": [p, p],
        "gie_launch_count": [p, C.POINTER(C.c_longlong)], "gie_warmup": [],
        "gie_hashmap_check_edt": [p, i, p, p],
        "gie_locmap_create_slab": [C.POINTER(p), i, i, i, i, i], "gie_slab_alias_inputs": [p, p],
        "gie_slab_input_buffers": [p, C.POINTER(p), C.POINTER(C.c_size_t), C.POINTER(p), C.POINTER(C.c_size_t), C.POINTER(p), C.POINTER(C.c_size_t)],
        "gie_slab_set_compact": [p, i], "gie_edt_pack": [p, p, p], "gie_edt_slab_sweeps": [p, i], "gie_ipc_export": [p, p],
        "gie_locmap_attach_slabs": [p, i, i, p, p, p, p],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _LIB = lib
    return lib


def _check(rc):
    if rc != 0:
        raise GieError(f"gie status {rc}: {load_library().gie_last_error().decode()}")


def _hostptr(a):
    return a.ctypes.data_as(C.c_void_p)


class LocMap:
    """Dense local volume.  Reference: class LocMap (include/map_structure/local_batch.h:35-60 ctor arguments)."""

    def __init__(self, voxel_size, local_size, occupancy_threshold=180, ogm_min_h=-10.0, ogm_max_h=10.0,
                 cutoff_grids_sq=100, fast_mode=False):
        self.lib = load_library()
        self._h = C.c_void_p()
        self._local_size = tuple(int(v) for v in local_size)
        self._voxel_width = float(voxel_size)
        X, Y, Z = self._local_size
        _check(self.lib.gie_locmap_create(C.byref(self._h), voxel_size, X, Y, Z, occupancy_threshold, ogm_min_h, ogm_max_h,
                                          cutoff_grids_sq, int(fast_mode)))
        self._map_volume = X * Y * Z
        self._bdr_num = 2 * (X * Y + Y * Z + X * Z)

    def close(self):
        if self._h:
            self.lib.gie_locmap_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        _check(self.lib.gie_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_pose(self, q_wxyz, t_xyz):
        """trans2proj + calculate_pivot_origin + calculate_update_pivot (volumetric_mapper.cpp:144-155)."""
        q = np.ascontiguousarray(q_wxyz, dtype=np.float32)
        t = np.ascontiguousarray(t_xyz, dtype=np.float32)
        _check(self.lib.gie_locmap_set_pose(self._h, _hostptr(q), _hostptr(t)))

    def pivots(self):
        out = np.zeros(6, np.int32)
        org = np.zeros(3, np.float32)
        _check(self.lib.gie_locmap_get_pivots(self._h, _hostptr(out), _hostptr(org)))
        return out[:3].copy(), out[3:].copy(), org

    def download(self, which):
        X, Y, Z = self._local_size
        out = np.empty(self._map_volume, dtype=_ARR_DTYPE[which])
        _check(self.lib.gie_locmap_download(self._h, which, _hostptr(out)))
        return out.reshape(Z, Y, X)

    def device_ptr(self, which):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _check(self.lib.gie_locmap_device_ptr(self._h, which, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def upload_glb_type(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.int8).reshape(-1)
        assert a.size == self._map_volume
        _check(self.lib.gie_locmap_upload_glb_type(self._h, _hostptr(a)))

    def copy_ogm_2_host(self):
        return self.download(ARR_GLB_TYPE)

    def copy_edt_2_host(self):
        return self.download(ARR_EDT)

    def convertCostMap(self):
        out = np.empty(self._map_volume, dtype=SEENDIST_DTYPE)
        _check(self.lib.gie_locmap_convert_costmap(self._h, _hostptr(out)))
        return out

    def batchEDTUpdate(self):
        """EDT_OCC::batchEDTUpdate (src/kernel/edt/local_edt.cu:7-28)."""
        _check(self.lib.gie_edt_batch_update(self._h))

    def edt_xy_sweeps(self):
        _check(self.lib.gie_edt_xy_sweeps(self._h))

    def edt_z_sweep(self, max_width_override=0):
        _check(self.lib.gie_edt_z_sweep(self._h, int(max_width_override)))

    def edt_slice_columns(self):
        """int32 [Z]: obstacle-bearing columns per z-slice found by the last batch EDT (0 = the sweeps skipped the slice)."""
        ptr, nbytes = self.device_ptr(ARR_EDT_NCOLS)
        out = np.empty(self._local_size[2], np.int32)
        _check(self.lib.gie_locmap_download(self._h, ARR_EDT_NCOLS, _hostptr(out)))
        return out

    def profile_enable(self, on=True):
        _check(self.lib.gie_profile_enable(self._h, int(on)))

    def profile_last(self):
        ms = np.zeros(len(STAGE_NAMES), np.float32)
        _check(self.lib.gie_profile_last(self._h, _hostptr(ms)))
        return dict(zip(STAGE_NAMES, ms.tolist()))

    def launch_count(self):
        n = C.c_longlong()
        _check(self.lib.gie_launch_count(self._h, C.byref(n)))
        return n.value


class GlbHashMap:
    """Voxel-block hashed global map.  Reference: struct GlbHashMap (include/par_wave/glb_hash_map.h:11-65)."""

    def __init__(self, loc_map, bucket_max=10000, block_max=19997):
        self.lib = loc_map.lib
        self._lMap = loc_map
        self._h = C.c_void_p()
        _check(self.lib.gie_hashmap_create(C.byref(self._h), loc_map._h, bucket_max, block_max))
        self.hash_table_H_std = {}   # block key -> GlbVoxel[512], filled by streamPipeline

    def close(self):
        if self._h:
            self.lib.gie_hashmap_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- OGM entry points (the reference passes VB_keys_loc_D; here the map handle) -----------------------------
    def ogm_pointcloud(self, pts, for_motion_planner=False, rbt_r2_grids=0, device_ptr=None, n=None):
        """PNTCLD_RAYCAST::localOGMKernels / PntcldMapMaker::updateLocalOGM.  pts: float32 [n,3] host array, or a
        device pointer (device_ptr, n)."""
        lm = self._lMap
        if device_ptr is not None:
            _check(self.lib.gie_ogm_pointcloud_dev(lm._h, self._h, C.c_void_p(int(device_ptr)), int(n), int(for_motion_planner), rbt_r2_grids))
        else:
            a = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
            _check(self.lib.gie_ogm_pointcloud_host(lm._h, self._h, _hostptr(a), a.shape[0], int(for_motion_planner), rbt_r2_grids))

    def ogm_scan2d(self, scan, theta_inc, theta_min, for_motion_planner=False, rbt_r2_grids=0, device_ptr=None):
        lm = self._lMap
        if device_ptr is not None:
            _check(self.lib.gie_ogm_scan2d_dev(lm._h, self._h, C.c_void_p(int(device_ptr)), int(scan), theta_inc, theta_min, int(for_motion_planner), rbt_r2_grids))
        else:
            a = np.ascontiguousarray(scan, dtype=np.float32).reshape(-1)
            _check(self.lib.gie_ogm_scan2d_host(lm._h, self._h, _hostptr(a), a.size, theta_inc, theta_min, int(for_motion_planner), rbt_r2_grids))

    def ogm_vlp16(self, ranges, theta_inc, theta_min, phi_inc, phi_min, for_motion_planner=False, rbt_r2_grids=0,
                  device_ptr=None, shape=None):
        lm = self._lMap
        if device_ptr is not None:
            ring_num, scan_num = shape
            _check(self.lib.gie_ogm_vlp16_dev(lm._h, self._h, C.c_void_p(int(device_ptr)), scan_num, ring_num, theta_inc, theta_min,
                                              phi_inc, phi_min, int(for_motion_planner), rbt_r2_grids))
        else:
            a = np.ascontiguousarray(ranges, dtype=np.float32)
            ring_num, scan_num = a.shape
            _check(self.lib.gie_ogm_vlp16_host(lm._h, self._h, _hostptr(a), scan_num, ring_num, theta_inc, theta_min, phi_inc,
                                               phi_min, int(for_motion_planner), rbt_r2_grids))

    def ogm_depth(self, img, cx, cy, fx, fy, valid_nan=True, for_motion_planner=False, rbt_r2_grids=0, device_ptr=None,
                  shape=None):
        lm = self._lMap
        if device_ptr is not None:
            rows, cols = shape
            _check(self.lib.gie_ogm_depth_dev(lm._h, self._h, C.c_void_p(int(device_ptr)), rows, cols, cx, cy, fx, fy,
                                              int(valid_nan), int(for_motion_planner), rbt_r2_grids))
        else:
            a = np.ascontiguousarray(img, dtype=np.float32)
            rows, cols = a.shape
            _check(self.lib.gie_ogm_depth_host(lm._h, self._h, _hostptr(a), rows, cols, cx, cy, fx, fy, int(valid_nan),
                                               int(for_motion_planner), rbt_r2_grids))

    def ogm_vlp16_pointcloud2(self, data, point_step, off_x, off_y, off_ring, scan_num, ring_num, theta_inc, theta_min, phi_inc,
                              phi_min, for_motion_planner=False, rbt_r2_grids=0):
        """Vlp16MapMaker::updateLocalOGM on raw PointCloud2 bytes (uint8 array of n * point_step)."""
        a = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        n = a.size // point_step
        _check(self.lib.gie_ogm_vlp16_pointcloud2_host(self._lMap._h, self._h, _hostptr(a), n, point_step, off_x, off_y, off_ring, scan_num,
                                                       ring_num, theta_inc, theta_min, phi_inc, phi_min, int(for_motion_planner), rbt_r2_grids))
        return n

    def vlp16_last_ranges(self, n_points, point_step, scan_num, ring_num):
        out = np.zeros((ring_num, scan_num), np.float32)
        _check(self.lib.gie_vlp16_last_ranges(self._lMap._h, n_points, point_step, scan_num, ring_num, _hostptr(out)))
        return out

    def ogm_pointcloud2(self, data, point_step, off_x=0, max_points=0, for_motion_planner=False, rbt_r2_grids=0):
        """PntcldMapMaker::updateLocalOGM on raw PointCloud2 bytes."""
        a = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        _check(self.lib.gie_ogm_pointcloud2_host(self._lMap._h, self._h, _hostptr(a), a.size // point_step, point_step, off_x, max_points,
                                                 int(for_motion_planner), rbt_r2_grids))

    # --- per-frame stages --------------------------------------------------------------------------------------
    def updateHashOGM(self, input_pynt, map_ct, stream_glb_ogm=False, ext_obsv=None):
        """ext_obsv: None or (ll float32 [n,3], ur float32 [n,3], activated uint8 [n]) — Ext_Obs_Wrapper's boxes."""
        if ext_obsv is None:
            _check(self.lib.gie_hashmap_update_ogm(self._h, int(input_pynt), map_ct, int(stream_glb_ogm), 0, None, None, None))
        else:
            ll = np.ascontiguousarray(ext_obsv[0], np.float32).reshape(-1, 3)
            ur = np.ascontiguousarray(ext_obsv[1], np.float32).reshape(-1, 3)
            act = np.ascontiguousarray(ext_obsv[2], np.uint8).reshape(-1)
            _check(self.lib.gie_hashmap_update_ogm(self._h, int(input_pynt), map_ct, int(stream_glb_ogm), ll.shape[0],
                                                   _hostptr(ll), _hostptr(ur), _hostptr(act)))

    def mergeNewObsv(self, map_ct, display_glb_edt=False):
        _check(self.lib.gie_hashmap_merge_new_obsv(self._h, map_ct, int(display_glb_edt)))

    def streamPipeline(self):
        """GlbHashMap::streamPipeline: (keys int32 [n,3], voxels GLBVOXEL_DTYPE [n,512]) of the blocks changed since the
        last call; merges them into the host mirror VB_keys_H / VB_values_H / hash_table_H_std."""
        n = C.c_int()
        _check(self.lib.gie_hashmap_num_changed(self._h, C.byref(n)))
        keys = np.zeros((n.value, 3), np.int32)
        vox = np.zeros((n.value, 512), dtype=GLBVOXEL_DTYPE)
        if n.value:
            _check(self.lib.gie_hashmap_stream_changed(self._h, _hostptr(keys), _hostptr(vox), n.value, C.byref(n)))
            keys, vox = keys[:n.value], vox[:n.value]
        for k, v in zip(keys.tolist(), vox):
            self.hash_table_H_std[tuple(k)] = v
        return keys, vox

    def sync(self):
        _check(self.lib.gie_sync(self._h))

    def num_blocks(self):
        n = C.c_int()
        _check(self.lib.gie_hashmap_num_blocks(self._h, C.byref(n)))
        return n.value

    def export_blocks(self):
        """Returns (keys int32 [n,3], voxels GLBVOXEL_DTYPE [n,512]) in the reference's layout and voxel order."""
        n = self.num_blocks()
        keys = np.zeros((n, 3), np.int32)
        vox = np.zeros((n, 512), dtype=GLBVOXEL_DTYPE)
        if n:
            _check(self.lib.gie_hashmap_export_blocks(self._h, _hostptr(keys), _hostptr(vox), n))
        return keys, vox

    def check_edt(self, glb=False, want_truth=False):
        """Gnd_truth_checker::cmp_dist (include/gt_checker.h:30-80) on the device.  Returns a dict of the error statistics
        (metres) and, with want_truth (local mode only), the squared nearest-obstacle distance per local voxel (-1 = unchecked)."""
        res = EdtCheck()
        truth = None
        if want_truth:
            X, Y, Z = self._lMap._local_size
            truth = np.empty(X * Y * Z, np.int32)
        _check(self.lib.gie_hashmap_check_edt(self._h, int(glb), _hostptr(truth) if truth is not None else None, C.byref(res)))
        out = {k: getattr(res, k) for k, _ in EdtCheck._fields_}
        if truth is not None:
            X, Y, Z = self._lMap._local_size
            out["truth_sq"] = truth.reshape(Z, Y, X)
        return out

    def wave_stats(self):
        out = np.zeros(8, np.int64)
        _check(self.lib.gie_hashmap_wave_stats(self._h, _hostptr(out)))
        return dict(zip(["fA", "fB", "fC", "levelsA", "levelsB", "levelsC", "fB_after_A", "fC_after_B"], out.tolist()))


class Mapper:
    """ROS-free replay of VOLMAPNODE::publishMap's call order (src/volumetric_mapper.cpp:138-224)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.loc_map = LocMap(cfg["voxel_width"], cfg["local_size"], cfg.get("occupancy_threshold", 180),
                              cfg.get("ogm_min_h", -10.0), cfg.get("ogm_max_h", 10.0), cfg["cutoff_grids_sq"],
                              cfg.get("fast_mode", False))
        self.hash_map = GlbHashMap(self.loc_map, cfg.get("bucket_max", 10000), cfg.get("block_max", 19997))
        self._time = 0

    def close(self):
        self.hash_map.close()
        self.loc_map.close()

    def integrate(self, frame, device_input=None):
        """OGM half of publishMap: pose, sensor integration, updateHashOGM."""
        cfg = self.cfg
        self._time += 1
        self.loc_map.set_pose(frame["q"], frame["t"])
        fmp, r2 = cfg.get("for_motion_planner", False), cfg.get("robot_r2_grids", 0)
        s = cfg["sensor"]
        hm = self.hash_map
        if s == "pointcloud":
            if device_input is not None:
                hm.ogm_pointcloud(None, fmp, r2, device_ptr=device_input, n=frame["points"].shape[0])
            else:
                hm.ogm_pointcloud(frame["points"], fmp, r2)
        elif s == "scan2d":
            sp = cfg["scan_param"]
            if device_input is not None:
                hm.ogm_scan2d(sp["scan_num"], sp["theta_inc"], sp["theta_min"], fmp, r2, device_ptr=device_input)
            else:
                hm.ogm_scan2d(frame["scan"], sp["theta_inc"], sp["theta_min"], fmp, r2)
        elif s == "vlp16":
            sp = cfg["scan_param"]
            if device_input is not None:
                hm.ogm_vlp16(None, sp["theta_inc"], sp["theta_min"], sp["phi_inc"], sp["phi_min"], fmp, r2,
                             device_ptr=device_input, shape=(sp["ring_num"], sp["scan_num"]))
            else:
                hm.ogm_vlp16(frame["ranges"], sp["theta_inc"], sp["theta_min"], sp["phi_inc"], sp["phi_min"], fmp, r2)
        elif s == "depth":
            cp = cfg["cam_param"]
            if device_input is not None:
                hm.ogm_depth(None, cp["cx"], cp["cy"], cp["fx"], cp["fy"], cp.get("valid_NaN", True), fmp, r2,
                             device_ptr=device_input, shape=(cp["rows"], cp["cols"]))
            else:
                hm.ogm_depth(frame["depth"], cp["cx"], cp["cy"], cp["fx"], cp["fy"], cp.get("valid_NaN", True), fmp, r2)
        else:
            raise GieError(f"unknown sensor {s}")
        hm.updateHashOGM(s == "pointcloud", self._time, cfg.get("display_glb_ogm", False) and not cfg.get("display_glb_edt", False),
                         frame.get("ext_obs"))

    def update_edt(self):
        """EDT half of publishMap: batchEDTUpdate + mergeNewObsv."""
        self.loc_map.batchEDTUpdate()
        self.hash_map.mergeNewObsv(self._time, self.cfg.get("display_glb_edt", False))

    def publishMap(self, frame, device_input=None):
        self.integrate(frame, device_input)
        self.update_edt()
