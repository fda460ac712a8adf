"""ctypes binding of the CPU oracle (oracle/gie_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GVOX_DTYPE = np.dtype([("occ_val", np.uint8), ("vox_type", np.int8), ("_pad", np.int16), ("update_ct", np.int32),
                       ("coc_glb", np.int32, 3), ("dist_sq", np.int32), ("wave_layer", np.int32), ("_pad2", np.int32),
                       ("dist_id_pair", np.uint64)])


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libgie_oracle.so")
        if not os.path.exists(path):
            build()
        l = C.CDLL(path)
        l.gor_create.restype = C.c_void_p
        l.gor_create.argtypes = [C.c_int] * 3 + [C.c_float, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
        for name in ["gor_ray_count", "gor_inst_type", "gor_glb_type", "gor_edt", "gor_aux", "gor_coc_aux", "gor_pair",
                     "gor_touched", "gor_stats"]:
            getattr(l, name).restype = C.c_void_p
            getattr(l, name).argtypes = [C.c_void_p]
        l.gor_destroy.argtypes = [C.c_void_p]
        l.gor_set_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.gor_ogm_pointcloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        l.gor_ogm_scan2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
        l.gor_ogm_vlp16.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int, C.c_int]
        l.gor_ogm_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int] * 3
        l.gor_update_hash_ogm.argtypes = [C.c_void_p, C.c_int, C.c_int]
        l.gor_batch_edt.argtypes = [C.c_void_p]
        l.gor_batch_edt_bruteforce.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.gor_merge_new_obsv.argtypes = [C.c_void_p, C.c_int]
        l.gor_get_pivots.argtypes = [C.c_void_p, C.c_void_p]
        l.gor_vlp16_bin.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_float, C.c_void_p]
        l.gor_set_ext_obs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        l.gor_set_stream.argtypes = [C.c_void_p, C.c_int, C.c_int]
        l.gor_take_changed.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        l.gor_num_blocks.argtypes = [C.c_void_p]
        l.gor_export_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.gor_set_threads.argtypes = [C.c_int]
        l.gor_cuda_atan2f.restype = C.c_float
        l.gor_cuda_atan2f.argtypes = [C.c_float, C.c_float]
        _LIB = l
    return _LIB


def set_threads(n):
    """Host threads for the oracle's batch EDT (OpenMP over columns); everything else is serial."""
    lib().gor_set_threads(int(n))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleMapper:
    """Same call order as gie_mapping_b200.Mapper, computed by the C oracle."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.l = lib()
        X, Y, Z = cfg["local_size"]
        self.shape = (Z, Y, X)
        self.n = X * Y * Z
        self.h = self.l.gor_create(X, Y, Z, cfg["voxel_width"], cfg.get("occupancy_threshold", 180), cfg.get("ogm_min_h", -10.0),
                                   cfg.get("ogm_max_h", 10.0), cfg["cutoff_grids_sq"], int(cfg.get("fast_mode", False)))
        self._time = 0

    def close(self):
        if self.h:
            self.l.gor_destroy(self.h)
            self.h = None

    def _view(self, fn, dtype):
        ptr = getattr(self.l, fn)(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(self.n,)).reshape(self.shape)

    ray_count = property(lambda s: s._view("gor_ray_count", np.int32))
    inst_type = property(lambda s: s._view("gor_inst_type", np.int8))
    glb_type = property(lambda s: s._view("gor_glb_type", np.int8))
    edt = property(lambda s: s._view("gor_edt", np.float32))
    aux = property(lambda s: s._view("gor_aux", np.int32))
    coc_aux = property(lambda s: s._view("gor_coc_aux", np.int32))
    pair = property(lambda s: s._view("gor_pair", np.uint64))

    def stats(self):
        ptr = self.l.gor_stats(self.h)
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(8,))
        return dict(zip(["fA", "fB", "fC", "levelsA", "levelsB", "levelsC", "fB_after_A", "fC_after_B"], a.tolist()))

    def pivots(self):
        out = np.zeros(6, np.int32)
        self.l.gor_get_pivots(self.h, _p(out))
        return out[:3].copy(), out[3:].copy()

    def set_glb_type(self, arr):
        self.glb_type[...] = np.asarray(arr, np.int8).reshape(self.shape)

    def integrate(self, frame):
        cfg = self.cfg
        self._time += 1
        q = np.ascontiguousarray(frame["q"], np.float32)
        t = np.ascontiguousarray(frame["t"], np.float32)
        self.l.gor_set_pose(self.h, _p(q), _p(t))
        fmp, r2 = int(cfg.get("for_motion_planner", False)), cfg.get("robot_r2_grids", 0)
        s = cfg["sensor"]
        if s == "pointcloud":
            a = np.ascontiguousarray(frame["points"], np.float32)
            self.l.gor_ogm_pointcloud(self.h, _p(a), a.shape[0], fmp, r2)
        elif s == "scan2d":
            sp = cfg["scan_param"]
            a = np.ascontiguousarray(frame["scan"], np.float32)
            self.l.gor_ogm_scan2d(self.h, _p(a), a.size, sp["theta_inc"], sp["theta_min"], fmp, r2)
        elif s == "vlp16":
            sp = cfg["scan_param"]
            a = np.ascontiguousarray(frame["ranges"], np.float32)
            self.l.gor_ogm_vlp16(self.h, _p(a), sp["scan_num"], sp["ring_num"], sp["theta_inc"], sp["theta_min"], sp["phi_inc"],
                                 sp["phi_min"], fmp, r2)
        elif s == "depth":
            cp = cfg["cam_param"]
            a = np.ascontiguousarray(frame["depth"], np.float32)
            self.l.gor_ogm_depth(self.h, _p(a), cp["rows"], cp["cols"], cp["cx"], cp["cy"], cp["fx"], cp["fy"],
                                 int(cp.get("valid_NaN", True)), fmp, r2)
        else:
            raise KeyError(s)
        self.l.gor_set_stream(self.h, int(cfg.get("display_glb_ogm", False) and not cfg.get("display_glb_edt", False)),
                              int(cfg.get("display_glb_edt", False)))
        eo = frame.get("ext_obs")
        if eo is not None:
            ll = np.ascontiguousarray(eo[0], np.float32).reshape(-1, 3)
            ur = np.ascontiguousarray(eo[1], np.float32).reshape(-1, 3)
            act = np.ascontiguousarray(eo[2], np.uint8).reshape(-1)
            self.l.gor_set_ext_obs(self.h, ll.shape[0], _p(ll), _p(ur), _p(act))
        else:
            self.l.gor_set_ext_obs(self.h, 0, None, None, None)
        self.l.gor_update_hash_ogm(self.h, int(s == "pointcloud"), self._time)

    def batch_edt(self):
        self.l.gor_batch_edt(self.h)

    def update_edt(self):
        self.l.gor_batch_edt(self.h)
        self.l.gor_merge_new_obsv(self.h, self._time)

    def publishMap(self, frame):
        self.integrate(frame)
        self.update_edt()

    def take_changed(self):
        """Block keys recorded as changed since the last call (streamPipeline's key set), int32 [n,3]."""
        n = self.l.gor_take_changed(self.h, None, 0)
        keys = np.zeros((n, 3), np.int32)
        if n:
            self.l.gor_take_changed(self.h, _p(keys), n)
        return keys

    def export_blocks(self):
        """Blocks in the reference layout; dist_id_pair in the reference's word order (low word = dist_sq)."""
        n = self.l.gor_num_blocks(self.h)
        keys = np.zeros((n, 3), np.int32)
        vox = np.zeros((n, 512), dtype=GVOX_DTYPE)
        if n:
            self.l.gor_export_blocks(self.h, _p(keys), _p(vox))
            p = vox["dist_id_pair"]
            vox["dist_id_pair"] = (p >> np.uint64(32)) | (p << np.uint64(32))
        return keys, vox


def vlp16_bin(data, point_step, off_x, off_y, off_ring, scan_num, ring_num, theta_inc):
    """Range image [ring_num, scan_num] from raw PointCloud2 bytes (Vlp16MapMaker::convertPyntCld)."""
    a = np.ascontiguousarray(data, np.uint8).reshape(-1)
    out = np.zeros((ring_num, scan_num), np.float32)
    lib().gor_vlp16_bin(_p(a), a.size // point_step, point_step, off_x, off_y, off_ring, scan_num, ring_num, theta_inc, _p(out))
    return out


def batch_edt_bruteforce(glb_type):
    """glb_type int8 [Z,Y,X] -> (dist_sq, coc) by the O(N * L) statement of the contract (small volumes only)."""
    Z, Y, X = glb_type.shape
    t = np.ascontiguousarray(glb_type, np.int8)
    d = np.zeros((Z, Y, X), np.int32)
    c = np.zeros((Z, Y, X), np.int32)
    lib().gor_batch_edt_bruteforce(_p(t), X, Y, Z, _p(d), _p(c))
    return d, c
