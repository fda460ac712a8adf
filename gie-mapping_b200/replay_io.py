"""Frame files for the C++ host driver gie-mapping_b200/gie_replay (format documented in host/gie_replay.cpp)."""
import os
import struct
import subprocess

import numpy as np

from .engine import GLBVOXEL_DTYPE, SEENDIST_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
SENSOR_ID = {"pointcloud": 0, "scan2d": 1, "vlp16": 2, "depth": 3}
PAYLOAD_KEY = {"pointcloud": "points", "scan2d": "scan", "vlp16": "ranges", "depth": "depth"}


def replay_binary():
    return os.path.join(_HERE, "gie_replay")


def write_frames(path, cfg, frames):
    X, Y, Z = cfg["local_size"]
    sp, cp = cfg.get("scan_param", {}), cfg.get("cam_param", {})
    ints = [0x47494531, SENSOR_ID[cfg["sensor"]], X, Y, Z, cfg.get("occupancy_threshold", 180), cfg["cutoff_grids_sq"],
            int(cfg.get("fast_mode", False)), cfg.get("bucket_max", 10000), cfg.get("block_max", 19997),
            int(cfg.get("for_motion_planner", False)), cfg.get("robot_r2_grids", 0), len(frames),
            cp.get("rows", 0), cp.get("cols", 0), sp.get("scan_num", 0), sp.get("ring_num", 0), int(cp.get("valid_NaN", True))]
    floats = [cfg["voxel_width"], cfg.get("ogm_min_h", -10.0), cfg.get("ogm_max_h", 10.0), sp.get("theta_inc", 0.0),
              sp.get("theta_min", 0.0), sp.get("phi_inc", 0.0), sp.get("phi_min", 0.0), cp.get("cx", 0.0), cp.get("cy", 0.0),
              cp.get("fx", 0.0), cp.get("fy", 0.0)]
    with open(path, "wb") as f:
        f.write(struct.pack(f"<{len(ints)}i{len(floats)}f", *ints, *floats))
        for fr in frames:
            f.write(np.asarray(fr["q"], np.float32).tobytes())
            f.write(np.asarray(fr["t"], np.float32).tobytes())
            p = np.ascontiguousarray(fr[PAYLOAD_KEY[cfg["sensor"]]], np.float32).ravel()
            f.write(struct.pack("<i", p.size))
            f.write(p.tobytes())


def read_output(path, cfg, nframes, stream=False, costmap=False):
    X, Y, Z = cfg["local_size"]
    n = X * Y * Z
    out = []
    with open(path, "rb") as f:
        for _ in range(nframes):
            d = {"glb_type": np.frombuffer(f.read(n), np.int8).reshape(Z, Y, X),
                 "aux": np.frombuffer(f.read(4 * n), np.int32).reshape(Z, Y, X),
                 "coc_aux": np.frombuffer(f.read(4 * n), np.int32).reshape(Z, Y, X),
                 "pair": np.frombuffer(f.read(8 * n), np.uint64).reshape(Z, Y, X),
                 "edt": np.frombuffer(f.read(4 * n), np.float32).reshape(Z, Y, X)}
            if costmap:
                d["costmap"] = np.frombuffer(f.read(SEENDIST_DTYPE.itemsize * n), SEENDIST_DTYPE).reshape(Z, Y, X)
            out.append(d)
        mirror = None
        if stream:
            nb = struct.unpack("<i", f.read(4))[0]
            mirror = {}
            for _ in range(nb):
                key = tuple(np.frombuffer(f.read(12), np.int32).tolist())
                mirror[key] = np.frombuffer(f.read(512 * GLBVOXEL_DTYPE.itemsize), GLBVOXEL_DTYPE)
    return out, mirror


def run_replay(cfg, frames, workdir, stream=False, costmap=False, timing=False, mapmakers=False):
    """Runs the C++ host driver on the frames; returns (per-frame arrays, streamed host mirror or None, stdout)."""
    exe = replay_binary()
    if not os.path.exists(exe):
        raise RuntimeError(f"{exe} is missing: build it with `make -C gie-mapping_b200/host`")
    inp, outp = os.path.join(workdir, "frames.bin"), os.path.join(workdir, "out.bin")
    write_frames(inp, cfg, frames)
    # timing runs dump nothing (the per-frame arrays of a 512^3 volume are 2.8 GB)
    cmd = [exe, inp, "-" if timing else outp] + (["--stream"] if stream else []) + (["--costmap"] if costmap else []) + (["--time"] if timing else []) + \
          (["--mapmakers"] if mapmakers else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gie_replay failed ({res.returncode}): {res.stdout[-1000:]} {res.stderr[-2000:]}")
    if timing:
        return None, None, res.stdout
    out, mirror = read_output(outp, cfg, len(frames), stream, costmap)
    return out, mirror, res.stdout
