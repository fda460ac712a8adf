"""I/O with oracle/_ref/ref_driver_* (the reference's own CUDA sources, see build_ref.sh).  Test infrastructure only."""
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SENSOR_ID = {"pointcloud": 0, "scan2d": 1, "vlp16": 2, "depth": 3}
PAYLOAD_KEY = {"pointcloud": "points", "scan2d": "scan", "vlp16": "ranges", "depth": "depth"}
BOXVOX_DTYPE = np.dtype([("alloc", np.int8), ("type", np.int8), ("occ", np.uint8), ("pad", np.uint8), ("dist", np.int32),
                         ("coc", np.int32, 3)])


def driver_path(flavour):
    return os.path.join(_HERE, "_ref", f"ref_driver_{flavour}")


def available(flavour="parity"):
    return os.path.exists(driver_path(flavour))


def write_input(path, cfg, frames):
    X, Y, Z = cfg["local_size"]
    sp = cfg.get("scan_param", {})
    cp = cfg.get("cam_param", {})
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


def iter_output(path, cfg, nframes, halo):
    """Yields the per-frame outputs one at a time (a 512^3 frame is 2.8 GB)."""
    X, Y, Z = cfg["local_size"]
    n = X * Y * Z
    box = (Z + 2 * halo, Y + 2 * halo, X + 2 * halo)
    with open(path, "rb") as f:
        for _ in range(nframes):
            d = {}
            d["glb_type"] = np.fromfile(f, np.int8, n).reshape(Z, Y, X)
            for k in ["aux", "coc_aux", "pair_dist", "pair_id"]:
                d[k] = np.fromfile(f, np.int32, n).reshape(Z, Y, X)
            d["edt"] = np.fromfile(f, np.float32, n).reshape(Z, Y, X)
            nb = box[0] * box[1] * box[2]
            d["box"] = np.fromfile(f, BOXVOX_DTYPE, nb).reshape(box)
            yield d


def read_output(path, cfg, nframes, halo):
    return list(iter_output(path, cfg, nframes, halo))


def run_to_file(cfg, frames, flavour="parity", halo=4, workdir="/tmp"):
    """Runs the reference driver and returns the path of its output file (read it with iter_output, then delete it)."""
    inp = os.path.join(workdir, f"gie_ref_in_{os.getpid()}.bin")
    outp = os.path.join(workdir, f"gie_ref_out_{os.getpid()}.bin")
    write_input(inp, cfg, frames)
    res = subprocess.run([driver_path(flavour), inp, outp, "--halo", str(halo)], capture_output=True, text=True)
    os.remove(inp)
    if res.returncode != 0:
        raise RuntimeError(f"reference driver failed ({res.returncode}): {res.stdout[-2000:]} {res.stderr[-2000:]}")
    return outp


def run(cfg, frames, flavour="parity", halo=4, workdir="/tmp", timing=False):
    """Runs the reference driver; returns per-frame outputs (timing=False) or the list of (ogm_ms, edt_ms)."""
    inp = os.path.join(workdir, f"gie_ref_in_{os.getpid()}.bin")
    outp = os.path.join(workdir, f"gie_ref_out_{os.getpid()}.bin")
    write_input(inp, cfg, frames)
    cmd = [driver_path(flavour), inp, "-" if timing else outp, "--halo", str(halo)] + (["--time"] if timing else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    os.remove(inp)
    if res.returncode != 0:
        raise RuntimeError(f"reference driver failed ({res.returncode}): {res.stdout[-2000:]} {res.stderr[-2000:]}")
    if timing:
        t = []
        for line in res.stdout.splitlines():
            if line.startswith("frame "):
                p = line.split()
                t.append((float(p[3]), float(p[5])))
        return t
    out = read_output(outp, cfg, len(frames), halo)
    os.remove(outp)
    return out
