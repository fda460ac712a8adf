"""Seeded synthetic scenes and sensor models.  Pure numpy; produces the inputs the reference's
localOGMKernels take: sensor-frame point clouds, 2-D scans, VLP-16 range images, depth images, and poses.

World: axis-aligned boxes (floor, optional ceiling, outer walls, K random boxes); ranges come from analytic
ray/AABB intersection so that they are exact.  The sensor moves +x at 0.5 m/frame with yaw 0.02 rad/frame and a small
constant pitch/roll so that the full rotation matrix is exercised.
"""
import math

import numpy as np


def quat_from_euler(roll, pitch, yaw):
    cr, sr = math.cos(roll / 2), math.sin(roll / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    q = np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                  cr * cp * sy - sr * sp * cy], dtype=np.float64)
    return (q / np.linalg.norm(q)).astype(np.float32)


def quat_to_rot(q):
    w, x, y, z = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


class World:
    def __init__(self, extent, height, n_boxes, seed=42, ceiling=True):
        """extent: half side length (m) of the walled area; height: ceiling height / wall height."""
        rng = np.random.RandomState(seed)
        e, h = float(extent), float(height)
        boxes = [(-e - 1, -e - 1, -1.0, e + 1, e + 1, 0.0)]             # floor
        if ceiling:
            boxes.append((-e - 1, -e - 1, h, e + 1, e + 1, h + 1.0))
        boxes += [(-e - 1, -e - 1, 0, -e, e + 1, h), (e, -e - 1, 0, e + 1, e + 1, h),
                  (-e - 1, -e - 1, 0, e + 1, -e, h), (-e - 1, e, 0, e + 1, e + 1, h)]
        for _ in range(n_boxes):
            cx, cy = rng.uniform(-e * 0.9, e * 0.9, 2)
            sx, sy = rng.uniform(0.3, max(0.6, e * 0.12), 2)
            sz = rng.uniform(0.5, h * 0.9)
            if abs(cx) < 1.5 and abs(cy) < 1.5:      # keep the start position free
                cx += 3.0
            boxes.append((cx - sx, cy - sy, 0.0, cx + sx, cy + sy, sz))
        self.lo = np.array([b[:3] for b in boxes], dtype=np.float64)
        self.hi = np.array([b[3:] for b in boxes], dtype=np.float64)

    def _cast_cuda(self, origin, dirs):
        """The same slab test with torch on the GPU (float64, identical IEEE operations): the 640x480 depth frames of cfg3
        against 400 boxes take seconds per frame in numpy.  Input generation only — never part of a timed region."""
        import torch
        dev = torch.device("cuda")
        d = torch.from_numpy(np.ascontiguousarray(dirs)).to(dev)
        o = torch.from_numpy(np.asarray(origin, np.float64)).to(dev)
        inv = 1.0 / d
        best = torch.full((d.shape[0],), float("inf"), dtype=torch.float64, device=dev)
        ninf, pinf = torch.tensor(float("-inf"), dtype=torch.float64, device=dev), torch.tensor(float("inf"), dtype=torch.float64, device=dev)
        for lo, hi in zip(self.lo, self.hi):
            t0 = (torch.from_numpy(lo).to(dev) - o) * inv
            t1 = (torch.from_numpy(hi).to(dev) - o) * inv
            lo_t, hi_t = torch.minimum(t0, t1), torch.maximum(t0, t1)      # NaN (0 * inf) propagates, then is ignored as in nanmax / nanmin
            tmin = torch.where(torch.isnan(lo_t), ninf, lo_t).max(dim=1).values
            tmax = torch.where(torch.isnan(hi_t), pinf, hi_t).min(dim=1).values
            hit = (tmax >= torch.clamp(tmin, min=0.0)) & (tmin > 1e-6)
            best = torch.where(hit & (tmin < best), tmin, best)
        return best.cpu().numpy()

    def cast(self, origin, dirs):
        """Distance along each unit direction to the first box, inf if none.  origin [3], dirs [n,3] (float64)."""
        n = dirs.shape[0]
        if n * len(self.lo) > 20_000_000:
            try:
                import torch
                if torch.cuda.is_available():
                    return self._cast_cuda(origin, dirs)
            except ImportError:
                pass
        best = np.full(n, np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / dirs
            for lo, hi in zip(self.lo, self.hi):
                t0 = (lo - origin) * inv
                t1 = (hi - origin) * inv
                tmin = np.nanmax(np.minimum(t0, t1), axis=1)
                tmax = np.nanmin(np.maximum(t0, t1), axis=1)
                hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 1e-6)
                best = np.where(hit & (tmin < best), tmin, best)
        return best


def trajectory(n_frames, start=(0.0, 0.0, 1.5), step=0.5, yaw_rate=0.02, pitch=0.03, roll=-0.02):
    out = []
    for k in range(n_frames):
        t = np.array([start[0] + step * k, start[1] + 0.1 * k, start[2]], dtype=np.float32)
        out.append((quat_from_euler(roll, pitch, yaw_rate * k), t))
    return out


def lidar3d_points(world, q, t, rings, az, elev_min_deg, elev_max_deg, max_range):
    """Sensor-frame float32 [n,3] point cloud; no-return rays dropped (cfg4/cfg5: OS-32 / OS-64 pattern)."""
    el = np.deg2rad(np.linspace(elev_min_deg, elev_max_deg, rings))
    azs = -np.pi + 2 * np.pi * np.arange(az) / az
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d_s = np.stack([ce * np.cos(azs)[None], ce * np.sin(azs)[None], np.broadcast_to(se, (rings, az))], axis=-1).reshape(-1, 3)
    R = quat_to_rot(q)
    rng = world.cast(np.asarray(t, np.float64), d_s @ R.T)
    ok = np.isfinite(rng) & (rng < max_range)
    return (d_s[ok] * rng[ok, None]).astype(np.float32)


def scan2d(world, q, t, scan_num, theta_min, theta_inc, max_range):
    th = theta_min + theta_inc * np.arange(scan_num)
    d_s = np.stack([np.cos(th), np.sin(th), np.zeros_like(th)], axis=-1)
    rng = world.cast(np.asarray(t, np.float64), d_s @ quat_to_rot(q).T)
    out = np.where(np.isfinite(rng) & (rng < max_range), rng, np.nan)
    return out.astype(np.float32)


def vlp16_ranges(world, q, t, scan_num, ring_num, theta_min, theta_inc, phi_min, phi_inc, max_range):
    """float32 [ring_num, scan_num] HORIZONTAL ranges, INFINITY = no return (Vlp16MapMaker layout)."""
    th = theta_min + theta_inc * np.arange(scan_num)
    ph = phi_min + phi_inc * np.arange(ring_num)
    cp, sp = np.cos(ph)[:, None], np.sin(ph)[:, None]
    d_s = np.stack([cp * np.cos(th)[None], cp * np.sin(th)[None], np.broadcast_to(sp, (ring_num, scan_num))], axis=-1).reshape(-1, 3)
    rng = world.cast(np.asarray(t, np.float64), d_s @ quat_to_rot(q).T).reshape(ring_num, scan_num)
    hor = rng * cp
    return np.where(np.isfinite(rng) & (rng < max_range), hor, np.inf).astype(np.float32)


def depth_image(world, q, t, rows, cols, cx, cy, fx, fy, near, far, nan_frac, rng_state):
    """float32 [rows, cols] depth (metres along the camera x axis; x forward, y left, z up), NaN outside [near, far]."""
    u, v = np.meshgrid(np.arange(cols), np.arange(rows))
    d_s = np.stack([np.ones(u.size), (cx - u.ravel()) / fx, (cy - v.ravel()) / fy], axis=-1)
    norm = np.linalg.norm(d_s, axis=1)
    d_u = d_s / norm[:, None]
    rng = world.cast(np.asarray(t, np.float64), d_u @ quat_to_rot(q).T)
    depth = rng / norm
    depth = np.where(np.isfinite(depth) & (depth >= near) & (depth <= far), depth, np.nan)
    if nan_frac > 0:
        depth = np.where(rng_state.rand(depth.size) < nan_frac, np.nan, depth)
    return depth.reshape(rows, cols).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# Workload configurations (BASELINE.json "configs").  `scale` shrinks the volume for CPU-sized tests while
# keeping the sensor model.
def make_config(name, local_size=None):
    if name == "cfg1":   # 2-D LiDAR, 128x128x32 @ 0.2 m, cutoff 2 m, fast_mode
        size = local_size or (128, 128, 32)
        return dict(name=name, sensor="scan2d", voxel_width=0.2, local_size=size, cutoff_grids_sq=100, fast_mode=True,
                    ogm_min_h=-10.0, ogm_max_h=10.0, occupancy_threshold=180, bucket_max=20000, block_max=40000,
                    scan_param=dict(scan_num=1081, max_r=30.0, theta_inc=float(np.float32(np.deg2rad(0.25))),
                                    theta_min=float(np.float32(np.deg2rad(-135.0)))),
                    world=dict(extent=40.0, height=3.0, n_boxes=160, ceiling=True), start=(-12.0, -3.0, 1.5))
    if name == "cfg2":   # VLP-16, 256^3 @ 0.2 m, cutoff 3 m, full waves
        size = local_size or (256, 256, 256)
        return dict(name=name, sensor="vlp16", voxel_width=0.2, local_size=size, cutoff_grids_sq=225, fast_mode=False,
                    ogm_min_h=-10.0, ogm_max_h=10.0, occupancy_threshold=180, bucket_max=100000, block_max=150000,
                    scan_param=dict(scan_num=440, ring_num=16, max_r=10.0, theta_inc=float(np.float32(2.0 * np.pi / 440)),
                                    theta_min=float(np.float32(-np.pi)), phi_inc=float(np.float32(np.deg2rad(2.0))),
                                    phi_min=float(np.float32(np.deg2rad(-15.0)))),
                    world=dict(extent=40.0, height=3.0, n_boxes=160, ceiling=True), start=(-12.0, -3.0, 1.5))
    if name == "cfg3":   # depth camera, 256^3 @ 0.1 m, cutoff 3 m
        size = local_size or (256, 256, 256)
        return dict(name=name, sensor="depth", voxel_width=0.1, local_size=size, cutoff_grids_sq=900, fast_mode=False,
                    ogm_min_h=-10.0, ogm_max_h=10.0, occupancy_threshold=180, bucket_max=100000, block_max=150000,
                    cam_param=dict(rows=480, cols=640, cx=320.5, cy=240.5, fx=554.26, fy=554.26, valid_NaN=True),
                    world=dict(extent=40.0, height=3.0, n_boxes=400, ceiling=True), start=(-12.0, -3.0, 1.5))
    if name == "cfg4":   # headline: OS-32 65 536 pts, 512^3 @ 0.1 m, cutoff 5 m, full waves
        size = local_size or (512, 512, 512)
        return dict(name=name, sensor="pointcloud", voxel_width=0.1, local_size=size, cutoff_grids_sq=2500, fast_mode=False,
                    ogm_min_h=-10.0, ogm_max_h=10.0, occupancy_threshold=180, bucket_max=400000, block_max=600000,
                    lidar=dict(rings=32, az=2048, elev_min=-22.5, elev_max=22.5, max_range=50.0),
                    world=dict(extent=40.0, height=6.0, n_boxes=200, ceiling=False), start=(-12.0, -3.0, 1.5))
    if name == "cfg5":   # multi-GPU: OS-64 131 072 pts, 1024 x 1024 x 1016 @ 0.1 m (Z <= 1022: 10-bit z of the coc codec), cutoff 5 m
        size = local_size or (1024, 1024, 1016)
        return dict(name=name, sensor="pointcloud", voxel_width=0.1, local_size=size, cutoff_grids_sq=2500, fast_mode=False,
                    ogm_min_h=-10.0, ogm_max_h=10.0, occupancy_threshold=180, bucket_max=800000, block_max=1200000,
                    lidar=dict(rings=64, az=2048, elev_min=-22.5, elev_max=22.5, max_range=60.0),
                    world=dict(extent=60.0, height=8.0, n_boxes=400, ceiling=False), start=(-12.0, -3.0, 1.5))
    raise KeyError(name)


def small_config(name, local_size, cutoff_grids_sq=None):
    cfg = make_config(name, tuple(local_size))
    ext = 0.5 * min(local_size[0], local_size[1]) * cfg["voxel_width"] * 0.8
    cfg["world"] = dict(cfg["world"], extent=max(ext, 2.0), n_boxes=10)
    cfg["bucket_max"], cfg["block_max"] = 4000, 12000
    cfg["start"] = (0.0, 0.0, 1.5)
    if cutoff_grids_sq is not None:
        cfg["cutoff_grids_sq"] = cutoff_grids_sq
    return cfg


def make_frames(cfg, n_frames, seed=42, start=None, dynamic=False):
    """List of frame dicts: q, t and the sensor payload for cfg['sensor'].  dynamic=True swaps the box set every
    second frame (obstacles appear and vanish), which exercises the raise-out / lower-out wavefronts."""
    w = cfg["world"]
    worlds = [World(w["extent"], w["height"], w["n_boxes"], seed=seed, ceiling=w["ceiling"])]
    if dynamic:
        worlds.append(World(w["extent"], w["height"], w["n_boxes"], seed=seed + 1000, ceiling=w["ceiling"]))
    rs = np.random.RandomState(seed + 1)
    step = min(0.5, 2.0 * cfg["voxel_width"] * max(1, cfg["local_size"][0] // 32))
    traj = trajectory(n_frames, start=start or cfg.get("start", (0.0, 0.0, 1.5)), step=step)
    frames = []
    for k, (q, t) in enumerate(traj):
        f = dict(q=q, t=t)
        world = worlds[(k // 2) % len(worlds)]
        s = cfg["sensor"]
        if s == "pointcloud":
            l = cfg["lidar"]
            f["points"] = lidar3d_points(world, q, t, l["rings"], l["az"], l["elev_min"], l["elev_max"], l["max_range"])
        elif s == "scan2d":
            sp = cfg["scan_param"]
            f["scan"] = scan2d(world, q, t, sp["scan_num"], sp["theta_min"], sp["theta_inc"], sp["max_r"])
        elif s == "vlp16":
            sp = cfg["scan_param"]
            f["ranges"] = vlp16_ranges(world, q, t, sp["scan_num"], sp["ring_num"], sp["theta_min"], sp["theta_inc"],
                                       sp["phi_min"], sp["phi_inc"], 100.0)
        elif s == "depth":
            cp = cfg["cam_param"]
            f["depth"] = depth_image(world, q, t, cp["rows"], cp["cols"], cp["cx"], cp["cy"], cp["fx"], cp["fy"], 0.3, 6.0,
                                     0.05, rs)
        frames.append(f)
    return frames


# ---------------------------------------------------------------------------------------------------------------------
# Raw sensor_msgs/PointCloud2 payloads (what the reference's MapMakers receive before their host-side conversion)
VLP16_POINT_DTYPE = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                              "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"], "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22})


def vlp16_pointcloud2(world, q, t, az=1800, rings=16, elev_min_deg=-15.0, elev_step_deg=2.0, max_range=100.0):
    """velodyne_pointcloud layout (point_step 22: x, y, z, intensity f32, ring u16, time f32), azimuth-major firing order,
    no-return rays dropped.  Returns (uint8 bytes, point_step, field offsets dict)."""
    el = np.deg2rad(elev_min_deg + elev_step_deg * np.arange(rings))
    azs = -np.pi + 2 * np.pi * (np.arange(az) + 0.37) / az
    ce, se = np.cos(el)[None, :], np.sin(el)[None, :]
    d_s = np.stack([ce * np.cos(azs)[:, None], ce * np.sin(azs)[:, None], np.broadcast_to(se, (az, rings))], axis=-1).reshape(-1, 3)
    ring = np.broadcast_to(np.arange(rings)[None, :], (az, rings)).reshape(-1)
    rng = world.cast(np.asarray(t, np.float64), d_s @ quat_to_rot(q).T)
    ok = np.isfinite(rng) & (rng < max_range)
    pts = np.zeros(int(ok.sum()), dtype=VLP16_POINT_DTYPE)
    p = (d_s[ok] * rng[ok, None]).astype(np.float32)
    pts["x"], pts["y"], pts["z"], pts["ring"] = p[:, 0], p[:, 1], p[:, 2], ring[ok]
    pts["intensity"] = 1.0
    return pts.view(np.uint8).reshape(-1).copy(), 22, dict(x=0, y=4, z=8, ring=16)
