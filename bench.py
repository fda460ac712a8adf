#!/usr/bin/env python
"""bench.py — frames/s of the per-frame OGM + hash merge + batch EDT + wavefront merge path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4] [--impl ours|reference]

A "step" is one frame of VOLMAPNODE::publishMap's hot path (reference src/volumetric_mapper.cpp:138-224) on a moving
synthetic sensor (gie-mapping_b200/scenes.py).  N=1 runs the headline configuration cfg4 (512^3 @ 0.1 m, 65 536-point
OS-32 scan, full wavefronts).  N>1 runs ONE volume of the multi-GPU configuration (cfg5: 1024 x 1024 x 1016 @ 0.1 m,
131 072-ray scan) sharded over the N GPUs (gie-mapping_b200/sharded.py): strong scaling; rank 0 also times the
same frames on its GPU alone, so the line carries its own single-GPU figure.

  value      frames/s with the sensor frames already resident in HBM (CUDA events on the launch stream)
  e2e        frames/s through the host-buffer C ABI: pinned host points -> H2D inside the timed region, and a D2H read of
             the frame's result record (device status + wavefront statistics)
  roofline   bytes the batch-DT sweeps have to move for the frames at hand / their CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the C oracle (oracle/gie_oracle.c, a port) at the full volume: one thread, and all host cores for the batch EDT

--impl reference times the reference's own CUDA sources recompiled for sm_100a (oracle/_ref/ref_driver_fast: original
Release flags, cuTT replaced by a gather shim) on the same frames; the reference has no CPU implementation of this path.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_pkg():
    if "gie_mapping_b200" in sys.modules:
        return sys.modules["gie_mapping_b200"]
    pkg_dir = os.path.join(ROOT, "gie-mapping_b200")
    spec = importlib.util.spec_from_file_location("gie_mapping_b200", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gie_mapping_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


# Algorithmic bytes of the batch DT.  Two figures are reported and labelled:
#   * `achieved` uses the bytes the three sweeps of THIS engine have to move for the frame at hand — a function of s, the
#     fraction of z-slices that hold an obstacle (slices without one are never written by the x sweep nor read by the z sweep):
#       y pass : 0.25 R (bit words scanned for links; the bits themselves are set from the merge's block list, a few MB)
#       x sweep: 0.25 s R (bit words) + 8 s W (g2, cxy)
#       z sweep: 8 s R + 8 W (aux, coc_aux)                                     => 8.25 + 16.25 s  bytes per voxel
#     (ncu's dram__bytes for the same launches is `traffic`; the two agree within a few percent, profiles/README.md);
#   * `survey_41B_figure` is SURVEY §8d's packed-intermediate 41 B/voxel (P1 1R+8W, P2 8R+8W, P3 8R+8W), the traffic of a
#     design that materialises every pass for every voxel; it exceeds what is moved here and is NOT used for `frac`.
SURVEY_BATCH_DT_BYTES = 41.0
BATCH_DT_STAGES = ("edt_pack", "edt_x", "edt_z")
OTHER_STAGES = ("ogm", "hash_merge", "mark_frontier", "waves", "commit")


def batch_dt_bytes_per_voxel(s):
    return 8.25 + 16.25 * s


def workload_string(cfg, frames):
    """One description of the workload, identical in both arms (the driver compares the strings)."""
    X, Y, Z = cfg["local_size"]
    key = {"pointcloud": "points", "scan2d": "scan", "vlp16": "ranges", "depth": "depth"}[cfg["sensor"]]
    n = int(np.mean([np.asarray(f[key]).size for f in frames]))
    per = f"{n // 3} points/frame" if cfg["sensor"] == "pointcloud" else f"{n} range values/frame"
    return (f"{cfg['name']}: {X}x{Y}x{Z} @ {cfg['voxel_width']} m, {cfg['sensor']} {per}, cutoff_grids_sq={cfg['cutoff_grids_sq']}, "
            f"fast_mode={cfg['fast_mode']}")


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"nproc": os.cpu_count(), "cpu_model": model}


class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons DURING the timed region, in-process through NVML (a polling `nvidia-smi`
    subprocess stalls host<->device synchronisation for tens of ms per query, which lands in the e2e number)."""

    def __init__(self, index, period_s=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.sm, self.reasons = [], set()
        self.sm_max = None
        self.stop_flag = False
        self.active = False     # samples are kept only while a timed region is open

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                if self.active:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:   # no NVML: report nothing rather than a made-up clock
            self.error = repr(e)

    def stop(self):
        self.stop_flag = True

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": "NVML, 20 ms period, timed regions only"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def cpu_baseline(gie, cfg, frames):
    """The C oracle (a port: the reference has no CPU implementation of this path) at the FULL size of the workload, on a
    bounded sample of its frames: single-threaded, and with the batch EDT's column loops spread over all host cores
    (pthreads; ray cast, hash merge and wavefronts stay serial).  `value` is the all-core figure."""
    from oracle import oracle_py
    info = host_info()
    out = {"unit": "frames/s", "kind": "port", **info}
    runs = {}
    big = cfg["local_size"][0] * cfg["local_size"][1] * cfg["local_size"][2] > 300e6   # cfg5: ~75 s per frame on one core
    for label, threads, n in (("single_thread", 1, 1 if big else 2), ("all_cores", info["nproc"] or 1, 2 if big else 3)):
        oracle_py.set_threads(threads)
        om = oracle_py.OracleMapper(cfg)
        n = min(n, len(frames))
        t0 = time.perf_counter()
        for f in frames[:n]:
            om.publishMap(f)
        dt = time.perf_counter() - t0
        om.close()
        runs[label] = {"value": n / dt, "frames": n, "seconds": dt, "threads": threads}
    oracle_py.set_threads(1)
    X, Y, Z = cfg["local_size"]
    out.update({"value": runs["all_cores"]["value"], "cores": runs["all_cores"]["threads"],
                "single_thread": runs["single_thread"], "all_cores": runs["all_cores"],
                "mvoxels_per_s": runs["all_cores"]["value"] * X * Y * Z / 1e6,
                "sample": f"the first {runs['single_thread']['frames']} (1 thread, {runs['single_thread']['seconds']:.1f} s) and "
                          f"{runs['all_cores']['frames']} ({runs['all_cores']['threads']} threads, {runs['all_cores']['seconds']:.1f} s) frames of "
                          f"the {cfg['name']} stream at the full {X}x{Y}x{Z} volume"})
    return out


def batch_dt_dense_case(gie, cfg, stream):
    """The batch DT alone on a volume with obstacles in EVERY slice and most columns (random 0.2 % occupancy): nothing for
    the sweeps to skip, so this is the bandwidth/latency-bound regime of the three kernels."""
    import torch
    X, Y, Z = cfg["local_size"]
    rng = np.random.RandomState(5)
    t = np.where(rng.rand(Z, Y, X) < 0.002, 2, 1).astype(np.int8)
    lm = gie.LocMap(cfg["voxel_width"], (X, Y, Z), cutoff_grids_sq=cfg["cutoff_grids_sq"])
    try:
        lm.set_stream(stream.cuda_stream)
        lm.upload_glb_type(t)
        for _ in range(3):
            lm.batchEDTUpdate()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            lm.batchEDTUpdate()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    finally:
        lm.close()
    nvox = X * Y * Z
    peak, _ = measured_peak()
    bpv = batch_dt_bytes_per_voxel(1.0)
    ach = bpv * nvox / (ms * 1e-3) / 1e9
    return {"workload": f"{X}x{Y}x{Z}, random 0.2 % occupancy in every slice (s = 1: nothing to skip)", "ms": ms,
            "bytes_per_voxel": bpv, "achieved": ach, "frac": ach / peak,
            "survey_41B_figure": {"achieved": SURVEY_BATCH_DT_BYTES * nvox / (ms * 1e-3) / 1e9}}


def cpp_host_leg(gie, cfg, frames, warmup):
    """The same frames through the C++ host surface (include/gie_compat: LocMap / GlbHashMap / localOGMKernels /
    batchEDTUpdate / mergeNewObsv, driven by gie-mapping_b200/gie_replay) with pageable host buffers and blocking copies, timed
    per frame on the host with a device synchronise at both ends — the way the reference times itself
    (volumetric_mapper.cpp:152-203) and the way `--impl reference` is timed."""
    import tempfile
    try:
        with tempfile.TemporaryDirectory() as d:
            _, _, out = gie.replay_io.run_replay(cfg, frames, d, timing=True)
        t = []
        for line in out.splitlines():
            if line.startswith("frame "):
                p = line.split()
                t.append(float(p[3]) + float(p[5]))
        t = t[warmup:]
        ms = float(np.mean(t))
        return {"value": 1000.0 / ms, "unit": "frames/s", "ms_per_step": ms, "frames": len(t),
                "how": "gie_replay --time: C++ host, pageable buffers, cudaMemcpy + cudaDeviceSynchronize per half frame"}
    except Exception as e:
        return {"error": repr(e)[:300]}


def run_sharded(args, gie, world, rank, local_rank):
    """N > 1: ONE local volume of BASELINE configs[4] (1024 x 1024 x 1016 @ 0.1 m, 64 x 2048-ray scan) sharded over the N GPUs
    (gie-mapping_b200/sharded.py): strong scaling.  Rank 0 also times the same frames on its own GPU alone afterwards, so the
    line carries its own single-GPU figure for the same workload."""
    import torch
    import torch.distributed as dist
    from gie_mapping_b200 import sharded
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    cfg = gie.scenes.make_config(args.config if args.config != "cfg4" else "cfg5")
    X, Y, Z = cfg["local_size"]
    nvox = X * Y * Z
    nframes = args.warmup + args.steps
    frames = gie.scenes.make_frames(cfg, nframes, seed=42) if rank == 0 else [None] * nframes
    host_in = [torch.from_numpy(np.ascontiguousarray(f["points"], np.float32)).pin_memory() for f in frames] if rank == 0 else []
    dev_in = [h.to(dev) for h in host_in]

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(make, device_resident):
        mp = make()
        for k in range(args.warmup):
            mp.publishMap(frames[k], device_input=dev_in[k].data_ptr() if (rank == 0 and device_resident) else None)
        if rank == 0:
            mp.hash_map.sync()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active = True
        e0.record(stream)
        for k in range(args.warmup, nframes):
            if rank == 0 and not device_resident:
                f = dict(frames[k])
                f["points"] = host_in[k].numpy()
                mp.publishMap(f)
                mp.hash_map.sync()
                mp.hash_map.wave_stats()
            else:
                mp.publishMap(frames[k], device_input=dev_in[k].data_ptr() if rank == 0 else None)
        e1.record(stream)
        barrier()
        sampler.active = False
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), mp

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    make_sharded = lambda: sharded.ShardedMapper(cfg, rank=rank, world=world)
    ms_dev, mp = timed(make_sharded, True)
    launches = mp.owner.loc_map.launch_count() if rank == 0 else 0
    wave_stats = mp.hash_map.wave_stats() if rank == 0 else None
    nblocks = mp.hash_map.num_blocks() if rank == 0 else None
    bytes_rx = mp.bytes_received
    mp.close()
    ms_e2e, mp = timed(make_sharded, False)
    # stage profile of the sharded frame: the owner's stages on rank 0, the slab sweeps on every rank (max over ranks)
    if rank == 0:
        mp.owner.loc_map.profile_enable(True)
    _check_slab = mp.slab
    gie.load_library().gie_profile_enable(_check_slab._h, 1)
    prof, sweeps = {}, 0.0
    import ctypes as C
    reps = min(10, args.steps)
    for k in range(nframes - reps, nframes):
        mp.publishMap(frames[k], device_input=dev_in[k].data_ptr() if rank == 0 else None)
        ms = np.zeros(len(gie.STAGE_NAMES), np.float32)
        gie.load_library().gie_profile_last(_check_slab._h, ms.ctypes.data_as(C.c_void_p))
        sweeps += float(ms[gie.STAGE_NAMES.index("edt_x")] + ms[gie.STAGE_NAMES.index("edt_z")]) / reps
        if rank == 0:
            for name, v in mp.owner.loc_map.profile_last().items():
                prof[name] = prof.get(name, 0.0) + v / reps
    t = torch.tensor([sweeps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sweeps_max = float(t.item())
    rx = torch.tensor([float(bytes_rx)], device=dev)
    dist.all_reduce(rx, op=dist.ReduceOp.MAX)
    mp.close()
    sampler.stop()
    barrier()
    # the same frames on rank 0's GPU alone (the other ranks wait)
    single = None
    if rank == 0 and not args.no_single_gpu_leg:
        try:
            m1 = gie.Mapper(cfg)
            m1.loc_map.set_stream(stream.cuda_stream)
            for k in range(args.warmup):
                m1.publishMap(frames[k], device_input=dev_in[k].data_ptr())
            m1.hash_map.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(args.warmup, nframes):
                m1.publishMap(frames[k], device_input=dev_in[k].data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            single = {"ms_per_step": e0.elapsed_time(e1) / args.steps}
            single["value"] = 1000.0 / single["ms_per_step"]
            m1.loc_map.profile_enable(True)
            sp = {}
            for k in range(nframes - reps, nframes):
                m1.publishMap(frames[k], device_input=dev_in[k].data_ptr())
                for name, v in m1.loc_map.profile_last().items():
                    sp[name] = sp.get(name, 0.0) + v / reps
            single["stage_ms"] = sp
            m1.close()
        except Exception as e:
            single = {"error": repr(e)[:300]}
    barrier()
    if rank == 0:
        fps, fps_e2e = 1000.0 / ms_dev, 1000.0 / ms_e2e
        peak, peak_src = measured_peak()
        line = {"metric": "EDT+OGM frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
                "config": {"workload": workload_string(cfg, frames),
                           "parallelism": f"ONE local volume sharded over {world} GPUs: {world} slabs of {Y // world} rows y; dense half (x and z sweeps of the "
                                          "batch EDT) on every rank, y pass broadcast from rank 0 (NCCL, obstacle-bearing slices only), sparse half "
                                          "(ray cast, hash merge, wavefronts) on rank 0 reading the slabs through CUDA IPC over NVLink",
                           "l2": "per-frame working set (>= 8.5 GB of batch-EDT output) exceeds the 126 MB L2; no explicit flush"},
                "mvoxels_per_s": fps * nvox / 1e6,
                "e2e": {"value": fps_e2e, "unit": "frames/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(np.mean([h.numel() * 4 for h in host_in])) + 28, "d2h_bytes_per_step": 4 + 64},
                "gpu_launches": int(launches), "stage_ms_rank0": prof, "slab_sweeps_ms_max_over_ranks": sweeps_max,
                "broadcast_bytes_per_frame": float(rx.item()) / max(1, nframes), "wave_stats": wave_stats, "blocks": nblocks,
                "single_gpu_same_workload": single,
                "strong_scaling": ({"speedup": single["ms_per_step"] / ms_dev, "efficiency": single["ms_per_step"] / ms_dev / world}
                                   if single and "ms_per_step" in single else None),
                "roofline": {"bound": "hbm", "kernel": "slab sweeps = k_edt_xsweep + k_edt_zsweep on the slab of the slowest rank",
                             "achieved": (batch_dt_bytes_per_voxel(0.0) - 0.25) * nvox / world / (sweeps_max * 1e-3) / 1e9 if sweeps_max > 0 else 0.0,
                             "peak": peak, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, burst copy)", "unit": "GB/s",
                             "frac": ((batch_dt_bytes_per_voxel(0.0) - 0.25) * nvox / world / (sweeps_max * 1e-3) / 1e9 / peak) if sweeps_max > 0 else 0.0,
                             "traffic": None,
                             "note": "lower bound of the bytes of one slab: 8 B/voxel of aux + coc_aux written (the slice-dependent x-sweep output and "
                                     "z-sweep input are left out); per-GPU figure"},
                "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def run_reference(args, gie, cfg, frames):
    from oracle import ref_io
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    line = {"impl": "reference", "metric": "EDT+OGM frames/sec", "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "int32+f32", "data": "synthetic", "config": {"workload": workload_string(cfg, frames)}}
    if args.gpus > 1:
        line["config"]["parallelism"] = "none: the reference is a single-GPU program; it ran on one GPU of the box"
    X, Y, Z = cfg["local_size"]
    nvox = X * Y * Z
    t = None
    if ref_io.available("fast"):
        try:
            t = ref_io.run(cfg, frames, "fast", timing=True)
        except Exception as e:      # e.g. no GPU on this host: fall back to the CPU port below
            line["reference_driver_error"] = repr(e)[:200]
    if t:
        t = t[args.warmup:]
        per_frame = np.array([a + b for a, b in t])
        # the reference's OGM half allocates and sorts 134 M keys per frame: its frame time has a long tail and moves between
        # boxes, so the MEDIAN frame is reported (mean, quartiles and both halves are kept beside it)
        ms = float(np.median(per_frame))
        fps = 1000.0 / ms
        pts_bytes = int(np.mean([f[ref_io.PAYLOAD_KEY[cfg["sensor"]]].nbytes for f in frames]))
        q = lambda v, p: float(np.percentile(v, p))
        ogm, edt = np.array([a for a, _ in t]), np.array([b for _, b in t])
        line.update({"value": fps, "ms_per_step": ms, "mvoxels_per_s": fps * nvox / 1e6, "statistic": f"median of {len(t)} frames",
                     "spread_ms": {"mean": float(per_frame.mean()), "p25": q(per_frame, 25), "p75": q(per_frame, 75), "min": float(per_frame.min()),
                                   "max": float(per_frame.max())},
                     "stage_ms": {"ogm_half": {"median": float(np.median(ogm)), "p25": q(ogm, 25), "p75": q(ogm, 75)},
                                  "edt_half": {"median": float(np.median(edt)), "p25": q(edt, 25), "p75": q(edt, 75)}},
                     "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "reference", **host_info(),
                                      "sample": "the reference's own CUDA sources (unmodified, Release flags, cuTT replaced by a gather "
                                                "shim) recompiled for sm_100a and run on the GPU; it has no CPU implementation; "
                                                "host-timed per frame with cudaDeviceSynchronize as the reference does"},
                     "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": pts_bytes, "d2h_bytes_per_step": 0}})
    else:
        cb = cpu_baseline(gie, cfg, frames)
        line.update({"value": cb["value"], "ms_per_step": 1000.0 / cb["value"],
                     "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense-case", action="store_true")
    ap.add_argument("--no-single-gpu-leg", action="store_true", help="N > 1: skip timing the same frames on rank 0's GPU alone")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    gie = load_pkg()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        if args.impl == "reference":
            # The reference is a single-GPU program: rank 0 alone runs it, on the multi-GPU arm's workload (the ONE cfg5 volume
            # that arm shards), the other ranks leave at once.  Its own size check ("Local map size too big!!!",
            # local_batch.h:54-58) already fires for the 512^3 headline volume and is compiled out in Release builds; it runs on.
            if rank != 0:
                return
            cfg = gie.scenes.make_config(args.config if args.config != "cfg4" else "cfg5")
            frames = gie.scenes.make_frames(cfg, args.warmup + args.steps, seed=42)
            run_reference(args, gie, cfg, frames)
            return
        run_sharded(args, gie, world, rank, local_rank)
        return
    cfg = gie.scenes.make_config(args.config)
    nframes = args.warmup + args.steps

    frames = gie.scenes.make_frames(cfg, nframes, seed=42 + rank)
    if args.impl == "reference":
        run_reference(args, gie, cfg, frames)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()

    key = {"pointcloud": "points", "scan2d": "scan", "vlp16": "ranges", "depth": "depth"}[cfg["sensor"]]
    host_in = [torch.from_numpy(np.ascontiguousarray(f[key], np.float32)).pin_memory() for f in frames]
    dev_in = [h.to(dev) for h in host_in]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_pass(device_resident):
        """Fresh map, W warm-up frames, K timed frames.  Returns ms/step (max over ranks) and the mapper."""
        mp = gie.Mapper(cfg)
        mp.loc_map.set_stream(stream.cuda_stream)
        result = torch.zeros(9, dtype=torch.int64).pin_memory()
        for k in range(args.warmup):
            mp.publishMap(frames[k], device_input=dev_in[k].data_ptr() if device_resident else None)
        mp.hash_map.sync()
        barrier()
        l0 = mp.loc_map.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active = True
        e0.record(stream)
        for k in range(args.warmup, nframes):
            if device_resident:
                mp.publishMap(frames[k], device_input=dev_in[k].data_ptr())
            else:
                f = dict(frames[k])
                f[key] = host_in[k].numpy()
                mp.publishMap(f)               # H2D copy of the frame happens inside the C ABI call
                mp.hash_map.sync()             # D2H of the device status word
                st = mp.hash_map.wave_stats()  # D2H result record (pinned, written by the device)
                result[0] = st["fC"]
        e1.record(stream)
        barrier()
        sampler.active = False
        ms = e0.elapsed_time(e1) / args.steps
        launches = mp.loc_map.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, mp

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ms_dev, launches, mp = run_pass(True)
    wave_stats = mp.hash_map.wave_stats()
    nblocks = mp.hash_map.num_blocks()
    mp.close()
    ms_e2e, _, mp2 = run_pass(False)
    mp2.close()
    sampler.stop()
    clocks = sampler.summary()
    # Stage profile: a THIRD pass over the very same frames with CUDA events around every stage (the events themselves cost a
    # little, so this pass is not the timed one).  frame - sum(stages) is then the time spent outside kernels on these frames.
    mp3 = gie.Mapper(cfg)
    mp3.loc_map.set_stream(stream.cuda_stream)
    for k in range(args.warmup):
        mp3.publishMap(frames[k], device_input=dev_in[k].data_ptr())
    mp3.loc_map.profile_enable(True)
    prof, slice_frac = {}, 0.0
    for k in range(args.warmup, nframes):
        mp3.publishMap(frames[k], device_input=dev_in[k].data_ptr())
        for name, v in mp3.loc_map.profile_last().items():
            prof[name] = prof.get(name, 0.0) + v / args.steps
        slice_frac += float((mp3.loc_map.edt_slice_columns() > 0).mean()) / args.steps
    mp3.close()

    X, Y, Z = cfg["local_size"]
    nvox = X * Y * Z
    fps = world * 1000.0 / ms_dev
    fps_e2e = world * 1000.0 / ms_e2e
    peak, peak_src = measured_peak()
    stage_sum = sum(prof.values())
    prof["batch_dt"] = sum(prof.get(k, 0.0) for k in BATCH_DT_STAGES)
    dense = batch_dt_dense_case(gie, cfg, stream) if (rank == 0 and not args.no_dense_case) else None

    # The roofline object is reported for the batch-DT sweep group (EDT_OCC::batchEDTUpdate), the kernel north_star grades.
    bpv = batch_dt_bytes_per_voxel(slice_frac)
    achieved = bpv * nvox / (prof["batch_dt"] * 1e-3) / 1e9 if prof["batch_dt"] > 0 else 0.0
    traffic = {}
    for tf in ("traffic_r02.json", "traffic_r01.json"):
        tp = os.path.join(ROOT, "profiles", tf)
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp))
                traffic["_file"] = tf
                break
            except Exception:
                traffic = {}

    def per_kernel(k):
        d = {"ms": prof.get(k, 0.0), "ncu_dram_bytes": traffic.get(k)}
        if traffic.get(k) and prof.get(k, 0.0) > 0:
            d["dram_gbs"] = traffic[k] / (prof[k] * 1e-3) / 1e9      # measured bytes of one ncu capture / live event time
        return d

    roofline = {"bound": "hbm", "kernel": "batch_dt = k_edt_ycols + k_edt_slices + k_edt_xsweep + k_edt_zsweep (EDT_OCC::batchEDTUpdate; the y-pass bits are set by the OGM merge)",
                "achieved": achieved, "peak": peak, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, burst copy)", "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic.get("batch_dt"), "traffic_source": traffic.get("_file"),
                "algorithmic_bytes_per_voxel": bpv, "algorithmic_bytes_per_launch": bpv * nvox,
                "bytes_formula": "8.25 + 16.25 s B/voxel, s = fraction of z-slices holding an obstacle (bytes this engine's sweeps must move)",
                "obstacle_slice_fraction": slice_frac,
                "survey_41B_figure": {"bytes_per_voxel": SURVEY_BATCH_DT_BYTES, "achieved": SURVEY_BATCH_DT_BYTES * nvox / (prof["batch_dt"] * 1e-3) / 1e9
                                      if prof["batch_dt"] > 0 else None,
                                      "note": "traffic of a design that materialises all three passes for every voxel; larger than what is moved here, not used for frac"},
                "kernel_ms": prof["batch_dt"], "kernel_ms_parts": {k: prof.get(k, 0.0) for k in BATCH_DT_STAGES},
                "note": f"live CUDA-event time of the sweep kernels on the engine stream, mean over the {args.steps} timed frames (profiled pass)",
                "dense_case": dense,
                "longest_stage": max((k for k in prof if k not in BATCH_DT_STAGES and k != "batch_dt"), key=lambda k: prof[k]),
                "per_kernel": {k: per_kernel(k) for k in BATCH_DT_STAGES + OTHER_STAGES}}
    line = {"metric": "EDT+OGM frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+f32", "data": "synthetic",
            "config": {"workload": workload_string(cfg, frames),
                       "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas of the per-frame path (one map per GPU, no "
                                                                  "data-path collective); the sharded batch EDT is reported separately",
                       "l2": "per-frame working set (>= 1.5 GB) exceeds the 126 MB L2; no explicit flush"},
            "mvoxels_per_s": fps * nvox / 1e6,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(np.mean([h.numel() * 4 for h in host_in])) + 28, "d2h_bytes_per_step": 4 + 64},
            "gpu_launches": int(launches),
            "stage_ms": prof, "outside_kernels_ms": ms_dev - stage_sum,
            "wave_stats": wave_stats, "blocks": nblocks,
            "roofline": roofline, "clocks": clocks}
    if rank == 0 and world == 1:
        line["e2e_cpp_host"] = cpp_host_leg(gie, cfg, frames, args.warmup)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(gie, cfg, frames)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
