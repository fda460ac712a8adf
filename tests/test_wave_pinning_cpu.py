"""Pins the wavefront stage (MarkLimitedObserve -> obtainFrontiers -> waves A/B/C -> UpdateHashBatch) against the reference
ITSELF and against ground truth.

Fixtures: tests/golden/wavevar/*.npz, produced on a B200 by oracle/wave_variance.py from the reference's own CUDA sources
(oracle/_ref/ref_driver_parity): run 0's occupancy and committed (dist, coc id) per frame, the OCCUPIED voxels of the global
map in a halo box wider than the cut-off distance, and the mismatch counts of 5 further runs of the reference against run 0.

What the fixtures established:
  * the reference's distances are reproducible run to run on a B200 (0 differing distances in 6 runs of every frame; at most
    2 coc ids differ, among equidistant obstacles) — differences against it are therefore ours to explain, not noise;
  * all but a handful of the round-1 differences (up to 524 of 21 271 voxels) came from ONE rule: the reference leaves the
    (dist, id) pair of UNKNOWN voxels stale (unify_helper.cuh:217-218) and its wave C relaxes against those stale words.
    With that restated exactly, what remains is the tie rule (first arrival there, smaller coc id here).
The arbiter is the reference's own definition of correctness, Gnd_truth_checker::cmp_dist (include/gt_checker.h:30-80): the
distance to the nearest OCCUPIED voxel of the global map, here from an exact KD-tree query."""
import glob
import json
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

WAVEVAR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wavevar")


def _cases():
    return sorted(glob.glob(os.path.join(WAVEVAR, "*.npz")))


def compare_with_reference(g, k, dist, glb_type):
    """Accounting of one frame: `dist` int64 [Z,Y,X] committed squared distances of the implementation under test."""
    H = int(g["halo"])
    rt, rd = g[f"f{k}_glb_type"], g[f"f{k}_pair_dist"].astype(np.int64)
    known = rt != 0
    out = dict(known=int(known.sum()), type_mismatch=int((glb_type != rt).sum()), dist_mismatch=0, sse_ours=0.0, sse_ref=0.0,
               worst_excess=0.0, max_abs_delta=0.0)
    dm = known & (dist != rd)
    out["dist_mismatch"] = int(dm.sum())
    if out["dist_mismatch"]:
        occ = np.argwhere(g[f"f{k}_box_type"] == 2)
        tree = cKDTree(occ[:, ::-1].astype(np.float64))
        zz, yy, xx = np.nonzero(dm)
        truth, _ = tree.query(np.stack([xx + H, yy + H, zz + H], 1).astype(np.float64))
        o, r = np.sqrt(dist[dm].astype(np.float64)), np.sqrt(rd[dm].astype(np.float64))
        out["sse_ours"], out["sse_ref"] = float(((o - truth) ** 2).sum()), float(((r - truth) ** 2).sum())
        out["worst_excess"] = float((np.abs(o - truth) - np.abs(r - truth)).max())
        out["max_abs_delta"] = float(np.abs(o - r).max())
    return out


def check_accounting(tag, acc, waves_ran):
    # a frontier (FNT) mark depends on whether the voxel's closest obstacle lies inside the volume, i.e. on earlier ties
    assert acc["type_mismatch"] <= (2 if waves_ran else 0), f"{tag}: occupancy / frontier marks differ in {acc['type_mismatch']} voxels"
    if not waves_ran:
        assert acc["dist_mismatch"] == 0, f"{tag}: {acc['dist_mismatch']} distances differ before any wavefront ran"
        return
    assert acc["dist_mismatch"] <= max(2, acc["known"] // 1000), f"{tag}: {acc['dist_mismatch']} of {acc['known']} distances differ"
    # where we differ we must not be farther from ground truth than the reference, as a whole (squared error) ...
    assert acc["sse_ours"] <= acc["sse_ref"] + 1e-9, f"{tag}: squared error vs ground truth {acc['sse_ours']:.4f} > reference's {acc['sse_ref']:.4f}"
    # ... and per voxel by no more than one propagation step at distance >= 8 (|sqrt(d+1) - sqrt(d)| < 0.0625 voxel)
    assert acc["worst_excess"] <= 0.0625 + 1e-9, f"{tag}: a voxel is {acc['worst_excess']:.4f} voxel farther from truth than the reference"
    assert acc["max_abs_delta"] <= 0.5, f"{tag}: |ours - reference| = {acc['max_abs_delta']:.3f} voxel"


def test_reference_is_reproducible():
    """The measured self-variance of the reference (6 runs per frame on a B200): distances never differ."""
    rep = json.load(open(os.path.join(WAVEVAR, "self_variance.json")))
    assert set(rep) >= {"pc_static", "pc_dynamic", "vlp16", "pc_dynamic_96"}
    for name, frames in rep.items():
        for f in frames:
            assert max(f["self_dist_mismatch"]) == 0, (name, f)
            assert max(f["self_type_mismatch"]) == 0, (name, f)
            assert max(f["self_id_mismatch"]) <= 2, (name, f)


@pytest.mark.parametrize("path", _cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_pinned_against_reference(gie, oracle, path):
    g = np.load(path)
    cfg = gie.scenes.small_config(str(g["cfg_name"]), tuple(int(v) for v in g["size"]), cutoff_grids_sq=int(g["cutoff"]))
    frames = gie.scenes.make_frames(cfg, int(g["nframes"]), dynamic=bool(g["dynamic"]))
    om = oracle.OracleMapper(cfg)
    waves_ran, exact, total_mismatch, total_known = False, 0, 0, 0
    try:
        for k, f in enumerate(frames):
            om.publishMap(f)
            st = om.stats()
            waves_ran = waves_ran or (st["fA"] + st["fB"] + st["fC"]) > 0
            dist = (om.pair >> np.uint64(32)).astype(np.int64)
            acc = compare_with_reference(g, k, dist, om.glb_type)
            check_accounting(f"{os.path.basename(path)} frame {k}", acc, waves_ran)
            exact += acc["dist_mismatch"] == 0
            total_mismatch += acc["dist_mismatch"]
            total_known += acc["known"]
    finally:
        om.close()
    assert waves_ran or "scan" in path
    assert exact >= 3
    assert total_mismatch <= total_known * 5e-4
