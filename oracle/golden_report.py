"""Regenerates tests/golden/report.json and tests/golden/wavevar/accounting.json on the CPU: the oracle against the committed
fixtures of the reference's own CUDA sources (tests/golden/*.npz from oracle/gen_golden.py, tests/golden/wavevar/*.npz from
oracle/wave_variance.py).  Per frame: differing occupancy / distance / coc-id counts, and for the wavevar cases the ground-truth
accounting of tests/test_wave_pinning_cpu.py.  Test infrastructure only.  Run:  python oracle/golden_report.py"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
from oracle import oracle_py  # noqa: E402
from test_wave_pinning_cpu import compare_with_reference  # noqa: E402


def run_case(gie, path, wavevar):
    g = np.load(path)
    cfg = gie.scenes.small_config(str(g["cfg_name"]), tuple(int(v) for v in g["size"]), cutoff_grids_sq=int(g["cutoff"]))
    frames = gie.scenes.make_frames(cfg, int(g["nframes"]), dynamic=bool(g["dynamic"]))
    om = oracle_py.OracleMapper(cfg)
    rep = []
    for k, f in enumerate(frames):
        om.publishMap(f)
        rt = g[f"f{k}_glb_type"]
        known = rt != 0
        od = (om.pair >> np.uint64(32)).astype(np.int64)
        oid = (om.pair & np.uint64(0xffffffff)).astype(np.int64)
        rd = g[f"f{k}_pair_dist"].astype(np.int64)
        rid = g[f"f{k}_pair_id"].astype(np.int64) & 0xffffffff
        row = dict(frame=k, known=int(known.sum()), glb_type_mismatch=int((om.glb_type != rt).sum()),
                   pair_dist_mismatch=int((od[known] != rd[known]).sum()), pair_id_mismatch=int((oid[known] != rid[known]).sum()), stats=om.stats())
        if wavevar:
            acc = compare_with_reference(g, k, od, om.glb_type)
            row.update(sse_vs_truth_ours=acc["sse_ours"], sse_vs_truth_reference=acc["sse_ref"], worst_excess_voxel=acc["worst_excess"],
                       max_abs_delta_voxel=acc["max_abs_delta"],
                       reference_self_dist_mismatch=[int(v) for v in g[f"f{k}_self_dist_mismatch"]],
                       reference_self_id_mismatch=[int(v) for v in g[f"f{k}_self_id_mismatch"]])
        rep.append(row)
    om.close()
    return rep


def main():
    gie = load_pkg()
    oracle_py.build()
    gold = os.path.join(ROOT, "tests", "golden")
    report = {os.path.basename(p)[:-4]: run_case(gie, p, False) for p in sorted(glob.glob(os.path.join(gold, "*.npz")))}
    json.dump(report, open(os.path.join(gold, "report.json"), "w"), indent=1)
    acc = {os.path.basename(p)[:-4]: run_case(gie, p, True) for p in sorted(glob.glob(os.path.join(gold, "wavevar", "*.npz")))}
    json.dump(acc, open(os.path.join(gold, "wavevar", "accounting.json"), "w"), indent=1)
    for name, rep in {**report, **{"wavevar/" + k: v for k, v in acc.items()}}.items():
        tot = sum(r["pair_dist_mismatch"] for r in rep)
        print(f"{name:24s} frames {len(rep):2d}  known voxels {sum(r['known'] for r in rep):8d}  differing distances {tot:4d}  types {sum(r['glb_type_mismatch'] for r in rep)}")


if __name__ == "__main__":
    main()
