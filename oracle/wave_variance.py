"""Measures how much the reference's wavefront stage differs FROM ITSELF run to run, and collects what an arbiter needs.

Runs oracle/_ref/ref_driver_parity (the reference's own CUDA sources, IEEE flags; see build_ref.sh) K times on each seeded
case and stores, per frame and run, the reference's glb_type / committed (dist, coc id) pair, plus the hash voxels of a
halo box wide enough to hold every obstacle within the cut-off distance of the volume (run 0).  tests/ then compares the
oracle (== engine, bit for bit) with every run and with the brute-force nearest-OCCUPIED distance (the reference's own
notion of correctness: Gnd_truth_checker::cmp_dist, include/gt_checker.h:30-80).

Run on a GPU box:  gpurun -- python oracle/wave_variance.py   (writes gpurun_out/wavevar/*.npz; copied to tests/golden/wavevar/).
Test infrastructure only.
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
from oracle import ref_io  # noqa: E402

K_RUNS = 6
CASES = [
    # name, cfg, size, cutoff, frames, dynamic
    ("pc_static", "cfg4", (48, 48, 24), 64, 5, False),
    ("pc_dynamic", "cfg4", (48, 40, 24), 64, 8, True),
    ("vlp16", "cfg2", (64, 64, 32), 49, 5, True),
    ("pc_dynamic_96", "cfg4", (96, 96, 48), 100, 8, True),
]


def main():
    gie = load_pkg()
    outdir = os.path.join(ROOT, "gpurun_out", "wavevar")
    os.makedirs(outdir, exist_ok=True)
    summary = {}
    for name, cname, size, cutoff, nframes, dynamic in CASES:
        cfg = gie.scenes.small_config(cname, size, cutoff_grids_sq=cutoff)
        frames = gie.scenes.make_frames(cfg, nframes, dynamic=dynamic)
        halo = int(math.ceil(math.sqrt(cutoff))) + 2
        runs = [ref_io.run(cfg, frames, "parity", halo=halo) for _ in range(K_RUNS)]
        save = {}
        rep = []
        for k in range(nframes):
            base = runs[0][k]
            known = base["glb_type"] != 0
            d_mis, id_mis, t_mis = [], [], []
            for r in runs[1:]:
                d_mis.append(int((r[k]["pair_dist"][known] != base["pair_dist"][known]).sum()))
                id_mis.append(int((r[k]["pair_id"][known] != base["pair_id"][known]).sum()))
                t_mis.append(int((r[k]["glb_type"] != base["glb_type"]).sum()))
            rep.append(dict(frame=k, known=int(known.sum()), self_dist_mismatch=d_mis, self_id_mismatch=id_mis, self_type_mismatch=t_mis))
            # run 0 is stored; the other runs only through their mismatch counts against it (self_* arrays)
            save[f"f{k}_glb_type"] = base["glb_type"]
            save[f"f{k}_pair_dist"] = base["pair_dist"]
            save[f"f{k}_pair_id"] = base["pair_id"]
            save[f"f{k}_box_type"] = base["box"]["type"]
            save[f"f{k}_self_dist_mismatch"] = np.array(d_mis)
            save[f"f{k}_self_id_mismatch"] = np.array(id_mis)
        np.savez_compressed(os.path.join(outdir, f"{name}.npz"), cfg_name=cname, size=np.array(size), cutoff=cutoff,
                            nframes=nframes, dynamic=dynamic, halo=halo, runs=K_RUNS, **save)
        summary[name] = rep
        print(name, json.dumps(rep), flush=True)
    with open(os.path.join(outdir, "self_variance.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
