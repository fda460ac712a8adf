"""Generates tests/golden/*.npz on a GPU box by running the reference's OWN CUDA sources (oracle/_ref/ref_driver_parity,
built by oracle/build_ref.sh from /root/reference with IEEE float flags) on small seeded scenes, and reports how the CPU
oracle compares.  Run:  gpurun -- python oracle/gen_golden.py   (writes gpurun_out/golden/, copied to tests/golden/).

Stored per case: the config name/size, frame count, and per frame the reference's glb_type, batch dist_sq / coc
(_aux/_coc_idx_aux after MarkLimitedObserve), final (dist, coc id) pair and a halo box of hash voxels."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
from oracle import ref_io, oracle_py  # noqa: E402

CASES = [
    # name, cfg, size, cutoff, frames, dynamic
    ("pc_static", "cfg4", (48, 48, 24), 64, 5, False),
    ("pc_dynamic", "cfg4", (48, 40, 24), 64, 8, True),
    ("scan2d", "cfg1", (64, 64, 16), 100, 4, False),
    ("vlp16", "cfg2", (64, 64, 32), 49, 5, True),
    ("depth", "cfg3", (64, 64, 32), 100, 5, True),
]
HALO = 4


def main():
    gie = load_pkg()
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    report = {}
    for name, cname, size, cutoff, nframes, dynamic in CASES:
        cfg = gie.scenes.small_config(cname, size, cutoff_grids_sq=cutoff)
        frames = gie.scenes.make_frames(cfg, nframes, dynamic=dynamic)
        ref = ref_io.run(cfg, frames, "parity", halo=HALO)
        om = oracle_py.OracleMapper(cfg)
        rep = []
        save = {}
        for k, (f, r) in enumerate(zip(frames, ref)):
            om.publishMap(f)
            known = r["glb_type"] != 0
            opair_d = (om.pair >> np.uint64(32)).astype(np.int64)
            opair_id = (om.pair & np.uint64(0xffffffff)).astype(np.int64)
            rid = r["pair_id"].astype(np.int64) & 0xffffffff
            rep.append(dict(frame=k,
                            glb_type_mismatch=int((om.glb_type != r["glb_type"]).sum()),
                            known=int(known.sum()),
                            pair_dist_mismatch=int((opair_d[known] != r["pair_dist"][known]).sum()),
                            pair_id_mismatch=int((opair_id[known] != rid[known]).sum()),
                            edt_max_abs_diff=float(np.abs(om.edt[known] - r["edt"][known]).max()) if known.any() else 0.0,
                            stats=om.stats()))
            for key in ["glb_type", "aux", "coc_aux", "pair_dist", "pair_id"]:
                save[f"f{k}_{key}"] = r[key]
            save[f"f{k}_box"] = r["box"]
        om.close()
        np.savez_compressed(os.path.join(outdir, f"{name}.npz"), cfg_name=cname, size=np.array(size), cutoff=cutoff,
                            nframes=nframes, dynamic=dynamic, halo=HALO, **save)
        report[name] = rep
        print(name, json.dumps(rep))
    with open(os.path.join(outdir, "report.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
