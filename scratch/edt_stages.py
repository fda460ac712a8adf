"""Per-stage CUDA-event times of the batch EDT (and of the whole frame) for the headline scene and for a dense volume.
usage: python scratch/edt_stages.py [cfg4] [nframes]   (env GIE_XS_RPI=16|32 selects the x-sweep shape)"""
import sys, json, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cfg = gie.scenes.make_config(name)
X, Y, Z = cfg["local_size"]
out = {}
# dense: random occupancy in every slice
import os
for dens in (() if os.environ.get('GIE_STAGES_SCENE_ONLY') else (0.002, 0.02)):
    rng = np.random.RandomState(5)
    t = np.where(rng.rand(Z, Y, X) < dens, 2, 1).astype(np.int8)
    lm = gie.LocMap(cfg["voxel_width"], (X, Y, Z), cutoff_grids_sq=cfg["cutoff_grids_sq"])
    lm.upload_glb_type(t)
    lm.profile_enable(True)
    acc = {}
    for k in range(8):
        lm.batchEDTUpdate()
        if k >= 3:
            for s, v in lm.profile_last().items():
                acc[s] = acc.get(s, 0.0) + v / 5
    out[f"dense_{dens}"] = {k: round(v, 4) for k, v in acc.items() if k.startswith("edt")}
    lm.close()
# headline scene
frames = gie.scenes.make_frames(cfg, n)
mp = gie.Mapper(cfg)
mp.loc_map.profile_enable(True)
acc, cnt = {}, 0
for k, f in enumerate(frames):
    mp.publishMap(f)
    if k >= n - 10:
        cnt += 1
        for s, v in mp.loc_map.profile_last().items():
            acc[s] = acc.get(s, 0.0) + v
mp.hash_map.sync()
out["scene"] = {k: round(v / cnt, 4) for k, v in acc.items()}
out["scene"]["sum"] = round(sum(out["scene"].values()), 4)
out["wave_stats_last"] = mp.hash_map.wave_stats()
print(json.dumps(out))
