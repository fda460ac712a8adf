"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the frame on ragged sizes,
dynamic scene (all three wavefronts), the sharded emulation, the EDT checker, streaming."""
import sys, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
from gie_mapping_b200 import sharded
for name, size, cutoff in (("cfg4", (40, 64, 33), 64), ("cfg2", (64, 64, 32), 49), ("cfg3", (48, 32, 24), 36), ("cfg1", (64, 64, 16), 36)):
    cfg = gie.scenes.small_config(name, size, cutoff_grids_sq=cutoff)
    cfg["display_glb_edt"] = True
    frames = gie.scenes.make_frames(cfg, 5, dynamic=True)
    mp = gie.Mapper(cfg)
    for f in frames:
        mp.publishMap(f)
        mp.hash_map.streamPipeline()
    mp.hash_map.sync()
    print(name, mp.hash_map.wave_stats(), mp.hash_map.check_edt()["rms"], mp.hash_map.check_edt(glb=True)["n"])
    mp.loc_map.convertCostMap()
    mp.close()
cfg = gie.scenes.small_config("cfg4", (40, 64, 33), cutoff_grids_sq=64)
frames = gie.scenes.make_frames(cfg, 4, dynamic=True)
sm = sharded.ShardedMapper(cfg, emulate_slabs=2)
for f in frames:
    sm.publishMap(f)
sm.hash_map.sync()
print("sharded", sm.hash_map.wave_stats())
sm.close()
# dense batch EDT (serial z sweep with deep stacks -> ring spill path)
rng = np.random.RandomState(1)
t = np.where(rng.rand(96, 40, 72) < 0.05, 2, 1).astype(np.int8)
lm = gie.LocMap(0.1, (72, 40, 96))
lm.upload_glb_type(t); lm.batchEDTUpdate(); print("dense", int(lm.download(gie.ARR_AUX).max())); lm.close()
print("sanitize run done")
