"""Short cfg4 run for ncu: 2 warm frames + 2 profiled frames (device-resident input)."""
import sys, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
import torch
cfg = gie.scenes.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
frames = gie.scenes.make_frames(cfg, n)
mp = gie.Mapper(cfg)
for f in frames:
    mp.publishMap(f)
mp.hash_map.sync()
print("done", mp.loc_map.launch_count(), mp.hash_map.wave_stats())
