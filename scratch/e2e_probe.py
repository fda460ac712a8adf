"""Per-frame time of cfg4 in the two bench modes (device-resident pipelined vs host buffers + sync per frame)."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
import torch
cfg = gie.scenes.make_config("cfg4")
n = 25
frames = gie.scenes.make_frames(cfg, n)
pin = [torch.from_numpy(np.ascontiguousarray(f["points"], np.float32)).pin_memory() for f in frames]
dev = [p.cuda() for p in pin]
stream = torch.cuda.current_stream()

def run(mode):
    mp = gie.Mapper(cfg)
    mp.loc_map.set_stream(stream.cuda_stream)
    if mode == "prof":
        mp.loc_map.profile_enable(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    host = []
    stages = []
    ev[0].record(stream)
    for k, f in enumerate(frames):
        t0 = time.perf_counter()
        if mode == "e2e":
            g = dict(f); g["points"] = pin[k].numpy()
            mp.publishMap(g)
            mp.hash_map.sync()
            mp.hash_map.wave_stats()
        else:
            mp.publishMap(f, device_input=dev[k].data_ptr())
            if mode == "prof":
                stages.append(mp.loc_map.profile_last())
        ev[k + 1].record(stream)
        host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    gpu = [ev[k].elapsed_time(ev[k + 1]) for k in range(n)]
    print(mode, "gpu ms/frame :", " ".join(f"{x:6.2f}" for x in gpu))
    print(mode, "host ms/frame:", " ".join(f"{x:6.2f}" for x in host))
    if stages:
        for name in stages[0]:
            print(f"  {name:14s}", " ".join(f"{s[name]:6.2f}" for s in stages))
    print(mode, "stats", mp.hash_map.wave_stats(), "blocks", mp.hash_map.num_blocks())
    mp.close()

for mode in (sys.argv[1:] or ["dev", "e2e", "prof"]):
    run(mode)
