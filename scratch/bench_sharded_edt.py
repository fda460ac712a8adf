"""Sharded batch EDT (gie-mapping_b200/sharded.py) on N GPUs: ms per update, bytes exchanged, max over ranks.
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scratch/bench_sharded_edt.py [X Y Z]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
from conftest import load_pkg
gie = load_pkg()
import torch, torch.distributed as dist
from gie_mapping_b200 import sharded
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
X, Y, Z = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 1024, 1016)
density = float(sys.argv[4]) if len(sys.argv) >= 5 else 0.001
Zs = Z // world
eng = sharded.ShardedBatchEDT(0.1, (X, Y, Z), cutoff_grids_sq=2500)
g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
t = torch.where(torch.rand((Zs, Y, X), device="cuda", generator=g) < density, 2, 1).to(torch.int8)
eng.set_slab_types(t)
del t
for _ in range(3):
    eng.update()
torch.cuda.synchronize()
if world > 1: dist.barrier()
reps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b0 = eng.exchanged_bytes
e0.record()
for _ in range(reps):
    eng.update()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# stage split (events around the local sweeps only)
s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
s0.record(); eng.slab.edt_xy_sweeps(); s1.record(); eng.cols.edt_z_sweep(X + Y + Z); s2.record(); torch.cuda.synchronize()
tt = torch.tensor([ms, s0.elapsed_time(s1), s1.elapsed_time(s2)], device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    nvox = X * Y * Z
    print(json.dumps({"what": "sharded batch EDT", "n_gpus": world, "volume": [X, Y, Z], "occupancy": density, "ms_per_update": float(tt[0]),
                      "mvoxels_per_s": nvox / float(tt[0]) / 1e3, "xy_sweeps_ms": float(tt[1]), "z_sweep_ms": float(tt[2]),
                      "exchange_ms": float(tt[0] - tt[1] - tt[2]),
                      "bytes_out_per_rank_per_update": (eng.exchanged_bytes - b0) // reps}), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
