"""Multi-GPU batch EDT: the z-slab <-> y-slab re-partition on CPU with gloo (world_size 2), the staged C ABI
on one GPU, and the sharded EDT against the oracle on 2 GPUs (skipped when fewer are visible)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _repartition_worker(rank, world, port, shape):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    load_pkg()
    from gie_mapping_b200 import sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    Z, Y, X = shape
    full = torch.from_numpy(np.random.RandomState(3).randint(-50, 50, size=shape).astype(np.int32))
    Zs, Ys = Z // world, Y // world
    slab = full[rank * Zs:(rank + 1) * Zs].contiguous()
    cols = torch.empty((Z, Ys, X), dtype=torch.int32)
    sent, _ = sharded.repartition_z_to_y(slab, cols)
    assert torch.equal(cols, full[:, rank * Ys:(rank + 1) * Ys, :]), "y-slab content"
    assert sent == slab.numel() * 4 * (world - 1) // world
    back = torch.zeros_like(slab)
    sharded.repartition_y_to_z(cols, back)
    assert torch.equal(back, slab), "round trip"
    dist.destroy_process_group()


def test_repartition_gloo_world2():
    import torch.multiprocessing as mp
    mp.spawn(_repartition_worker, args=(2, _free_port(), (8, 6, 5)), nprocs=2, join=True)


@pytest.mark.gpu
def test_staged_edt_equals_fused(gie, oracle):
    """gie_edt_xy_sweeps + gie_edt_z_sweep on one map == gie_edt_batch_update == oracle."""
    Z, Y, X = 40, 50, 33
    rng = np.random.RandomState(5)
    t = np.where(rng.rand(Z, Y, X) < 0.01, 2, 1).astype(np.int8)
    lm = gie.LocMap(0.1, (X, Y, Z))
    om = oracle.OracleMapper(dict(local_size=(X, Y, Z), voxel_width=0.1, cutoff_grids_sq=100))
    try:
        lm.upload_glb_type(t)
        lm.edt_xy_sweeps()
        lm.edt_z_sweep()
        om.set_glb_type(t)
        om.batch_edt()
        assert np.array_equal(lm.download(gie.ARR_AUX), om.aux) and np.array_equal(lm.download(gie.ARR_COC_AUX), om.coc_aux)
    finally:
        lm.close()
        om.close()


def _sharded_worker(rank, world, port, shape, density, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    load_pkg()
    from gie_mapping_b200 import sharded
    from oracle import oracle_py
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    Z, Y, X = shape
    rng = np.random.RandomState(17)
    t = np.where(rng.rand(Z, Y, X) < density, 2, 1).astype(np.int8)
    if density < 0.001:
        t[: Z // 2] = 1            # sparse case: rank 0's slab holds no obstacle at all
    Zs = Z // world
    eng = sharded.ShardedBatchEDT(0.1, (X, Y, Z))
    eng.set_slab_types(t[rank * Zs:(rank + 1) * Zs])
    eng.update()
    d, c = eng.result()
    eng.close()
    om = oracle_py.OracleMapper(dict(local_size=(X, Y, Z), voxel_width=0.1, cutoff_grids_sq=100))
    om.set_glb_type(t)
    om.batch_edt()
    ok = np.array_equal(d, om.aux[rank * Zs:(rank + 1) * Zs]) and np.array_equal(c, om.coc_aux[rank * Zs:(rank + 1) * Zs])
    om.close()
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,density", [((48, 64, 96), 0.01), ((32, 40, 50), 0.0005), ((64, 64, 64), 0.2)])
def test_sharded_batch_edt_two_gpus(shape, density):
    """2 GPUs, NCCL: every rank's z-slab of the sharded batch EDT equals the oracle's EDT of the whole volume bit for bit
    (dense, sparse with an obstacle-free slab, ragged X)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    mp.spawn(_sharded_worker, args=(2, _free_port(), shape, density, q), nprocs=2, join=True)
    assert q.get(timeout=10) == 1
