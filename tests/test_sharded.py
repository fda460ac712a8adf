"""One volume sharded over GPUs (gie-mapping_b200/sharded.py): the per-frame broadcast protocol on CPU with gloo
(world_size 2), the slab arithmetic on ONE GPU (G slab maps in one process: same kernels, same pointer tables, no NCCL / IPC),
and the real thing on 2 GPUs over NCCL + CUDA IPC (skipped when fewer are visible)."""
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


def _exchange_worker(rank, world, port):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    load_pkg()
    from gie_mapping_b200 import sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    Z, WYX, X = 12, 2 * 7, 7
    rng = np.random.RandomState(11)
    for ns in (0, 5, 12):                                    # nothing to send, a few planes, every plane
        ref_meta = rng.randint(0, 99, 2 * Z + 8).astype(np.int32)
        ref_meta[2 * Z] = ns
        ref_ytab = rng.randint(-2 ** 40, 2 ** 40, (Z, WYX)).astype(np.int64)
        ref_col = rng.randint(0, X, (Z, X)).astype(np.int32)
        meta = torch.from_numpy(ref_meta.copy()) if rank == 0 else torch.full((2 * Z + 8,), -1, dtype=torch.int32)
        ytab = torch.from_numpy(ref_ytab.copy()) if rank == 0 else torch.full((Z, WYX), -7, dtype=torch.int64)
        col = torch.from_numpy(ref_col.copy()) if rank == 0 else torch.full((Z, X), -7, dtype=torch.int32)
        got_ns, nbytes = sharded.exchange_edt_inputs(meta, ytab, col, Z, src=0)
        assert got_ns == ns and nbytes == (2 * Z + 8) * 4 + ns * (WYX * 8 + X * 4)
        assert np.array_equal(meta.numpy(), ref_meta)
        assert np.array_equal(ytab.numpy()[:ns], ref_ytab[:ns]) and np.array_equal(col.numpy()[:ns], ref_col[:ns])
        if rank != 0:                                        # planes beyond n_slices do not travel
            assert (ytab.numpy()[ns:] == -7).all() and (col.numpy()[ns:] == -7).all()
    dist.destroy_process_group()


def test_exchange_protocol_gloo_world2():
    import torch.multiprocessing as mp
    mp.spawn(_exchange_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_slab_layout(gie):
    from gie_mapping_b200 import sharded
    assert sharded.slab_layout(1024, 8) == [(128 * g, 128) for g in range(8)]
    assert sharded.slab_layout(64, 2) == [(0, 32), (32, 32)]
    for bad in ((48, 2), (64, 4), (1024, 9), (1024, 0)):
        with pytest.raises(gie.GieError):
            sharded.slab_layout(*bad)


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


def _cmp_sharded_frame(gie, mp, om, tag):
    lm = mp.loc_map
    mp.hash_map.sync()
    for which, ref, name in [(gie.ARR_GLB_TYPE, om.glb_type, "glb_type"), (gie.ARR_AUX, om.aux, "aux"), (gie.ARR_COC_AUX, om.coc_aux, "coc_aux"),
                             (gie.ARR_PAIR, om.pair, "pair")]:
        got = lm.download(which)
        bad = int((got != ref).sum())
        assert bad == 0, f"{tag}: {name} differs in {bad} voxels"
    known = om.glb_type != 0
    assert np.array_equal(lm.download(gie.ARR_EDT)[known].view(np.uint32), om.edt[known].view(np.uint32)), f"{tag}: edt differs"
    assert mp.hash_map.wave_stats() == om.stats(), f"{tag}: {mp.hash_map.wave_stats()} vs {om.stats()}"


@pytest.mark.gpu
@pytest.mark.parametrize("name,size,cutoff,G,nframes", [
    ("cfg4", (48, 64, 24), 64, 2, 8),        # ray cast, all three wavefronts
    ("cfg4", (40, 128, 33), 100, 4, 6),      # 4 slabs, ragged X and Z
    ("cfg2", (64, 64, 32), 49, 2, 5),        # projective sensor
])
def test_sharded_emulation_parity(gie, oracle, name, size, cutoff, G, nframes):
    """The sharded frame with G slab maps inside one process on one GPU — the slab kernels, the pointer tables the sparse
    stages read the batch-EDT result through, the pack / sweep split — against the oracle, every array, every frame."""
    from gie_mapping_b200 import sharded
    cfg = gie.scenes.small_config(name, size, cutoff_grids_sq=cutoff)
    frames = gie.scenes.make_frames(cfg, nframes, dynamic=True)
    mp = sharded.ShardedMapper(cfg, emulate_slabs=G)
    om = oracle.OracleMapper(cfg)
    try:
        waves = 0
        for k, f in enumerate(frames):
            mp.publishMap(f)
            om.publishMap(f)
            _cmp_sharded_frame(gie, mp, om, f"{name} G={G} frame {k}")
            st = om.stats()
            waves += st["levelsA"] + st["levelsB"] + st["levelsC"]
        assert waves > 0 or name != "cfg4"
    finally:
        mp.close()
        om.close()


def _sharded_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    gie = load_pkg()
    from gie_mapping_b200 import sharded
    from oracle import oracle_py
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    cfg = gie.scenes.small_config("cfg4", (48, 64 * world // 2, 24), cutoff_grids_sq=64)
    frames = gie.scenes.make_frames(cfg, 8, dynamic=True)
    mp = sharded.ShardedMapper(cfg, rank=rank, world=world)
    om = oracle_py.OracleMapper(cfg) if rank == 0 else None
    ok, why = 1, ""
    try:
        for k, f in enumerate(frames):
            mp.publishMap(f)
            if rank == 0:
                om.publishMap(f)
                try:
                    _cmp_sharded_frame(gie, mp, om, f"world {world} frame {k}")
                except AssertionError as e:
                    ok, why = 0, str(e)
                    break
    finally:
        flag = torch.tensor([ok], device="cuda")
        dist.broadcast(flag, src=0)
        if rank == 0:
            q.put((int(flag.item()), why, mp.hash_map.wave_stats() if ok else None))
        mp.close()
        if om is not None:
            om.close()
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_two_gpus_nccl_ipc():
    """2 processes, 2 GPUs: the y pass travels by NCCL broadcast, rank 0 reads rank 1's slab through a CUDA IPC mapping; every
    array on every frame equals the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    mp.spawn(_sharded_worker, args=(2, _free_port(), q), nprocs=2, join=True)
    ok, why, stats = q.get(timeout=10)
    assert ok == 1, why
