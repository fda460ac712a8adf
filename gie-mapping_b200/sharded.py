"""One local volume sharded over the GPUs of a node (BASELINE configs[4]; no reference counterpart — the reference is
single-GPU).

What is cut, and why there.  Per frame the path has a DENSE half — the x and z sweeps of the batch EDT, which write 8 bytes
for every voxel of the volume (8.5 GB at 1024 x 1024 x 1016) — and a SPARSE half: ray cast, hash merge, limited-observation
mark, frontiers, wavefronts and commit only touch the observed region (a few percent of the volume) and are latency-bound BFS
levels of microseconds each; a per-level exchange between GPUs would cost more than a level does.  So:

  * the volume is cut into G slabs of rows y, rank g holds rows [g Y/G, (g+1) Y/G) of the batch-EDT arrays.  Along y a slab
    holds whole x rows and whole z columns, so BOTH sweeps run on the slab as it lies: no re-partition of the 8 B/voxel
    intermediate (round 1 moved N*16*(G-1)/G^2 bytes per rank per frame through two all-to-alls for this);
  * what the slabs need from the rest of the volume is the y pass, and that is 1 bit per voxel: rank 0, which owns the
    occupancy, runs it for the whole volume (bit words + links + column / slice lists, gie_edt_pack) and BROADCASTS the planes
    of the obstacle-bearing slices only (NCCL; the sweeps never read the other planes);
  * rank 0 keeps the hashed global map and runs the sparse half; where that needs the batch-EDT result (known voxels in
    MarkLimitedObserve, a few look-ups in waves A/B) it reads the owning slab directly — its own, or a peer GPU's array mapped
    through CUDA IPC and read over NVLink (gie_locmap_attach_slabs);
  * two stream-ordered NCCL operations per frame bracket the slab sweeps (the broadcast, and a 4-byte all-reduce that tells
    rank 0 every slab is written).  Nothing else synchronises: all work is enqueued on the current CUDA stream.

Results are bit-identical to the single-GPU engine (tests/test_sharded.py: in-process emulation with G slabs on one GPU, and
world-size-2 NCCL).  All arithmetic happens in libgie_b200.so; this module is host-side plumbing.
"""
import ctypes as C

import numpy as np

from .engine import LocMap, GlbHashMap, Mapper, GieError, load_library, _check, ARR_AUX, ARR_COC_AUX


class _DevView:
    """Zero-copy torch view of a device array owned by the engine."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(loc_map, which, shape, typestr="<i4"):
    """torch view of one of a map's device arrays (gie_locmap_device_ptr)."""
    import torch
    ptr, nbytes = loc_map.device_ptr(which)
    assert int(np.prod(shape)) * int(typestr[2:]) <= nbytes
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=torch.device("cuda", torch.cuda.current_device()))


def slab_layout(Y, G):
    """Rows of slab g: [g * Y/G, (g+1) * Y/G).  Slabs are word aligned (32 rows) because the y pass packs 32 rows per word."""
    if G < 1 or G > 8 or Y % (32 * G):
        raise GieError(f"Y = {Y} cannot be cut into {G} slabs of a multiple of 32 rows")
    rows = Y // G
    return [(g * rows, rows) for g in range(G)]


def exchange_edt_inputs(meta, ytab_planes, col_planes, Z, src=0, group=None):
    """The per-frame broadcast of the y pass from the owner to the slab ranks.  meta: int32 [2Z + 8] (columns per slice,
    slice list, number of obstacle-bearing slices); ytab_planes / col_planes: [Z, ...] buffers whose first n_slices planes
    travel.  Backend-agnostic (NCCL on GPUs, gloo in the CPU test).  Returns (n_slices, bytes received by a non-source rank)."""
    import torch.distributed as dist
    dist.broadcast(meta, src=src, group=group)
    ns = int(meta[2 * Z].item())          # the one host read of the frame on the slab ranks: the size of what follows
    if ns > 0:
        dist.broadcast(ytab_planes[:ns], src=src, group=group)
        dist.broadcast(col_planes[:ns], src=src, group=group)
    nbytes = meta.numel() * meta.element_size()
    if ns > 0:
        nbytes += ytab_planes[:ns].numel() * ytab_planes.element_size() + col_planes[:ns].numel() * col_planes.element_size()
    return ns, nbytes


class SlabMap:
    """Batch-EDT arrays of rows [row0, row0 + rows) of an X x Y x Z volume (gie_locmap_create_slab)."""

    def __init__(self, local_size, row0, rows):
        self.lib = load_library()
        self._h = C.c_void_p()
        X, Y, Z = (int(v) for v in local_size)
        self._local_size, self.row0, self.rows = (X, Y, Z), int(row0), int(rows)
        _check(self.lib.gie_locmap_create_slab(C.byref(self._h), X, Y, Z, int(row0), int(rows)))

    def close(self):
        if self._h:
            self.lib.gie_locmap_destroy(self._h)
            self._h = C.c_void_p()

    def set_stream(self, cuda_stream):
        _check(self.lib.gie_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def alias_inputs(self, owner):
        _check(self.lib.gie_slab_alias_inputs(self._h, owner._h))

    def set_compact(self, on):
        _check(self.lib.gie_slab_set_compact(self._h, int(on)))

    def input_buffers(self):
        """(ytab, col_list, meta) device pointers and byte sizes of the buffers the y pass is received into."""
        p = [C.c_void_p() for _ in range(3)]
        n = [C.c_size_t() for _ in range(3)]
        _check(self.lib.gie_slab_input_buffers(self._h, C.byref(p[0]), C.byref(n[0]), C.byref(p[1]), C.byref(n[1]), C.byref(p[2]), C.byref(n[2])))
        return [(p[i].value, n[i].value) for i in range(3)]

    def output_ptrs(self):
        out = []
        for which in (ARR_AUX, ARR_COC_AUX):
            ptr, nbytes = C.c_void_p(), C.c_size_t()
            _check(self.lib.gie_locmap_device_ptr(self._h, which, C.byref(ptr), C.byref(nbytes)))
            out.append(ptr.value)
        return out

    def ipc_handles(self):
        hs = []
        for ptr in self.output_ptrs():
            buf = (C.c_ubyte * 64)()
            _check(self.lib.gie_ipc_export(C.c_void_p(ptr), buf))
            hs.append(bytes(buf))
        return hs

    def sweeps(self, max_width):
        _check(self.lib.gie_edt_slab_sweeps(self._h, int(max_width)))

    def download(self, which):
        X, Y, Z = self._local_size
        out = np.empty(Z * self.rows * X, np.int32)
        _check(self.lib.gie_locmap_download(self._h, which, out.ctypes.data_as(C.c_void_p)))
        return out.reshape(Z, self.rows, X)


class ShardedMapper:
    """VOLMAPNODE::publishMap's call order (src/volumetric_mapper.cpp:138-224) on a volume sharded over `world` ranks, or —
    world == 1, emulate_slabs = G — over G slab maps inside this process on this GPU (the same code path minus NCCL and IPC,
    so that the slab arithmetic is tested on a one-GPU box).  rank 0 is the owner: hashed global map + sparse stages."""

    def __init__(self, cfg, rank=0, world=1, emulate_slabs=0, stream=None):
        self.cfg, self.rank, self.world = cfg, int(rank), int(world)
        X, Y, Z = cfg["local_size"]
        self.G = self.world if self.world > 1 else int(emulate_slabs)
        if self.G < 2:
            raise GieError("ShardedMapper needs at least 2 slabs (world > 1 or emulate_slabs >= 2); use Mapper otherwise")
        self.layout = slab_layout(Y, self.G)
        self.max_width = X + Y + Z
        self.is_owner = self.rank == 0
        self.owner = Mapper(cfg) if self.is_owner else None
        self.bytes_received = 0
        lib = load_library()
        if self.world == 1:
            self.slabs = [SlabMap((X, Y, Z), r0, n) for r0, n in self.layout]
            if stream is not None:
                self.owner.loc_map.set_stream(stream)
            for s in self.slabs:
                s.alias_inputs(self.owner.loc_map)       # same device: read the owner's y pass in place (also takes its stream)
            ptrs = [s.output_ptrs() for s in self.slabs]
            aux = (C.c_void_p * self.G)(*[p[0] for p in ptrs])
            coc = (C.c_void_p * self.G)(*[p[1] for p in ptrs])
            _check(lib.gie_locmap_attach_slabs(self.owner.loc_map._h, self.G, self.layout[0][1], aux, coc, None, None))
            return
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        dev = torch.device("cuda", torch.cuda.current_device())
        r0, n = self.layout[self.rank]
        self.slab = SlabMap((X, Y, Z), r0, n)
        cs = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        self.slab.set_stream(cs)
        WY = (Y + 31) // 32
        if self.is_owner:
            self.owner.loc_map.set_stream(cs)
            self.slab.alias_inputs(self.owner.loc_map)
            # send buffers: the planes of the obstacle-bearing slices, gathered by gie_edt_pack
            self.ytab_t = torch.empty((Z, WY * X), dtype=torch.int64, device=dev)
            self.col_t = torch.empty((Z, X), dtype=torch.int32, device=dev)
            (_, _), (_, _), (mp, mb) = self._owner_meta()
            self.meta_t = torch.as_tensor(_DevView(mp, (mb // 4,), "<i4"), device=dev)
        else:
            (yp, yb), (cp, cb), (mp, mb) = self.slab.input_buffers()
            self.ytab_t = torch.as_tensor(_DevView(yp, (Z, WY * X), "<i8"), device=dev)
            self.col_t = torch.as_tensor(_DevView(cp, (Z, X), "<i4"), device=dev)
            self.meta_t = torch.as_tensor(_DevView(mp, (mb // 4,), "<i4"), device=dev)
            self.slab.set_compact(True)
        self.token = torch.zeros(1, dtype=torch.int32, device=dev)
        # the owner maps every peer slab's output arrays (CUDA IPC) and reads them over NVLink
        mine = self.slab.ipc_handles()
        gathered = [None] * self.world
        dist.all_gather_object(gathered, mine)
        if self.is_owner:
            own = self.slab.output_ptrs()
            aux = (C.c_void_p * self.G)(*([own[0]] + [None] * (self.G - 1)))
            coc = (C.c_void_p * self.G)(*([own[1]] + [None] * (self.G - 1)))
            ah = b"".join(g[0] for g in gathered)
            ch = b"".join(g[1] for g in gathered)
            _check(lib.gie_locmap_attach_slabs(self.owner.loc_map._h, self.G, n, aux, coc, ah, ch))
        dist.barrier()

    def _owner_meta(self):
        # the owner's edt_meta lives in its LocMap; the aliased slab reports the same buffers
        return self.slab.input_buffers()

    # -----------------------------------------------------------------------------------------------------------------
    def publishMap(self, frame, device_input=None):
        lib = load_library()
        if self.world == 1:
            self.owner.integrate(frame, device_input)
            _check(lib.gie_edt_pack(self.owner.loc_map._h, None, None))
            for s in self.slabs:
                s.sweeps(self.max_width)
            self.owner.hash_map.mergeNewObsv(self.owner._time, self.cfg.get("display_glb_edt", False))
            return
        Z = self.cfg["local_size"][2]
        if self.is_owner:
            self.owner.integrate(frame, device_input)
            _check(lib.gie_edt_pack(self.owner.loc_map._h, C.c_void_p(self.ytab_t.data_ptr()), C.c_void_p(self.col_t.data_ptr())))
        _, nbytes = exchange_edt_inputs(self.meta_t, self.ytab_t, self.col_t, Z, src=0)
        if not self.is_owner:
            self.bytes_received += nbytes
        self.slab.sweeps(self.max_width)
        self.dist.all_reduce(self.token)           # stream-ordered: the owner's merge below starts after every slab is written
        if self.is_owner:
            self.owner.hash_map.mergeNewObsv(self.owner._time, self.cfg.get("display_glb_edt", False))

    # convenience for tests
    @property
    def loc_map(self):
        return self.owner.loc_map

    @property
    def hash_map(self):
        return self.owner.hash_map

    def close(self):
        if self.owner is not None:
            self.owner.close()          # also unmaps the peers' slabs (CUDA IPC) ...
        if self.world > 1:
            self.dist.barrier()         # ... before their owners free them
        for s in getattr(self, "slabs", []) + ([self.slab] if hasattr(self, "slab") else []):
            s.close()
