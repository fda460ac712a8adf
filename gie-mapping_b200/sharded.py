"""Batch EDT of a local volume sharded across the GPUs of one node (no reference counterpart — the reference
is single-GPU).

One process per GPU (torch.distributed, NCCL over NVLink).  Rank r owns the z-slab  z in [r*Z/G, (r+1)*Z/G)  of the
X x Y x Z volume.  The y and x sweeps of the separable EDT (EDTphase1/2, reference src/kernel/edt/local_edt_core.h:14-135)
only look inside one z slice, so they run on the slab as it lies.  The z sweep (EDTphase3, :137-193) needs whole z columns:
the packed intermediate (in-slice squared distance + closest obstacle of the slice, 8 B/voxel) is re-partitioned from
z-slabs to y-slabs with ONE all-to-all, the z sweep runs on the y-slab, and a second all-to-all returns (dist_sq, coc) to
the z-slab owners.  Per rank N*8*(G-1)/G^2 bytes leave in each direction.  The result is bit-identical to the single-GPU
gie_edt_batch_update of the whole volume (tests/test_sharded.py).

Layout facts that make the exchange copy-free on the receive side: a z-slab [Zs][Y][X] splits along y into G blocks
[Zs][Ys][X]; block d goes to rank d; rank d receives G such blocks ordered by source rank = ordered by z, and their
concatenation IS its y-slab array [Z][Ys][X].

This module is host-side plumbing only (pointer wrapping, all-to-all calls); all arithmetic happens in libgie_b200.so.
"""
import numpy as np
import torch
import torch.distributed as dist

from .engine import LocMap, ARR_GLB_TYPE, ARR_AUX, ARR_COC_AUX, ARR_EDT_G2, ARR_EDT_CXY, ARR_EDT_NCOLS


class _DevView:
    """Zero-copy torch view of a device array owned by the engine."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(loc_map, which, shape, typestr="<i4"):
    ptr, nbytes = loc_map.device_ptr(which)
    assert int(np.prod(shape)) * int(typestr[2:]) <= nbytes
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=torch.device("cuda", torch.cuda.current_device()))


def repartition_z_to_y(slab, cols, group=None, async_op=False):
    """z-slab [Zs, Y, X] on every rank -> y-slab [Z, Ys, X] on every rank (one all-to-all).  Works on any device/backend.
    Returns (bytes leaving this rank, work handle or None)."""
    G = dist.get_world_size(group)
    Zs, Y, X = slab.shape
    send = slab.view(Zs, G, Y // G, X).permute(1, 0, 2, 3).contiguous()      # block d = the rows of y-owner d
    # rank d receives G blocks [Zs, Ys, X] ordered by source rank = ordered by z: their concatenation is its y-slab
    work = dist.all_to_all_single(cols.view(-1), send.view(-1), group=group, async_op=async_op)
    return send.numel() * send.element_size() * (G - 1) // G, work


def repartition_y_to_z(cols, slab, scratch=None, group=None):
    """y-slab [Z, Ys, X] on every rank -> z-slab [Zs, Y, X] on every rank (one all-to-all + a local block transpose).
    Returns the bytes leaving this rank."""
    G = dist.get_world_size(group)
    Z, Ys, X = cols.shape
    Zs = Z // G
    if scratch is None:
        scratch = torch.empty((G, Zs, Ys, X), dtype=cols.dtype, device=cols.device)
    dist.all_to_all_single(scratch.view(-1), cols.view(-1), group=group)       # chunk d of cols = the slices of z-owner d
    slab.view(Zs, G, Ys, X).copy_(scratch.permute(1, 0, 2, 3))
    return cols.numel() * cols.element_size() * (G - 1) // G


class ShardedBatchEDT:
    """EDT_OCC::batchEDTUpdate over a volume whose z-slabs live on different GPUs."""

    def __init__(self, voxel_size, size_xyz, cutoff_grids_sq=100, group=None):
        self.group = group
        self.G = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        X, Y, Z = (int(v) for v in size_xyz)
        if Y % self.G or Z % self.G:
            raise ValueError(f"Y={Y} and Z={Z} must be multiples of the number of GPUs ({self.G})")
        self.X, self.Y, self.Z = X, Y, Z
        self.Zs, self.Ys = Z // self.G, Y // self.G
        self.slab = LocMap(voxel_size, (X, Y, self.Zs), cutoff_grids_sq=cutoff_grids_sq)        # my z-slab
        self.cols = LocMap(voxel_size, (X, self.Ys, Z), cutoff_grids_sq=cutoff_grids_sq)        # my y-slab, whole z columns
        stream = torch.cuda.current_stream().cuda_stream
        self.slab.set_stream(stream)
        self.cols.set_stream(stream)
        sh_slab, sh_cols = (self.Zs, Y, X), (Z, self.Ys, X)
        self.t_type = device_tensor(self.slab, ARR_GLB_TYPE, sh_slab, "|i1")
        self.t_g2, self.t_cxy = device_tensor(self.slab, ARR_EDT_G2, sh_slab), device_tensor(self.slab, ARR_EDT_CXY, sh_slab)
        self.t_ncols = device_tensor(self.slab, ARR_EDT_NCOLS, (self.Zs,))
        self.t_aux, self.t_coc = device_tensor(self.slab, ARR_AUX, sh_slab), device_tensor(self.slab, ARR_COC_AUX, sh_slab)
        self.c_g2, self.c_cxy = device_tensor(self.cols, ARR_EDT_G2, sh_cols), device_tensor(self.cols, ARR_EDT_CXY, sh_cols)
        self.c_ncols = device_tensor(self.cols, ARR_EDT_NCOLS, (Z,))
        self.c_aux, self.c_coc = device_tensor(self.cols, ARR_AUX, sh_cols), device_tensor(self.cols, ARR_COC_AUX, sh_cols)
        self._back = [torch.empty((self.G, self.Zs, self.Ys, X), dtype=torch.int32, device=self.t_aux.device) for _ in range(2)]
        self.exchanged_bytes = 0

    def close(self):
        self.slab.close()
        self.cols.close()

    def set_slab_types(self, glb_type_slab):
        """glb_type of my z-slab: int8 [Zs, Y, X] torch tensor on this device or numpy array."""
        if isinstance(glb_type_slab, np.ndarray):
            glb_type_slab = torch.from_numpy(np.ascontiguousarray(glb_type_slab, np.int8)).to(self.t_type.device)
        self.t_type.copy_(glb_type_slab.view(self.Zs, self.Y, self.X))

    def update(self):
        """Batch EDT of the whole volume; afterwards result() holds (dist_sq, coc) of my z-slab."""
        G = self.G
        self.slab.edt_xy_sweeps()
        if G == 1:
            self.c_ncols.copy_(self.t_ncols)
            self.c_g2.copy_(self.t_g2.view_as(self.c_g2))
            self.c_cxy.copy_(self.t_cxy.view_as(self.c_cxy))
        else:
            dist.all_gather_into_tensor(self.c_ncols, self.t_ncols.contiguous(), group=self.group)
            works = []
            for src, dst in ((self.t_g2, self.c_g2), (self.t_cxy, self.c_cxy)):   # the second pack overlaps the first transfer
                nbytes, w = repartition_z_to_y(src, dst, self.group, async_op=True)
                self.exchanged_bytes += nbytes
                works.append(w)
            for w in works:
                w.wait()
        self.cols.edt_z_sweep(self.X + self.Y + self.Z)
        for k, (src, dst) in enumerate(((self.c_aux, self.t_aux), (self.c_coc, self.t_coc))):
            if G == 1:
                dst.copy_(src.view_as(dst))
            else:
                self.exchanged_bytes += repartition_y_to_z(src, dst, self._back[k], self.group)

    def result(self):
        """(dist_sq int32 [Zs, Y, X], coc int32 [Zs, Y, X]) of my z-slab as numpy arrays; coc = x | y << 11 | z << 22 with
        z counted in the WHOLE volume."""
        torch.cuda.synchronize()
        return self.t_aux.cpu().numpy(), self.t_coc.cpu().numpy()
