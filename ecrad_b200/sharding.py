"""Column sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

Columns are independent end to end (the reference already block-decomposes by istartcol:iendcol,
driver/ecrad_driver.F90:345-354), so the data path needs no collective: rank r owns a contiguous column range and
writes its own slice of flux_type.  The only exchange steps are the ones the reference's MPL layer has
(ifsrrtm/rrtm_kgb*.F90: rank 0 reads the tables, MPL_BROADCAST) and, where the host model wants all fluxes in one place,
a gather of the flux profiles; both are provided here over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(ncol_total: int, rank: int, world: int):
    """1-based inclusive (istartcol, iendcol) of rank `rank`: contiguous blocks of ceil(N/G) columns, like the offline
    driver's blocks.  Returns (1, 0) for ranks beyond the data."""
    per = -(-ncol_total // world)
    start = rank * per + 1
    end = min((rank + 1) * per, ncol_total)
    return (start, end) if start <= end else (1, 0)


def broadcast_table_blob(path: str, dist, device=None) -> bytes:
    """Rank 0 reads the packed ETB1 table blob, every rank receives the same bytes (one broadcast of ~2.5 MB)."""
    import torch

    rank = dist.get_rank()
    n = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == 0:
        raw = np.fromfile(path, dtype=np.uint8)
        n[0] = raw.size
    dist.broadcast(n, src=0)
    buf = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    if rank == 0:
        buf.copy_(torch.from_numpy(raw))
    dist.broadcast(buf, src=0)
    return buf.cpu().numpy().tobytes()


def gather_profiles(local, ncol_total: int, dist, dst: int = 0):
    """Gather a (nrows, ncol_local) tensor of flux profiles (row = half-level, column fastest: the reference layout of
    flux%lw_up etc.) from every rank's contiguous column range into (nrows, ncol_total) on rank `dst`.
    Ragged last shard is handled by padding to the common shard width."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-ncol_total // world)
    nrows = local.shape[0]
    send = torch.zeros((nrows, per), dtype=local.dtype, device=local.device)
    send[:, : local.shape[1]] = local
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((nrows, ncol_total), dtype=local.dtype, device=local.device)
    for r in range(world):
        s, e = shard_range(ncol_total, r, world)
        if e >= s:
            out[:, s - 1:e] = recv[r][:, : e - s + 1]
    return out


def gather_slab(local, dist, dst: int = 0, out=None):
    """ONE collective for all flux profiles: every rank's contiguous slab `local` (any shape, the same on all ranks: e.g.
    (nprofiles, nlev+1, ncol_local), what the flux kernels wrote) lands in out[rank] of a preallocated (world, *local.shape)
    tensor on `dst` -- NCCL's grouped send/recv straight into the destination, no padding, no per-profile calls, no copy-out.
    The result is blocked by rank exactly like the host model's column blocks (out[r] = the columns shard_range(.., r, world))."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    if rank == dst:
        if out is None:
            out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        dist.gather(local, [out[r] for r in range(world)], dst=dst)
        return out
    dist.gather(local, None, dst=dst)
    return None


# ---------------------------------------------------------------------------------------------------------
# Gather fused into the flux kernels: every rank stores its column slice straight into arrays that live on one GPU
# ---------------------------------------------------------------------------------------------------------
class PeerFluxArrays:
    """Flux profiles of ALL columns in the memory of rank `dst`, writable by every rank of the node over NVLink.

    Rank `dst` allocates one device arena holding `names` arrays of shape (nrows, ncol_total) (row = half-level, column
    fastest: the reference layout of flux%lw_up etc.); its CUDA IPC handle goes to the other ranks with one broadcast; each
    rank opens it (cudaIpcOpenMemHandle maps the peer memory into its address space) and hands
    `ecrad_b200_radiation_device_ld` pointers to the column slice it owns, with ld_out = ncol_total.  The kernels that produce
    the fluxes then write them where the host model wants them -- the reference's block decomposition
    (driver/ecrad_driver.F90:345-354) without a gather step.  No data-path collective; one barrier tells `dst` the step is done.
    """

    def __init__(self, names, nrows, ncol_total, dist, dst=0):
        import torch
        from cuda.bindings import runtime as rt

        self.rt, self.dist, self.dst = rt, dist, dst
        self.names, self.nrows, self.ncol_total = list(names), int(nrows), int(ncol_total)
        self.rank = dist.get_rank()
        self.plane = self.nrows * self.ncol_total * 8
        nbytes = self.plane * len(self.names)
        self.owner = self.rank == dst
        # every step is followed by an agreement over the process group, so that a failure on one rank (no peer access, IPC
        # not permitted in a sandbox, ...) makes ALL ranks raise instead of leaving the others in a barrier
        def agree(ok, what):
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                raise RuntimeError("PeerFluxArrays: " + what + " failed on at least one rank")

        ptr, payload, ok = 0, [None], True
        if self.owner:
            err, ptr = rt.cudaMalloc(nbytes)
            ok = err == rt.cudaError_t.cudaSuccess
            if ok:
                rt.cudaMemset(ptr, 0xFF, nbytes)   # NaN pattern: unwritten columns would show
                err, handle = rt.cudaIpcGetMemHandle(ptr)
                ok = err == rt.cudaError_t.cudaSuccess
                if ok:
                    payload = [bytes(handle.reserved)]
        agree(ok, "allocation / cudaIpcGetMemHandle")
        dist.broadcast_object_list(payload, src=dst)
        if not self.owner:
            handle = rt.cudaIpcMemHandle_t()
            handle.reserved = payload[0]
            err, ptr = rt.cudaIpcOpenMemHandle(handle, rt.cudaIpcMemLazyEnablePeerAccess)
            ok = err == rt.cudaError_t.cudaSuccess
        agree(ok, "cudaIpcOpenMemHandle")
        self.base = int(ptr)
        self._torch = torch

    def pointer(self, name, first_col):
        """Device address of element (row 0, column first_col) of array `name` (0-based column in the global numbering)."""
        return self.base + self.names.index(name) * self.plane + 8 * int(first_col)

    def tensor(self, name):
        """The whole (nrows, ncol_total) array as a torch tensor (owner rank only)."""
        assert self.owner
        torch = self._torch

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.nrows, self.ncol_total), "typestr": "<f8", "version": 3,
                                        "data": (self.base + self.names.index(name) * self.plane, False)}
        return torch.as_tensor(raw, device=torch.device("cuda", torch.cuda.current_device()))

    def slice_view(self, first_col, ncol_local):
        """(len(names), nrows, ncol_local) strided view of the columns [first_col, first_col + ncol_local) of ALL arrays, on whatever
        rank calls it (peer-mapped memory on the non-owners): `view.copy_(local_slab)` pushes a rank's fluxes to the owner with wide
        coalesced stores over NVLink, rows of ncol_local * 8 bytes at a time."""
        torch = self._torch

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (len(self.names), self.nrows, int(ncol_local)), "typestr": "<f8", "version": 3,
                                        "strides": (self.plane, self.ncol_total * 8, 8), "data": (self.base + 8 * int(first_col), False)}
        return torch.as_tensor(raw, device=torch.device("cuda", torch.cuda.current_device()))

    def close(self):
        self.dist.barrier()
        if self.owner:
            self.rt.cudaFree(self.base)
        else:
            self.rt.cudaIpcCloseMemHandle(self.base)
