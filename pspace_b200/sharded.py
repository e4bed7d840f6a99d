"""Quadrature-point sharding across GPUs: one process per GPU, torch.distributed plumbing.

The sum over q is associative, so rank r projects only its contiguous slice [q0_r, q1_r) of the tensor grid
(it needs only the rows J[q0_r:q1_r]) and produces a full partial block C_r; the partial blocks are combined
by ONE collective over the coefficient blocks:

    allreduce        every rank ends with the full C                      (N^2 m^2 doubles; cfg5: 377 MB)
    reduce_scatter   rank r ends with its block of basis-pair rows of C   (when C is large / consumed row-sharded)

There is no collective on the data path before that (shards are independent), and the integer / table kernels
are replicated per rank.  The reference has no counterpart: it never distributes quadrature points.
`combine` works on any backend (nccl on GPUs; gloo in the CPU tests).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist

ALIGN = 32  # shard boundaries on multiples of the kernel's K-chunk (32 points) keep every chunk full


def shard_bounds(Q: int, world: int, align: int = ALIGN) -> List[Tuple[int, int]]:
    """Contiguous, near-equal q-ranges, boundaries aligned to `align` (the last range takes the remainder)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    units = (Q + align - 1) // align
    out, start = [], 0
    for r in range(world):
        n_units = units // world + (1 if r < units % world else 0)
        end = min(Q, start + n_units * align)
        out.append((start, end))
        start = end
    assert out[-1][1] == Q and sum(b - a for a, b in out) == Q
    return out


def row_blocks(nrows: int, world: int) -> List[Tuple[int, int]]:
    """Equal row blocks for reduce_scatter (all ranks must hold equal-sized pieces: rows are padded up)."""
    per = (nrows + world - 1) // world
    return [(min(nrows, r * per), min(nrows, (r + 1) * per)) for r in range(world)]


def combine(partial: torch.Tensor, mode: str = "allreduce", group=None):
    """Sum the per-rank partial coefficient blocks.  partial: [rows, ...] real or complex tensor.

    allreduce: in place, returns `partial` holding the total.
    reduce_scatter: returns (my_rows, (r0, r1)) — the rank's row block of the total.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if mode == "allreduce":
            return partial
        return partial, (0, partial.shape[0])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.view_as_real(partial) if partial.is_complex() else partial
    if mode == "allreduce":
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return partial
    if mode != "reduce_scatter":
        raise ValueError(mode)
    rows = partial.shape[0]
    per = (rows + world - 1) // world
    flat = t.reshape(rows, -1)
    if per * world != rows:     # pad to equal blocks
        pad = torch.zeros((per * world - rows, flat.shape[1]), dtype=flat.dtype, device=flat.device)
        flat = torch.cat([flat, pad], 0)
    mine = torch.empty((per, flat.shape[1]), dtype=flat.dtype, device=flat.device)
    dist.reduce_scatter_tensor(mine, flat.contiguous(), op=dist.ReduceOp.SUM, group=group)
    r0, r1 = min(rows, rank * per), min(rows, (rank + 1) * per)
    mine = mine[: r1 - r0].reshape((r1 - r0,) + tuple(t.shape[1:]))
    if partial.is_complex():
        mine = torch.view_as_complex(mine.contiguous())
    return mine, (r0, r1)


class SharedPinnedOutput:
    """One host buffer for the whole job: a POSIX shared-memory file mapped by every rank of the node and page-locked in
    each process (cudaHostRegister), so that rank r's device-to-host copy lands directly in rows [r0, r1) of the result
    every rank (and the caller) sees."""

    def __init__(self, shape, dtype, group=None, tag="pspace_C"):
        import os
        self.shape, self.dtype = tuple(shape), dtype
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        name = [f"/dev/shm/{tag}_{os.getpid()}"] if rank == 0 else [None]
        if world > 1:
            dist.broadcast_object_list(name, src=0, group=group)
        self.path, self.owner = name[0], rank == 0
        n = 1
        for d in self.shape:
            n *= d
        width = 2 if dtype.is_complex else 1
        if self.owner:
            with open(self.path, "wb") as f:
                f.truncate(n * width * 8)
        if world > 1:
            dist.barrier(group=group)
        flat = torch.from_file(self.path, shared=True, size=n * width, dtype=torch.float64)
        self._flat = flat
        rc = torch.cuda.cudart().cudaHostRegister(flat.data_ptr(), flat.numel() * 8, 0)
        self.registered = int(rc) == 0 if not hasattr(rc, "value") else rc.value == 0
        self.tensor = torch.view_as_complex(flat.view(*self.shape, 2)) if dtype.is_complex else flat.view(*self.shape)
        if world > 1:
            dist.barrier(group=group)

    def close(self):
        import os
        try:
            torch.cuda.cudart().cudaHostUnregister(self._flat.data_ptr())
        except Exception:
            pass
        if self.owner:
            try:
                os.unlink(self.path)
            except OSError:
                pass


class ShardedProjector:
    """Rank-local projector: owns a DeviceContainer and this rank's q-range."""

    def __init__(self, container, rank: int = None, world: int = None, group=None):
        self.c = container
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.q0, self.q1 = shard_bounds(container.Q, self.world)[self.rank]

    @property
    def local_points(self):
        return self.q1 - self.q0

    def projectMatrix(self, J_local, i0=0, i1=None, j0=0, j1=None, out=None, mode="allreduce"):
        """J_local: device rows [q0,q1) of J.  Returns the combined block (allreduce) or (rows, range)."""
        C = self.c.projectMatrix(J_local, i0, i1, j0, j1, q0=self.q0, q1=self.q1, out=out)
        return combine(C, mode, self.group)

    def projectMatrixHost(self, J_local_host, out_host, i0=0, i1=None, j0=0, j1=None, dev_out=None):
        """End-to-end host form: this rank's J rows in (pinned) host memory -> slab-pipelined upload + projection ->
        allreduce on the device -> the combined block in `out_host` (pinned, one full copy PER RANK).  Returns out_host."""
        i1 = self.c.N if i1 is None else i1
        j1 = self.c.N if j1 is None else j1
        m2 = J_local_host.shape[1]
        if dev_out is None:
            dev_out = torch.empty((i1 - i0, j1 - j0, m2), dtype=self.c.t_dtype, device=self.c.device)
        self.c.projectMatrixHostIn(J_local_host, dev_out, i0, i1, j0, j1, q0=self.q0, q1=self.q1)
        combine(dev_out, "allreduce", self.group)
        out_host.copy_(dev_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def projectMatrixHostShared(self, J_local_host, out_shared, i0=0, i1=None, j0=0, j1=None):
        """End-to-end host form for ONE result on the host: `out_shared` is a SharedPinnedOutput mapped by every rank.
        This rank's J rows (pinned host) -> slab-pipelined upload + projection -> reduce-scatter over blocks of basis
        rows on the device -> each rank copies ONLY its row block (1/world of C) into its slice of the shared host
        buffer.  Round 1 let every rank read back the whole combined C (world x the bytes, contending for one host).
        The result is complete on every rank when this returns (barrier)."""
        i1 = self.c.N if i1 is None else i1
        j1 = self.c.N if j1 is None else j1
        m2 = J_local_host.shape[1]
        rows, per = i1 - i0, (i1 - i0 + self.world - 1) // self.world
        key = (rows, j1 - j0, m2)
        if getattr(self, "_rs_key", None) != key:
            # partial window padded to world equal row blocks (the pad rows stay zero), and my block of the total
            self._rs_part = torch.zeros((per * self.world, j1 - j0, m2), dtype=self.c.t_dtype, device=self.c.device)
            self._rs_mine = torch.empty((per, j1 - j0, m2), dtype=self.c.t_dtype, device=self.c.device)
            self._rs_key = key
        part, mine = self._rs_part, self._rs_mine
        self.c.projectMatrixHostIn(J_local_host, part[:rows], i0, i1, j0, j1, q0=self.q0, q1=self.q1)
        r0, r1 = min(rows, self.rank * per), min(rows, (self.rank + 1) * per)
        if self.world > 1:
            pr = torch.view_as_real(part) if part.is_complex() else part
            mr = torch.view_as_real(mine) if mine.is_complex() else mine
            dist.reduce_scatter_tensor(mr, pr, op=dist.ReduceOp.SUM, group=self.group)
        else:
            mine = part
        if r1 > r0:
            out_shared.tensor[r0:r1].copy_(mine[: r1 - r0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        return out_shared.tensor, (r0, r1)

    def projectVector(self, f_local, k0=0, k1=None, out=None, mode="allreduce"):
        F = self.c.projectVector(f_local, k0, k1, q0=self.q0, q1=self.q1, out=out)
        return combine(F, mode, self.group)
