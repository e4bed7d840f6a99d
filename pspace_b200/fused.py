"""Fused multi-GPU projection: compute + reduce-scatter + all-gather over peer-mapped memory (NVLink / NVSwitch).

The NCCL path (sharded.ShardedProjector) keeps a full partial C per rank and calls one allreduce after the kernel.
Here the exchange is part of the kernels: the projection kernel's epilogue pushes every finished partial tile straight
into the receive slot of the rank that owns it (peer stores over NVLink, overlapped with the math of the following
tiles); after a barrier, a second kernel adds the `world` partial blocks of each owned element in rank order and stores
the total into the C window of every rank.  Buffers are torch.distributed._symmetric_memory allocations (peer-mapped
into every process of the node); the two cross-rank barriers are the symmetric-memory handle's device-side barriers.
Results are bit-reproducible (fixed summation order) and equal to the NCCL path up to the order of the `world` adds.
"""
import ctypes

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

from .cabi import ProjectArgs, check
from .sharded import shard_bounds


class FusedShardedProjector:
    def __init__(self, container, i0=0, i1=None, j0=0, j1=None, ncols=None, group=None):
        assert dist.is_initialized(), "FusedShardedProjector needs torch.distributed (one process per GPU)"
        self.c = container
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.i0, self.i1 = i0, (container.N if i1 is None else i1)
        self.j0, self.j1 = j0, (container.N if j1 is None else j1)
        self.ncols = ncols
        bounds = shard_bounds(container.Q, self.world)
        # every rank sees every rank's range (no communication needed): fail on ALL ranks together instead of letting
        # the ranks with points block in the barrier while an empty one raises
        empty = [r for r, (a, b) in enumerate(bounds) if b <= a]
        if empty:
            raise ValueError(f"fused projection: {container.Q} quadrature points leave rank(s) {empty} of {self.world} without a "
                             "shard; use fewer ranks or the NCCL path (ShardedProjector handles empty shards)")
        self.q0, self.q1 = bounds[self.rank]
        self.args = ProjectArgs(self.i0, self.i1, self.j0, self.j1, self.q0, self.q1, ncols, 0, 0, 0, None, 0)
        L = container.L
        n = ctypes.c_longlong()
        check(L.pspace_cuda_peer_layout(container.h, ctypes.byref(self.args), self.world, ctypes.byref(n)))
        dev = container.device
        width = 2 if container.is_complex else 1
        # receive slots [world][owned tiles][BM][ncols_pad] and the output window, both peer-mapped
        self.recv = symm_mem.empty(n.value, dtype=torch.float64, device=dev)
        self.out_flat = symm_mem.empty((self.i1 - self.i0) * (self.j1 - self.j0) * ncols * width, dtype=torch.float64, device=dev)
        self.h_recv = symm_mem.rendezvous(self.recv, self.group)
        self.h_out = symm_mem.rendezvous(self.out_flat, self.group)
        PtrArr = ctypes.c_void_p * self.world
        self.recv_ptrs = PtrArr(*[int(p) for p in self.h_recv.buffer_ptrs])
        self.out_ptrs = PtrArr(*[int(p) for p in self.h_out.buffer_ptrs])
        ws = L.pspace_cuda_project_workspace(container.h, 1, ctypes.byref(self.args))
        self.ws = torch.empty(max(ws, 16), dtype=torch.uint8, device=dev)
        self.ws_bytes = ws
        shape = (self.i1 - self.i0, self.j1 - self.j0, ncols)
        self.out = (torch.view_as_complex(self.out_flat.view(*shape, 2)) if container.is_complex else self.out_flat.view(*shape))

    @property
    def local_points(self):
        return self.q1 - self.q0

    def projectMatrix(self, J_local, ev_start=None, ev_stop=None):
        """J_local: device rows [q0,q1) of J.  Returns this rank's (peer-mapped) window holding the COMBINED block.
        ev_start / ev_stop (torch.cuda.Event) are recorded around this rank's projection kernels (bench roofline)."""
        c, L = self.c, self.c.L
        assert J_local.is_cuda and J_local.is_contiguous() and J_local.shape == (self.q1 - self.q0, self.ncols)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        with torch.cuda.device(c.device):
            if ev_start is not None:
                ev_start.record()
            check(L.pspace_cuda_project_matrix_peer(c.h, ctypes.byref(self.args), J_local.data_ptr(), self.recv_ptrs, self.world,
                                                    self.rank, self.ws.data_ptr(), self.ws_bytes, st))
            if ev_stop is not None:
                ev_stop.record()
            self.h_recv.barrier(channel=0)          # every rank's pushes have landed
            check(L.pspace_cuda_peer_reduce_gather(c.h, ctypes.byref(self.args), self.recv.data_ptr(), self.out_ptrs, self.world,
                                                   self.rank, st))
            self.h_out.barrier(channel=1)           # every rank's window is complete (and the slots may be reused)
        return self.out
