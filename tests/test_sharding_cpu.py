"""CPU tests of the multi-GPU path's host logic: shard bounds and the collective combine (gloo, world_size 2),
with the oracle supplying each rank's partial block (q-range restricted projection)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pspace_b200.sharded import row_blocks, shard_bounds


def test_shard_bounds_cover_and_align():
    for Q in (1, 31, 32, 33, 64, 1000, 3125, 1048576):
        for world in (1, 2, 3, 4, 8):
            b = shard_bounds(Q, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == Q
            for (a0, a1), (b0, b1) in zip(b[:-1], b[1:]):
                assert a1 == b0 and a0 <= a1
            for q0, q1 in b[:-1]:
                assert q1 % 32 == 0 or q1 == Q
            sizes = [q1 - q0 for q0, q1 in b]
            assert max(sizes) - min(sizes) <= 64
    assert shard_bounds(1048576, 8)[3] == (393216, 524288)
    assert row_blocks(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cplx, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import harness as H
        from pspace_b200.sharded import combine, shard_bounds as sb
        ps = [("normal", 0.3, 1.2, 3), ("uniform", -1.0, 2.0, 3), ("exponential", 0.0, 0.5, 2)]
        c = H.CpuContainer(H.oracle_api(cplx), ps, 1).initialize()
        rng = np.random.default_rng(3)
        m2 = 6
        J = rng.normal(size=(c.Q, m2)).astype(c.dtype)
        if cplx:
            J = J + 1j * rng.normal(size=J.shape)
        q0, q1 = sb(c.Q, world, align=4)[rank]
        part = c.project_matrix_qrange(J[q0:q1], q0, q1, nthreads=2)          # [N][N][m2]
        full = c.project_matrix(J, nthreads=2)
        N = c.N
        # allreduce
        t = torch.from_numpy(part.reshape(N * N, m2).copy())
        tot = combine(t, "allreduce").numpy().reshape(N, N, m2)
        err_ar = float(np.max(np.abs(tot - full)) / np.max(np.abs(full)))
        # reduce-scatter over basis-pair rows
        t = torch.from_numpy(part.reshape(N * N, m2).copy())
        mine, (r0, r1) = combine(t, "reduce_scatter")
        ref_rows = full.reshape(N * N, m2)[r0:r1]
        err_rs = float(np.max(np.abs(mine.numpy() - ref_rows)) / np.max(np.abs(full)))
        q.put((rank, err_ar, err_rs, r0, r1, N * N))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cplx", [False, True])
def test_two_rank_combine_gloo(cplx):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cplx, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] < 1e-13 and r[2] < 1e-13 for r in res), res
    # the row blocks tile [0, N*N)
    assert res[0][3] == 0 and res[0][4] == res[1][3] and res[1][4] == res[1][5]


def test_cabi_shard_rules_match_the_python_ones():
    """C++ callers shard with pspace_cuda_shard_bounds / pspace_cuda_row_block (host arithmetic, no device needed); they must
    cut the quadrature points and the basis rows exactly as pspace_b200.sharded does, so that mixed callers agree."""
    import ctypes
    from pspace_b200 import cabi
    from pspace_b200.sharded import row_blocks, shard_bounds
    L = ctypes.CDLL(cabi.LIB_PATH)
    L.pspace_cuda_shard_bounds.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    L.pspace_cuda_row_block.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    for Q in (0, 1, 31, 32, 33, 64, 1000, 3125, 65536, 1048576, 1048577):
        for world in (1, 2, 3, 4, 7, 8, 16):
            ref = shard_bounds(Q, world) if Q > 0 else [(0, 0)] * world
            for rank in range(world):
                a, b = ctypes.c_longlong(), ctypes.c_longlong()
                assert L.pspace_cuda_shard_bounds(Q, world, rank, ctypes.byref(a), ctypes.byref(b)) == 0
                assert (a.value, b.value) == ref[rank], (Q, world, rank)
    for rows in (0, 1, 5, 286, 1296):
        for world in (1, 2, 3, 8):
            ref = row_blocks(rows, world)
            for rank in range(world):
                r0, r1, per = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
                assert L.pspace_cuda_row_block(rows, world, rank, ctypes.byref(r0), ctypes.byref(r1), ctypes.byref(per)) == 0
                assert (r0.value, r1.value) == ref[rank] and per.value == (rows + world - 1) // world
    assert L.pspace_cuda_shard_bounds(10, 2, 2, None, None) != 0          # bad arguments are refused, not crashed on
