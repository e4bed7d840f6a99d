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
