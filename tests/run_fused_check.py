#!/usr/bin/env python3
"""N-GPU check of the two exchange paths at cfg5 size (run under torchrun): every rank's partial block is gathered and
summed in rank order on the device (the exact definition of the result), and the NCCL-allreduce result and the fused
peer-memory result are compared with it.  Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pspace_b200 import workloads as WL  # noqa: E402
from pspace_b200.device import DeviceContainer, fill_synthetic  # noqa: E402
from pspace_b200.fused import FusedShardedProjector  # noqa: E402
from pspace_b200.sharded import shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    w = WL.cfg5()
    qfrac = int(sys.argv[1]) if len(sys.argv) > 1 else 4            # use Q/qfrac points to keep the run short
    c = DeviceContainer(w.params, w.basis_type, False, device=local).initialize()
    N, Q, m2 = c.N, c.Q, w.m2
    q0, q1 = shard_bounds(Q, world)[rank]
    q1 = q0 + (q1 - q0) // qfrac
    J = fill_synthetic((q1 - q0) * m2, w.seed + 1000, offset=q0 * m2, device=local).reshape(q1 - q0, m2)
    part = c.projectMatrix(J, q0=q0, q1=q1)
    # exact: gather every rank's partial and add in rank order
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    ref = parts[0].clone()
    for r in range(1, world):
        ref += parts[r]
    del parts
    nccl = part.clone()
    dist.all_reduce(nccl, op=dist.ReduceOp.SUM)
    fp = FusedShardedProjector(c, 0, N, 0, N, ncols=m2)
    fp.q0, fp.q1 = q0, q1
    fp.args.q0, fp.args.q1 = q0, q1
    fused = fp.projectMatrix(J).clone()
    fused2 = fp.projectMatrix(J).clone()
    den = float(ref.abs().max())
    if os.environ.get("FUSED_DEBUG"):
        bad = (fused - ref).abs() > 1e-13 * den
        idx = bad.nonzero()
        info = {"rank": rank, "nbad": int(bad.sum())}
        if idx.shape[0]:
            ti, tj, tn = idx[:, 0] // 16, idx[:, 1] // 8, idx[:, 2] // 96
            tiles = torch.unique(torch.stack([ti, tj, tn], 1), dim=0)
            tid = (tiles[:, 0] * 36 + tiles[:, 1]) * 6 + tiles[:, 2]
            info.update(ntiles=int(tiles.shape[0]), tile_ids=sorted(int(x) for x in tid)[:24], owners=sorted(set(int(x) for x in (tiles[:, 0] * 36 + tiles[:, 1]) // 81)),
                        maxdiff=float((fused - ref).abs().max()) / den,
                        remainder=sorted(set(bool(x >= 26 * 148) for x in tid)))
            # does the bad value look like a missing / doubled partial of some rank?
            i, j, cc = (int(v) for v in idx[0])
            info["sample"] = {"ijc": [i, j, cc], "fused": float(fused[i, j, cc]), "ref": float(ref[i, j, cc]), "mine": float(part[i, j, cc])}
        print(json.dumps(info), flush=True)
    e = torch.tensor([float((nccl - ref).abs().max()) / den, float((fused - ref).abs().max()) / den,
                      0.0 if torch.equal(fused, fused2) else 1.0, float(part.abs().max()) / den], dtype=torch.float64, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "points_per_rank": q1 - q0, "nccl_vs_rank_order_sum": float(e[0]), "fused_vs_rank_order_sum": float(e[1]),
                          "fused_repeatable": float(e[2]) == 0.0, "max_partial_over_max_total": float(e[3])}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
