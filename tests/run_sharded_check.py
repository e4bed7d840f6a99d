#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/run_sharded_check.py

Every rank projects its quadrature-point shard of cfg2 (real) and of a complex 4-parameter container with the CUDA
path; the partial blocks are combined with NCCL (allreduce and reduce-scatter) and compared on rank 0 with the CPU
oracle's full projection.  Also runs the host-input path (slab upload) per rank.  Prints one JSON line per case.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import harness as H  # noqa: E402
from pspace_b200 import workloads as WL  # noqa: E402
from pspace_b200.device import DeviceContainer  # noqa: E402
from pspace_b200.sharded import ShardedProjector  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("cfg2_real", WL.cfg2().params, 1, False, 64),
             ("complex_4param", [("normal", 0.3 + 1e-30j, 1.2, 3), ("exponential", 0.0, 0.5, 4), ("uniform", -1.0, 2.0, 5),
                                 ("normal", -1.0, 0.4, 3)], 0, True, 36)]
    ok = True
    for name, ps, bt, cplx, m2 in cases:
        c = DeviceContainer(ps, bt, cplx, device=local).initialize()
        sp = ShardedProjector(c)
        o = H.CpuContainer(H.oracle_api(cplx), ps, bt).initialize()
        rng = np.random.default_rng(5)
        J = rng.normal(size=(o.Q, m2)).astype(o.dtype)
        if cplx:
            J = J + 1j * rng.normal(size=J.shape)
        Jl = torch.from_numpy(np.ascontiguousarray(J[sp.q0:sp.q1])).cuda()
        C = sp.projectMatrix(Jl, 0, 40, 0, o.N)                                   # allreduce
        rows, (r0, r1) = sp.projectMatrix(Jl, 0, 40, 0, o.N, mode="reduce_scatter")
        Ch = torch.empty((40, o.N, m2), dtype=C.dtype).pin_memory()
        sp.projectMatrixHost(torch.from_numpy(np.ascontiguousarray(J[sp.q0:sp.q1])).pin_memory(), Ch, 0, 40, 0, o.N)
        F = sp.projectVector(Jl)
        # the C-ABI collective entry point (what a C++ caller uses) on torch's own NCCL communicator, when torch exposes it
        e_cabi = 0.0
        comm_ptr = None
        try:
            comm_ptr = dist.distributed_c10d._get_default_group()._get_backend(torch.device("cuda"))._comm_ptr()
        except Exception:
            comm_ptr = None
        Cc = None
        if comm_ptr:
            import ctypes
            from pspace_b200.cabi import check
            a = c._args(0, 40, 0, o.N, sp.q0, sp.q1, m2, False, 0, 0)
            ws, need = c._workspace(1, a)
            Cc = torch.empty((40, o.N, m2), dtype=C.dtype, device="cuda")
            check(c.L.pspace_cuda_project_matrix_allreduce(c.h, ctypes.byref(a), Jl.data_ptr(), Cc.data_ptr(), ws.data_ptr() if ws is not None else None,
                                                           need, ctypes.c_void_p(comm_ptr), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
        from pspace_b200.fused import FusedShardedProjector
        fp = FusedShardedProjector(c, 0, 40, 0, o.N, ncols=m2)
        Cf = fp.projectMatrix(Jl).clone()
        Cf2 = fp.projectMatrix(Jl).clone()                                       # slots reused: second call must agree bitwise
        ref = o.project_matrix(J, 0, 40, 0, o.N, nthreads=8)
        refF = o.project_vector(J)
        den = np.max(np.abs(ref))
        e_ar = float(np.max(np.abs(C.cpu().numpy() - ref)) / den)
        e_rs = float(np.max(np.abs(rows.cpu().numpy() - ref[r0:r1])) / den) if r1 > r0 else 0.0
        e_h = float(np.max(np.abs(Ch.numpy() - ref)) / den)
        e_f = float(np.max(np.abs(F.cpu().numpy() - refF)) / np.max(np.abs(refF)))
        e_fused = float(np.max(np.abs(Cf.cpu().numpy() - ref)) / den)
        if not torch.equal(Cf, Cf2):
            e_fused = 1.0
        if Cc is not None:
            e_cabi = float(np.max(np.abs(Cc.cpu().numpy() - ref)) / den)
        errs = torch.tensor([e_ar, e_rs, e_h, e_f, e_fused, e_cabi], dtype=torch.float64, device="cuda")
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        if rank == 0:
            e = [float(x) for x in errs]
            good = max(e) < 1e-12
            ok = ok and good
            print(json.dumps({"case": name, "world": world, "N": o.N, "Q": o.Q, "m2": m2, "allreduce_err": e[0], "reduce_scatter_err": e[1],
                              "host_input_err": e[2], "vector_err": e[3], "fused_peer_err": e[4], "cabi_allreduce_err": e[5],
                              "cabi_on_torch_comm": bool(comm_ptr), "shards": [sp.q0, sp.q1], "pass": good}))
        c.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
