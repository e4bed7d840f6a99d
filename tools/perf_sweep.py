import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pspace_b200 import workloads as WL
from pspace_b200.device import DeviceContainer, fill_synthetic, fp64_peak
cfgs = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [5]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
print("peak dmma TF", fp64_peak(1, 300.0), "dfma", fp64_peak(0, 300.0))
for cfg in cfgs:
    w = WL.ALL[cfg]()
    c = DeviceContainer(w.params, w.basis_type, w.complex_).initialize()
    N, Q, m2 = c.N, c.Q, w.m2
    wi, wj = (N, N) if w.pair_window is None else w.pair_window
    width = 2 if w.complex_ else 1
    J = fill_synthetic(Q * m2 * width, w.seed + 1000)
    J = J.view(torch.complex128).reshape(Q, m2) if w.complex_ else J.reshape(Q, m2)
    out = torch.empty((wi, wj, m2), dtype=J.dtype, device='cuda')
    torch.cuda.synchronize()
    for r in range(reps + 1):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        c.projectMatrix(J, 0, wi, 0, wj, out=out, variant=variant)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        upd = wi * wj * Q
        flop = upd * m2 * (8 if w.complex_ else 2)
        print(f"cfg{cfg} N={N} Q={Q} m2={m2} win={wi}x{wj} {ms:.2f} ms  {upd/ms*1e3:.3e} upd/s  {flop/ms*1e-9:.2f} TF  plan: {c.lastPlan()}")
    del J, out
    c.close()
