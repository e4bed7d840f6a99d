import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pspace_b200 import workloads as WL
from pspace_b200.device import DeviceContainer, fill_synthetic
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
qn = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
w = WL.ALL[cfg]()
c = DeviceContainer(w.params, w.basis_type, w.complex_).initialize()
N, Q, m2 = c.N, c.Q, w.m2
qn = min(qn, Q)
wi, wj = (N, N) if w.pair_window is None else w.pair_window
width = 2 if w.complex_ else 1
J = fill_synthetic(qn * m2 * width, w.seed + 1000)
J = J.view(torch.complex128).reshape(qn, m2) if w.complex_ else J.reshape(qn, m2)
out = torch.empty((wi, wj, m2), dtype=J.dtype, device='cuda')
for r in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    c.projectMatrix(J, 0, wi, 0, wj, q0=0, q1=qn, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"cfg{cfg} q={qn} {ms:.2f} ms {wi*wj*qn*m2*(8 if w.complex_ else 2)/ms*1e-9:.2f} TF {c.lastPlan()}")
