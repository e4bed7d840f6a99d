"""One launch each of the one-pass operators (after a warm-up) for ncu:  python tools/prof_onepass.py [cfg] [force_panel]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pspace_b200 import workloads as WL
from pspace_b200.device import DeviceContainer

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
w = WL.ALL[cfg]()
c = DeviceContainer(w.params, w.basis_type, w.complex_).initialize()
c.setOption("panel_cache", 0)
c.setOption("force_panel", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
tdt = torch.complex128 if w.complex_ else torch.float64
V = torch.randn((c.N, w.m), dtype=torch.float64, device="cuda").to(tdt)
f = torch.randn((c.Q, w.m), dtype=torch.float64, device="cuda").to(tdt)
U = torch.empty((c.Q, w.m), dtype=tdt, device="cuda")
F = torch.empty((c.N, w.m), dtype=tdt, device="cuda")
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for r in range(2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    c.projectVector(f, out=F, variant=variant)
    e[1].record()
    c.evaluateExpansion(V, out=U)
    e[2].record()
    torch.cuda.synchronize()
    print(f"cfg{cfg} project_vector {e[0].elapsed_time(e[1]):.3f} ms  evaluate_expansion {e[1].elapsed_time(e[2]):.3f} ms")
