#!/usr/bin/env python3
"""One-pass operators (a8 vector projection, f1 expansion evaluation) on one B200: factored-operand kernels vs the
round-1 panel kernels, against the FP64 DMMA peak (FLOP-bound: 2*N*m FLOP per point, x4 complex).

    python tools/bench_onepass.py [cfg ...]      one JSON line per (cfg, op, path)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pspace_b200 import cabi  # noqa: E402
from pspace_b200 import device as D  # noqa: E402
from pspace_b200 import workloads as WL  # noqa: E402


def gpu_time(fn, reps=7, warm=3, inner=10):
    """CUDA events around `inner` back-to-back calls (the GPU never idles waiting for the host), per call."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    cfgs = [int(a) for a in sys.argv[1:]] or [5]
    peak = D.fp64_peak(1, 300.0, 0)
    for cfg in cfgs:
        w = WL.ALL[cfg]()
        c = D.DeviceContainer(w.params, w.basis_type, w.complex_).initialize()
        c.setOption("panel_cache", 0)
        N, Q, m = c.N, c.Q, w.m
        tdt = torch.complex128 if w.complex_ else torch.float64
        V = torch.randn((N, m), dtype=torch.float64, device="cuda").to(tdt)
        f = torch.randn((Q, m), dtype=torch.float64, device="cuda").to(tdt)
        U = torch.empty((Q, m), dtype=tdt, device="cuda")
        F = torch.empty((N, m), dtype=tdt, device="cuda")
        flop = 2.0 * N * Q * m * (4 if w.complex_ else 1)
        abytes = 8.0 * (2 if w.complex_ else 1) * (Q * m + N * m)
        for path in (0, 1):
            if path == 1 and 8.0 * (2 if w.complex_ else 1) * N * Q > 20e9:
                continue                    # the panel would not fit comfortably
            c.setOption("force_panel", path)
            for op, fn in (("project_vector (a8)", lambda: c.projectVector(f, out=F)),
                           ("evaluate_expansion (f1)", lambda: c.evaluateExpansion(V, out=U))):
                best, med = gpu_time(fn)
                print(json.dumps({"workload": w.name, "op": op, "path": "panel (round 1)" if path else "factored operand",
                                  "N": N, "Q": Q, "m": m, "ms_best": round(best, 4), "ms_median": round(med, 4),
                                  "useful_flop": flop, "tflops": round(flop / best * 1e-9, 3), "fp64_dmma_peak_tflops": round(peak, 2),
                                  "frac_of_fp64_peak": round(flop / best * 1e-9 / peak, 4), "algorithmic_bytes": abytes,
                                  "plan": c.lastPlan()}), flush=True)
        c.close()


if __name__ == "__main__":
    main()
