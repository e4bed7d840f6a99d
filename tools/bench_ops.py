#!/usr/bin/env python3
"""Secondary operations of the path, measured on one B200 against the roofline that binds each of them — HBM bandwidth
for the integer / table / reduction kernels, the FP64 DMMA peak for the two one-pass GEMMs (a8 vector projection, f1
expansion evaluation: 2*N*m FLOP per point against 8*m bytes) — with the CPU oracle timed beside them (bounded samples).
The headline (matrix projection) is bench.py; this covers SURVEY.md §8 rows a1-a8 and f1-f3.

    python tools/bench_ops.py [cfg] > profiles/r02_ops_cfg5.jsonl       (default cfg5)

Timing: CUDA events around `inner` back-to-back calls (the GPU never waits for the host between them), best of 5 after 2
warm-ups, divided by `inner`."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import harness as H  # noqa: E402
from pspace_b200 import device as D  # noqa: E402
from pspace_b200 import workloads as WL  # noqa: E402


def gpu_time(fn, reps=5, warm=2, inner=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    w = WL.ALL[cfg]()
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json (measured copy bandwidth)"
    except Exception:
        hbm, src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    c = D.DeviceContainer(w.params, w.basis_type, w.complex_).initialize()
    c.setOption("panel_cache", 0)     # every call rebuilds its basis panel (as bench.py does): no work carried between calls
    o = H.CpuContainer(H.oracle_api(w.complex_), w.params, w.basis_type).initialize()
    N, Q, nv, m = c.N, c.Q, c.nvars, w.m
    wd = 2 if w.complex_ else 1
    tdt = torch.complex128 if w.complex_ else torch.float64
    rows = []

    fp64 = D.fp64_peak(1, 300.0, 0)

    def add(name, ms, bytes_, unit_count, unit, cpu_rate=None, cpu_note="", note="", flop=None):
        r = {"op": name, "workload": w.name, "ms": ms, "algorithmic_bytes": bytes_, "achieved_gbs": bytes_ / ms * 1e-6,
             "hbm_peak_gbs": hbm, "frac_of_hbm": bytes_ / ms * 1e-6 / hbm, "rate": unit_count / (ms * 1e-3), "unit": unit,
             "cpu_rate": cpu_rate, "cpu_note": cpu_note, "note": note, "bound": "hbm", "frac": bytes_ / ms * 1e-6 / hbm}
        if flop is not None:      # FP64-bound operator: the binding roofline is the DMMA peak measured in this run
            r.update({"bound": "fp64 (DMMA)", "useful_flop": flop, "achieved_tflops": flop / ms * 1e-9, "fp64_peak_tflops": fp64,
                      "frac": flop / ms * 1e-9 / fp64})
        rows.append(r)

    # (a1) multi-index construction
    pmax = np.array([p[3] for p in w.params], dtype=np.intc)
    ntup = float(np.prod(pmax + 1.0))
    ms = gpu_time(lambda: c.initializeBasis(pmax), reps=3, warm=1)
    t = time.perf_counter()
    o.initialize_basis(pmax)
    dt = time.perf_counter() - t
    add("basis_degrees (a1)", ms, 4.0 * nv * N, ntup, "tuples enumerated/s", ntup / dt, "oracle enumeration, 1 thread",
        "includes cudaMalloc + the D2H mirror of the table; latency bound (3 launches + sync)")
    # (a2,a3) rules + tables + digit/weight grid
    nq = pmax + 1
    ms = gpu_time(lambda: c.initializeQuadrature(nq), reps=3, warm=1)
    add("initialize_quadrature (a2,a3)", ms, Q * (16.0 if nv <= 16 else 32.0) + 8.0 * wd * Q, Q, "grid points/s",
        note="1-D rules, 1-D tables, digit table + weights; includes cudaMalloc/cudaFree of the grid arrays")
    # (a3) materialised grid
    qn = min(Q, 1 << 20)
    ms = gpu_time(lambda: c.grid(0, qn))
    add("tensor_grid Z,Y,W (a3)", ms, 8.0 * wd * (2 * nv + 1) * qn, qn, "grid points/s", note="includes 3 torch allocations")
    # (a5) batched basis at grid points
    kq = min(Q, 1 << 18)
    ms = gpu_time(lambda: c.evalBasisGrid(0, N, 0, kq), inner=4)
    add("eval_basis_grid (a5)", ms, 8.0 * wd * N * kq, float(N) * kq, "basis values/s", note=f"{kq} points, all {N} basis rows")
    # (f1) expansion evaluation and (a8) vector projection: one-pass GEMMs with a generated operand; algorithmic bytes are
    # the m columns read / written per point, the binding roofline is FP64 (2*N*m FLOP per point, x4 complex)
    V = torch.randn((N, m), dtype=torch.float64, device="cuda").to(tdt)
    U = torch.empty((Q, m), dtype=tdt, device="cuda")
    F = torch.empty((N, m), dtype=tdt, device="cuda")
    flop = 2.0 * N * Q * m * (4 if w.complex_ else 1)
    qs = min(Q, 4096)
    Vh = V.cpu().numpy()
    t = time.perf_counter()
    o.evaluate_expansion(Vh, 0, N, 0, qs)
    dt = time.perf_counter() - t
    for force_panel in (0, 1):
        if force_panel and 8.0 * wd * N * Q > 16e9:
            continue
        c.setOption("force_panel", force_panel)
        tag = " [round-1 panel kernels]" if force_panel else ""
        ms = gpu_time(lambda: c.evaluateExpansion(V, out=U), inner=10)
        add("evaluate_expansion (f1)" + tag, ms, 8.0 * wd * (m * Q + N * m), float(N) * Q, "(k,q) terms/s", float(N) * qs / dt,
            f"oracle, 1 thread, first {qs} points", c.lastPlan(), flop=flop)
        f = U
        ms = gpu_time(lambda: c.projectVector(f, out=F), inner=10)
        add("project_vector (a8)" + tag, ms, 8.0 * wd * (m * Q + N * m), float(N) * Q, "(k,q) updates/s", note=c.lastPlan(), flop=flop)
    c.setOption("force_panel", 0)
    # (f3) moments
    ms = gpu_time(lambda: c.moments(f), inner=10)
    fh = f[:min(Q, 1 << 16)].cpu().numpy()
    t = time.perf_counter()
    o.moments(fh, 0, fh.shape[0])
    dt = time.perf_counter() - t
    add("moments (f3)", ms, 8.0 * wd * m * Q + 8.0 * wd * Q, float(Q), "points/s", fh.shape[0] / dt,
        f"oracle, 1 thread, first {fh.shape[0]} points")
    # (f2) TACS scatter of the full coefficient matrix (cfg5: 286^2 blocks of 24x24 = 377 MB)
    if N * N * m * m * 8 * wd < 4e9:
        C = torch.randn((N, N, m * m), dtype=torch.float64, device="cuda").to(tdt)
        mat = torch.zeros((N * m, N * m), dtype=tdt, device="cuda")
        nn_, nd_ = (4, m // 4) if m % 4 == 0 else (1, m)
        ms = gpu_time(lambda: D.scatter_matrix_tacs(C, nn_, nd_, mat))
        add("scatter_matrix_tacs (f2)", ms, 8.0 * wd * 3 * N * N * m * m, float(N) * N * m * m, "elements/s",
            note="read C, read-modify-write mat")
        mk = c.pairMask([1] * nv)
        ms = gpu_time(lambda: c.pairMask([1] * nv))
        add("pair_mask (f2)", ms, float(N) * N, float(N) * N, "pairs/s",
            note=f"kept {int(mk.sum())} of {N * N} pairs for dmapf=1; includes a stream sync + malloc")
    for r in rows:
        print(json.dumps(r))
    sys.stderr.write(f"HBM peak source: {src}\n")


if __name__ == "__main__":
    main()
