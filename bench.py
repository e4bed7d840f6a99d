#!/usr/bin/env python3
"""bench.py — Galerkin projection throughput (BASELINE.json metric) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over the whole workload: the matrix (Jacobian) projection
C[i][j][:] = sum_q w_q psi_i(z_q) psi_j(z_q) J[q][:] of BASELINE.json configs[4] (cfg5: 10 mixed parameters,
complete degree-3 basis N = 286, 4^10 = 1 048 576 quadrature points, 24x24 blocks) — all N^2 ordered pairs, all
q: 8.58e10 updates of 576 FMAs.  At N GPUs the quadrature points are sharded (strong scaling: the problem is
fixed) and the partial coefficient blocks are combined by one NCCL allreduce inside the timed step.

Other workloads (`--workload cfg1..cfg4`, BASELINE.json configs[0..3]) run the same step on their own shapes: cfg3 projects
E' = 8 elements that share one container per step through pspace_cuda_project_matrix_batched (full 1296 x 1296 basis window,
62 GB of output; at N GPUs the ELEMENTS are split across ranks — independent objects, no collective); cfg4 is the complex
(USE_COMPLEX) build on a 64 x 64 basis window (the full N^2 output would be 39.6 PB) plus its full vector projection.
Both arms print the SAME `config` dict (a pure function of workload, N and the collective).

JSON keys (one line, rank 0):
  value        updates/s, whole job, J resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e          same metric through the host-buffer call (pinned host J in, host C out; H2D/D2H inside the timing)
  roofline     dominant kernel (k_project_ws, FP64 tensor pipe / DMMA): achieved = useful FLOP of one launch /
               its CUDA-event duration; peak = FP64 DMMA peak MEASURED IN THIS RUN (MEASURED_PEAKS.json has no
               FP64 figure) — "tensor" bound because DMMA.8x8x4 issues to the tensor pipe's FP64 sub-pipe
  cpu_baseline the CPU implementation of the same path on this box's host cores (bounded sample)
  gpu_launches kernels of this library launched inside the timed region
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Galerkin projection (quad-pt x basis-pair) updates/s"
UNIT = "updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sym", action="store_true")
    ap.add_argument("--collective", default="nccl", choices=["nccl", "peer"],
                    help="N > 1: how the partial coefficient blocks are combined in the timed step: one NCCL allreduce after the kernel "
                         "(default: measured faster), or the fused peer-memory exchange (pushes from the projection epilogue); the other "
                         "one is timed beside it")
    ap.add_argument("--ref-seconds", type=float, default=0.0, help="--impl reference: seconds per step (0 = derive from --steps)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of the CPU baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload(name):
    from pspace_b200 import workloads as WL
    return WL.ALL[int(name[3:])]()


BENCH_ELEMENTS = 8          # cfg3: elements per step (E' chosen so that E' N^2 m^2 8 B fits HBM; 8 x 7.7 GB = 62 GB)


def shape_of(w):
    """(rows i, cols j) of the basis window and elements per step of a workload."""
    N = w.N
    wi, wj = (N, N) if w.pair_window is None else (min(N, w.pair_window[0]), min(N, w.pair_window[1]))
    return wi, wj, (BENCH_ELEMENTS if w.elements > 1 else 1)


def config_for(w, world, collective):
    """`config` of the JSON line: identical in the b200 and the reference arm (workload, N GPUs, collective only)."""
    wi, wj, E = shape_of(w)
    width = 2 if w.complex_ else 1
    updates = float(E) * wi * wj * w.Q
    if E > 1:
        par = f"{E} elements per step split over {world} rank(s): independent elements, no collective" if world > 1 else "1 GPU"
    elif world > 1:
        par = f"q-shard x{world} + " + ("fused peer-memory reduce-scatter/all-gather of C (NVLink pushes from the projection epilogue)"
                                         if collective == "peer" else "NCCL allreduce of C")
    else:
        par = "1 GPU"
    return {"workload": f"{w.name}: matrix (Jacobian) projection, all ordered pairs of a {wi}x{wj} basis window x all q"
                        + (f", {E} elements sharing one container per step (of the nominal {w.elements}; rate per element is size-invariant)" if E > 1 else ""),
            "N": w.N, "Q": w.Q, "m": w.m, "params": w.nvars, "basis": "complete" if w.basis_type else "tensor",
            "window": [wi, wj], "elements_per_step": E, "updates_per_step": updates,
            "flop_per_step": updates * w.m2 * (8 if w.complex_ else 2), "parallelism": par,
            "l2": (f"inputs larger than L2: J {E * w.Q * w.m2 * width * 8 / 1e9:.3f} GB + C {E * wi * wj * w.m2 * width * 8 / 1e9:.3f} GB per step "
                   "vs 126 MB L2 (no flush needed)" if E * (w.Q + wi * wj) * w.m2 * width * 8 > 2.5e8 else
                   f"working set {E * (w.Q + wi * wj) * w.m2 * width * 8 / 1e6:.2f} MB fits L2: a 256 MB buffer is written between timed steps to flush it "
                   "(the flush is outside the timing)")}


def cpu_sample(w, seconds, api_kind="auto"):
    """Time the CPU implementation of the matrix projection on a bounded pair window x q-range with all host
    threads.  Uses the compiled reference (oracle/_ref, kind "reference") when it exists AND the workload has
    <= 5 parameters (the reference stops at 5: BasisHelper.cpp:63-65); otherwise the restated oracle ("port")."""
    import numpy as np
    from oracle import harness as H
    cores = os.cpu_count() or 1
    use_ref = H.ref_available(w.complex_) and w.nvars <= 5
    api = H.ref_api(w.complex_) if use_ref else H.oracle_api(w.complex_)
    c = H.CpuContainer(api, w.params, w.basis_type).initialize()
    N, Q, m2 = c.N, c.Q, w.m2
    width = 2 if w.complex_ else 1

    def make_J(qn):
        J = H.synthetic(qn * m2 * width, w.seed + 1000)
        return J.view(np.complex128).reshape(qn, m2) if w.complex_ else J.reshape(qn, m2)

    def run(rows, cols, qn):
        """rows i < rows, cols j < cols, quadrature points q < qn (the compiled reference always sweeps all Q)."""
        J = make_J(Q if use_ref else qn)
        t = time.perf_counter()
        if use_ref:
            c.project_matrix(J, 0, rows, 0, cols, nthreads=cores)
            done = rows * cols * Q
        else:
            c.project_matrix_qrange(J, 0, qn, 0, rows, 0, cols, nthreads=cores)
            done = rows * cols * qn
        return done, time.perf_counter() - t

    rows = min(N, cores)
    done, dt = run(rows, min(N, 8), min(Q, 1024))          # calibration
    want = done / dt * seconds                               # updates that fit the time budget
    cols = N
    if use_ref:
        rows = int(max(1, min(N, want / (cols * Q))))
        qn = Q
    else:
        qn = int(min(Q, max(1024, want / (rows * cols))))
    done, dt = run(rows, cols, qn)
    qdesc = f"all {Q}" if qn == Q else f"first {qn}"
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"rows i<{rows} x all {cols} j x {qdesc} quadrature points of {w.name}: {done:.3e} updates in {dt:.1f} s, "
                      f"{cores} threads (row-partitioned), (i,j,q) loop of TACSStochasticElement.cpp:380-414 over per-call "
                      f"basis()/quadrature()"}, (rows, cols, qn)


# ------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = workload(args.workload)
    per_step = args.ref_seconds if args.ref_seconds > 0 else max(5.0, min(30.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(w, min(per_step, 5.0))
    vals, info = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        info, _ = cpu_sample(w, per_step)
        vals.append(info["value"])
    dt = time.perf_counter() - t0
    v = statistics.mean(vals)
    info["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if not w.complex_ else "c128", "data": "synthetic",
            "config": config_for(w, max(1, args.gpus), args.collective),
            "note": "CPU arm: each step times a bounded window of the workload; updates/s is window-invariant",
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def file_sha(path):
    import hashlib
    try:
        return hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]
    except OSError:
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pspace_b200.device import DeviceContainer, fill_synthetic, fp64_peak
    from pspace_b200.sharded import ShardedProjector, SharedPinnedOutput, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the pspace B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = workload(args.workload)
    c = DeviceContainer(w.params, w.basis_type, w.complex_, device=local).initialize()
    c.setOption("panel_cache", 0)        # every step rebuilds the basis panels: no work is carried between steps
    N, Q, m2 = c.N, c.Q, w.m2
    wi, wj, E = shape_of(w)
    assert N == w.N and Q == w.Q
    width = 2 if w.complex_ else 1
    tdt = torch.complex128 if w.complex_ else torch.float64
    batched = E > 1
    if batched:
        # independent elements: split over the ranks, every rank sweeps all quadrature points, no exchange
        e_lo, e_hi = E * rank // world, E * (rank + 1) // world
        q0, q1 = 0, Q
    else:
        e_lo, e_hi = 0, 1
        q0, q1 = shard_bounds(Q, world)[rank]
    nE, nq = e_hi - e_lo, q1 - q0
    cfg = config_for(w, world, args.collective)
    small = E * (Q + wi * wj) * m2 * width * 8 <= 2.5e8            # working set within L2: flush between timed steps
    flush_buf = torch.empty(1 << 25, dtype=torch.float64, device=dev) if small else None

    def flush():
        if flush_buf is not None:
            flush_buf.fill_(1.0)

    # synthetic J rows of this rank's share (same global stream of numbers on any sharding)
    nj_el = nq * m2 * width
    Jflat = fill_synthetic(max(1, nE) * nj_el, w.seed + 1000, offset=(e_lo * Q + q0) * m2 * width, device=local)
    if batched:
        J = Jflat.reshape(max(1, nE), nq, m2)
        C = torch.empty((max(1, nE), wi, wj, m2), dtype=tdt, device=dev)
    else:
        J = Jflat.view(torch.complex128).reshape(nq, m2) if w.complex_ else Jflat.reshape(nq, m2)
        C = torch.empty((wi, wj, m2), dtype=tdt, device=dev)
    Cr = torch.view_as_real(C) if w.complex_ else C

    # N > 1 (q-sharded workloads): the fused peer-memory exchange is timed beside the NCCL allreduce
    fp, fused_note = None, None
    if world > 1 and not batched:
        try:
            from pspace_b200.fused import FusedShardedProjector
            fp = FusedShardedProjector(c, 0, wi, 0, wj, ncols=m2)
        except Exception as e:                      # no peer access / symmetric memory unavailable: NCCL path only
            fp, fused_note = None, f"fused peer path unavailable ({type(e).__name__}: {e}); NCCL allreduce used"
        ok = torch.tensor([1 if fp is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            fp = None
    use_peer = fp is not None and args.collective == "peer"

    def project():
        if batched:
            if nE > 0:
                c.projectMatrixBatched(J[:nE], 0, wi, 0, wj, out=C[:nE])
        else:
            c.projectMatrix(J, 0, wi, 0, wj, q0=q0, q1=q1, out=C)

    def nccl_step():
        project()
        if world > 1 and not batched:
            dist.all_reduce(Cr, op=dist.ReduceOp.SUM)

    def step():
        if use_peer:
            fp.projectMatrix(J)
        else:
            nccl_step()

    updates = cfg["updates_per_step"]                              # all ordered pairs of the window x all q x elements
    flop_per_update = m2 * (8 if w.complex_ else 2)

    # FP64 peak of this GPU, measured now: register-resident DMMA and DFMA loops and cuBLAS DGEMM 8192^3 (SURVEY.md §8d:
    # "use the larger")
    peak_dmma = fp64_peak(1, 400.0, local)
    peak_dfma = fp64_peak(0, 400.0, local)
    peak_dgemm = None
    try:
        n = 8192
        A_ = torch.randn((n, n), dtype=torch.float64, device=dev)
        B_ = torch.randn((n, n), dtype=torch.float64, device=dev)
        torch.matmul(A_, B_)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            g0.record()
            torch.matmul(A_, B_)
            g1.record()
            torch.cuda.synchronize()
            best = min(best, g0.elapsed_time(g1))
        peak_dgemm = 2.0 * n ** 3 / (best * 1e-3) * 1e-12
        del A_, B_
        torch.cuda.empty_cache()
    except Exception:
        peak_dgemm = None
    peak = max(peak_dmma, peak_dfma, peak_dgemm or 0.0)

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    launches0 = c.launchCount()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4 * args.steps + 2)]
    barrier()
    total_ms = 0.0
    for s in range(args.steps):
        flush()
        t0e, t1e = ev[4 * s], ev[4 * s + 1]            # the step
        a, b = ev[4 * s + 2], ev[4 * s + 3]            # this rank's projection kernels only (roofline)
        t0e.record()
        if use_peer:
            fp.projectMatrix(J, a, b)
        else:
            a.record()
            project()
            b.record()
            if world > 1 and not batched:
                dist.all_reduce(Cr, op=dist.ReduceOp.SUM)
        t1e.record()
        if small:
            torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if small:
        total_ms = sum(ev[4 * s].elapsed_time(ev[4 * s + 1]) for s in range(args.steps))     # flushes are outside the timing
    else:
        total_ms = ev[0].elapsed_time(ev[4 * (args.steps - 1) + 1])                             # back-to-back steps
    launches = c.launchCount() - launches0
    plan = c.lastPlan()
    total_ms = maxr(total_ms)
    ms_per_step = total_ms / args.steps
    value = updates / (ms_per_step * 1e-3)
    kern_ms = [ev[4 * s + 2].elapsed_time(ev[4 * s + 3]) for s in range(args.steps)]

    # N > 1: the same step with the other exchange, timed beside
    other = None
    if world > 1 and fp is not None:
        alt = nccl_step if use_peer else (lambda: fp.projectMatrix(J))
        alt()
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(args.steps):
            alt()
        n1.record()
        barrier()
        nt = maxr(n0.elapsed_time(n1) / args.steps)
        nccl_step()
        fp.projectMatrix(J)
        diff = float((fp.out - C).abs().max() / C.abs().max())
        other = {"collective": "nccl allreduce after the kernel" if use_peer else
                               "fused: partial tiles pushed to their owner rank over NVLink from the projection epilogue (peer-mapped "
                               "symmetric memory), then k_peer_reduce_gather (pspace_b200/fused.py)",
                 "ms_per_step": nt, "value": updates / (nt * 1e-3), "unit": UNIT, "max_rel_diff_between_the_two": diff}

    # ---- vector (residual-type) projection of the same container, m = block edge (SURVEY.md §8d asks for it beside) ------
    J2d = J[0] if batched else J
    if batched and w.complex_:
        J2d = J2d.view(torch.complex128)
    fq0, fq1 = (0, Q) if batched else (q0, q1)
    fvec = J2d[:, : w.m].contiguous()
    Fv = torch.empty((N, w.m), dtype=tdt, device=dev)

    def vec_step():
        c.projectVector(fvec, q0=fq0, q1=fq1, out=Fv)
        if world > 1 and not batched:
            dist.all_reduce(torch.view_as_real(Fv) if w.complex_ else Fv, op=dist.ReduceOp.SUM)
    vec_step()
    barrier()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record()
    for _ in range(10):
        vec_step()
    v1.record()
    barrier()
    vt = maxr(v0.elapsed_time(v1) / 10)
    vflop = 2.0 * N * Q * w.m * (4 if w.complex_ else 1)
    vector = {"value": float(N) * Q / (vt * 1e-3), "unit": "(k,q) updates/s of m FMAs", "ms_per_step": vt, "m": w.m,
              "frac_of_fp64_peak": (vflop / (vt * 1e-3) * 1e-12 / peak) if (world == 1 or batched) else None,
              "plan": c.lastPlan(),
              "note": "F[k][:] = sum_q w_q psi_k(z_q) f[q][:], all N basis rows, all q; f = first m columns of J; 10 back-to-back calls"
                      + (" (each rank the whole vector projection: elements are what is sharded here)" if batched and world > 1 else "")}

    # ---- live parity of this run's kernels against the CPU oracle: a 2x2 pair window over ALL of rank 0's points ----
    parity = None
    ia, ja = min(N // 2, max(0, wi - 2)), max(0, wj - 2)
    ib, jb = min(wi, ia + 2), min(wj, ja + 2)
    if rank == 0:
        import numpy as np
        from oracle import harness as H
        o = H.CpuContainer(H.oracle_api(w.complex_), w.params, w.basis_type).initialize()
        qa, qb = fq0, fq1
        Jw = J2d[: qb - qa].contiguous()
        got = c.projectMatrix(Jw, ia, ib, ja, jb, q0=qa, q1=qb).cpu().numpy()
        ref = o.project_matrix_qrange(Jw.cpu().numpy(), qa, qb, ia, ib, ja, jb, nthreads=min(8, os.cpu_count() or 1))
        # and the same pairs as they came out of the timed step (full window, this rank's partial before the exchange is not
        # kept, so at N > 1 this is the combined block against the oracle over ALL points when the workload is small enough)
        parity = {"normwise_err": float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))), "tolerance": 1e-12,
                  "window": f"pairs i in [{ia},{ib}) x j in [{ja},{jb}), ALL of rank 0's points q in [{qa},{qb}), all {m2} columns, "
                            "vs the oracle's (i,j,q) loop" + (" (extended oracle: the reference stops at 5 parameters, BasisHelper.cpp:63-65; "
                            "pinned bit-for-bit against the compiled reference for <= 5)" if w.nvars > 5 else "")}
        if world == 1:
            step()
            full = (C[0] if batched else C)[ia:ib, ja:jb].cpu().numpy()
            parity["timed_step_output_err"] = float(np.max(np.abs(full - ref)) / np.max(np.abs(ref)))

    # ---- optional: the same full window with the i<->j symmetry exploited (reported separately) ----
    sym = None
    if wi == wj and not args.no_sym and not batched:
        def sym_step():
            c.projectMatrix(J, 0, wi, 0, wj, q0=q0, q1=q1, out=C, symmetric=True)
            if world > 1:
                dist.all_reduce(Cr, op=dist.ReduceOp.SUM)
        sym_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            sym_step()
        s1.record()
        barrier()
        sms = maxr(s0.elapsed_time(s1)) / args.steps
        BI = 16
        tiles_i, tiles_j = (wi + BI - 1) // BI, (wj + 7) // 8
        done = sum(1 for ti in range(tiles_i) for tj in range(tiles_j) if min(wj - 1, tj * 8 + 7) >= ti * BI)
        exec_flop = done * (BI * 8) * float(nq) * ((m2 * width + 95) // 96 * 96 if m2 * width >= 96 else m2 * width) * 2.0 * (2 if w.complex_ else 1)
        sym = {"value": updates / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms,
               "executed_pair_tiles": done, "all_pair_tiles": tiles_i * tiles_j,
               "hw_frac_from_executed_flop": exec_flop / (sms * 1e-3) * 1e-12 / peak,
               "note": "same output (all ordered pairs of the window); only tiles holding a pair j >= i are computed and every block "
                       "is stored to (i,j) and (j,i).  'value' counts all N^2 Q updates delivered (effective rate); the reference "
                       "computes every ordered pair, so the headline `value` above is the dense run."}

    # ---- end to end: host buffers in, host buffer out, copies inside the timed region --------------------------
    e2e, e2e_pageable = None, None
    if not args.no_e2e:
        if batched:
            # host call per element (pspace_cuda_project_matrix_host): pinned J_e in, pinned C_e out; 2 elements per rank
            n_e2e = min(2, max(1, nE))
            Jh = torch.empty((n_e2e, Q, m2), dtype=tdt, pin_memory=True)
            Jh.copy_((J[:n_e2e].view(torch.complex128) if w.complex_ else J[:n_e2e]))
            del J, Jflat, C, Cr
            torch.cuda.empty_cache()
            Ch = torch.empty((n_e2e, wi, wj, m2), dtype=tdt, pin_memory=True)

            def e2e_step():
                for e in range(n_e2e):
                    c.projectMatrixHost(Jh[e], 0, wi, 0, wj, out=Ch[e])
            e2e_updates = float(n_e2e) * world * wi * wj * Q
            api = (f"pspace_cuda_project_matrix_host per element (C-ABI, pinned host J_e -> pinned host C_e), {n_e2e} element(s) per rank per step; "
                   "the 7.7 GB read-back of an element's blocks dominates")
            h2d, d2h = int(Jh.numel() * Jh.element_size()), int(Ch.numel() * Ch.element_size())
        else:
            Jh = torch.empty((nq, m2), dtype=tdt, pin_memory=True)
            Jh.copy_(J)
            del J, Jflat
            torch.cuda.empty_cache()
            e2e_updates = updates
            if world == 1:
                Ch = torch.empty((wi, wj, m2), dtype=tdt, pin_memory=True)

                def e2e_step():
                    c.projectMatrixHost(Jh, 0, wi, 0, wj, out=Ch)             # pspace_cuda_project_matrix_host
                api = "pspace_cuda_project_matrix_host (C-ABI, pinned host J -> host C; J uploaded in growing slabs overlapped with the projection)"
                h2d, d2h = int(Jh.numel() * Jh.element_size()), int(Ch.numel() * Ch.element_size())
            else:
                del C, Cr
                torch.cuda.empty_cache()
                sp = ShardedProjector(c, rank, world)
                shared = SharedPinnedOutput((wi, wj, m2), tdt)

                def e2e_step():
                    sp.projectMatrixHostShared(Jh, shared, 0, wi, 0, wj)
                per = (wi + world - 1) // world
                api = ("ShardedProjector.projectMatrixHostShared: pspace_cuda_project_matrix_hostin (slab-pipelined upload of the rank's pinned J "
                       "shard) -> NCCL reduce-scatter over blocks of basis rows -> each rank copies ITS row block into its slice of ONE "
                       "page-locked shared host buffer (the whole C once per job, not once per rank); bytes are per rank")
                h2d, d2h = int(Jh.numel() * Jh.element_size()), int(per * wj * m2 * width * 8)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = maxr(time.perf_counter() - t0)
        e2e = {"value": e2e_updates * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt / args.steps * 1e3, "api": api}
        if world > 1 and not batched:
            # the shared host result (row blocks written by different ranks) against a device recomputation of the parity window
            Cw = c.projectMatrix(Jh.to(dev), ia, ib, ja, jb, q0=q0, q1=q1)
            dist.all_reduce(torch.view_as_real(Cw) if w.complex_ else Cw, op=dist.ReduceOp.SUM)
            if rank == 0:
                hw = shared.tensor[ia:ib, ja:jb].to(dev)
                e2e["shared_host_result_vs_device_window"] = float((hw - Cw).abs().max() / Cw.abs().max())
            del Cw
            barrier()
            shared.close()
        if world == 1 and not batched and rank == 0:
            # the same call on PAGEABLE host memory (what a drop-in caller holding std::vector storage passes): one step
            import numpy as np
            Jp = np.empty((nq, m2), dtype=c.np_dtype)
            Jp[...] = Jh.numpy()
            Cp = np.empty((wi, wj, m2), dtype=c.np_dtype)
            c.projectMatrixHost(Jp, 0, wi, 0, wj, out=Cp) if Jp.nbytes < 1e9 else None
            t0 = time.perf_counter()
            c.projectMatrixHost(Jp, 0, wi, 0, wj, out=Cp)
            dtp = time.perf_counter() - t0
            e2e_pageable = {"value": updates / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3,
                            "note": "same C-ABI call with pageable numpy buffers (the driver stages the copies): what "
                                    "ParameterContainer::projectMatrix(m2, J, C) on std::vector storage costs; one step"}
            del Jp, Cp

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (rank 0's launch) ---------------------------------------------------
    kms = statistics.mean(kern_ms)
    flop_launch = float(max(1, nE)) * wi * wj * nq * flop_per_update
    achieved = flop_launch / (kms * 1e-3) * 1e-12
    traffic, traffic_src = None, "no ncu record of this source for this (workload, N): omitted"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        rec = tj.get(f"{w.name}@{world}")
        sha = file_sha(os.path.join(ROOT, "pspace_b200", "csrc", "project.cu"))
        if rec and rec.get("project_cu_sha256") == sha:
            traffic, traffic_src = rec["dram_bytes_per_launch"], rec.get("source", "profiles/traffic.json")
        elif rec:
            traffic_src = "profiles/traffic.json holds a record of an OLDER project.cu for this (workload, N): not reported"
    except Exception:
        pass
    fast = "tma" in plan
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": ("k_project_ws (persistent stream-K, TMA + mbarrier, DMMA.8x8x4 on the tensor pipe's FP64 sub-pipe)" if fast else
                           "k_project (generic all-warps cp.async kernel: operand TMA cannot address, e.g. an odd column count)"),
                "peak_source": f"measured in this run: FP64 DMMA {peak_dmma:.2f} / DFMA {peak_dfma:.2f} TFLOP/s register-resident loops"
                               + (f", cuBLAS DGEMM 8192^3 {peak_dgemm:.2f} TFLOP/s (torch.matmul f64)" if peak_dgemm else "")
                               + "; the largest is the denominator (MEASURED_PEAKS.json has no FP64 figure; bf16 peak does not apply to an FP64 path)",
                "kernel_ms": kms, "flop_per_launch": flop_launch,
                "note": "event time spans the whole project call" + (f" ({nE} elements)" if batched else "") +
                        ": basis-panel kernel (one launch for a square window, else two) + k_project_ws + k_fixup; useful FLOP only (fragment DMULs and padded tile rows are not counted)"}

    cpu = None
    if not args.no_cpu and world == 1:
        cpu, _ = cpu_sample(w, args.cpu_seconds)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128" if w.complex_ else "f64", "data": "synthetic", "config": cfg,
            "run": {"plan": plan, "panel_cache": 0, "l2_flush_between_steps": bool(small)},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline}
    if batched:
        line["per_element"] = {"ms": ms_per_step / E, "elements_per_s": E / (ms_per_step * 1e-3),
                               "nominal_elements": w.elements,
                               "extrapolated_s_for_nominal": w.elements * ms_per_step * 1e-3 / E,
                               "note": "elements are independent and identical in cost: the nominal 10k-element sweep is the per-element time x 10k (extrapolated, not run)"}
    if other:
        line["other_collective"] = other
    if fused_note:
        line["run"]["fused_note"] = fused_note
    line["vector_projection"] = vector
    line["parity"] = parity
    if sym:
        line["symmetric"] = sym
    if e2e:
        line["e2e"] = e2e
    if e2e_pageable:
        line["e2e_pageable"] = e2e_pageable
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.gpus > 1 and "RANK" not in os.environ:
        # the driver launches N > 1 through torch.distributed.run; do the same when invoked directly
        port = 29500 + (os.getpid() % 2000)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
