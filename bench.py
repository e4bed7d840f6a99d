#!/usr/bin/env python3
"""bench.py — Galerkin projection throughput (BASELINE.json metric) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over the whole workload: the matrix (Jacobian) projection
C[i][j][:] = sum_q w_q psi_i(z_q) psi_j(z_q) J[q][:] of BASELINE.json configs[4] (cfg5: 10 mixed parameters,
complete degree-3 basis N = 286, 4^10 = 1 048 576 quadrature points, 24x24 blocks) — all N^2 ordered pairs, all
q: 8.58e10 updates of 576 FMAs.  At N GPUs the quadrature points are sharded (strong scaling: the problem is
fixed) and the partial coefficient blocks are combined by one NCCL allreduce inside the timed step.

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
            "config": {"workload": w.name, "Q": w.Q, "m": w.m,
                       "note": "CPU arm: each step times a bounded window of the workload; updates/s is window-invariant"},
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pspace_b200.device import DeviceContainer, fill_synthetic, fp64_peak
    from pspace_b200.sharded import ShardedProjector, shard_bounds

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

    w = workload(args.workload)
    c = DeviceContainer(w.params, w.basis_type, w.complex_, device=local).initialize()
    c.setOption("panel_cache", 0)        # every step rebuilds the basis panels: no work is carried between steps
    N, Q, m2 = c.N, c.Q, w.m2
    wi, wj = (N, N) if w.pair_window is None else (min(N, w.pair_window[0]), min(N, w.pair_window[1]))
    width = 2 if w.complex_ else 1
    q0, q1 = shard_bounds(Q, world)[rank]
    nq = q1 - q0
    tdt = torch.complex128 if w.complex_ else torch.float64

    # synthetic J rows of this rank's shard (same global stream of numbers on any sharding)
    Jflat = fill_synthetic(nq * m2 * width, w.seed + 1000, offset=q0 * m2 * width, device=local)
    J = Jflat.view(torch.complex128).reshape(nq, m2) if w.complex_ else Jflat.reshape(nq, m2)
    C = torch.empty((wi, wj, m2), dtype=tdt, device=dev)
    Cr = torch.view_as_real(C) if w.complex_ else C

    # N > 1: the exchange of the partial coefficient blocks is fused into the kernels (peer-memory pushes from the
    # projection epilogue + reduce/all-gather kernel, pspace_b200/fused.py); the plain NCCL allreduce path is timed beside it.
    fp, fused_note = None, None
    if world > 1:
        try:
            from pspace_b200.fused import FusedShardedProjector
            fp = FusedShardedProjector(c, 0, wi, 0, wj, ncols=m2)
        except Exception as e:                      # no peer access / symmetric memory unavailable: NCCL path only
            fp, fused_note = None, f"fused peer path unavailable ({type(e).__name__}: {e}); NCCL allreduce used"
    ok = torch.tensor([1 if fp is not None else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            fp = None

    def nccl_step():
        c.projectMatrix(J, 0, wi, 0, wj, q0=q0, q1=q1, out=C)
        if world > 1:
            dist.all_reduce(Cr, op=dist.ReduceOp.SUM)

    def project_only():
        if fp is not None:
            fp.projectMatrix(J)
        else:
            c.projectMatrix(J, 0, wi, 0, wj, q0=q0, q1=q1, out=C)

    use_peer = fp is not None and args.collective == "peer"

    def step():
        if use_peer:
            fp.projectMatrix(J)
        else:
            nccl_step()

    updates = float(wi) * wj * Q                                   # all ordered pairs of the window x all q
    flop_per_update = m2 * (8 if w.complex_ else 2)

    # FP64 peak of this GPU, measured now (DMMA and DFMA register-resident loops); larger one is the denominator
    peak_dmma = fp64_peak(1, 400.0, local)
    peak_dfma = fp64_peak(0, 400.0, local)
    peak = max(peak_dmma, peak_dfma)

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    launches0 = c.launchCount()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 2)]
    barrier()
    ev[0].record()
    kern_ms = []
    for s in range(args.steps):
        a, b = ev[2 + 2 * s], ev[3 + 2 * s]        # events around this rank's projection kernels only (roofline)
        if use_peer:
            fp.projectMatrix(J, a, b)
        else:
            a.record()
            c.projectMatrix(J, 0, wi, 0, wj, q0=q0, q1=q1, out=C)
            b.record()
            if world > 1:
                dist.all_reduce(Cr, op=dist.ReduceOp.SUM)
    ev[1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[1])
    launches = c.launchCount() - launches0
    plan = c.lastPlan()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = updates / (ms_per_step * 1e-3)

    for s in range(args.steps):
        kern_ms.append(ev[2 + 2 * s].elapsed_time(ev[3 + 2 * s]))
    # N > 1: the same step with the exchange done by NCCL, timed beside
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
        nt = torch.tensor([n0.elapsed_time(n1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(nt, op=dist.ReduceOp.MAX)
        nccl_step()
        fp.projectMatrix(J)
        diff = float((fp.out - C).abs().max() / C.abs().max())
        other = {"collective": "nccl allreduce after the kernel" if use_peer else
                               "fused: partial tiles pushed to their owner rank over NVLink from the projection epilogue (peer-mapped "
                               "symmetric memory), then k_peer_reduce_gather (pspace_b200/fused.py)",
                 "ms_per_step": float(nt.item()), "value": updates / (float(nt.item()) * 1e-3), "unit": UNIT,
                 "max_rel_diff_between_the_two": diff}

    # ---- vector (residual-type) projection of the same container, m = block edge (SURVEY.md §8d asks for it beside) ------
    fvec = J[:, : w.m].contiguous()
    Fv = torch.empty((N, w.m), dtype=tdt, device=dev)

    def vec_step():
        c.projectVector(fvec, q0=q0, q1=q1, out=Fv)
        if world > 1:
            dist.all_reduce(torch.view_as_real(Fv) if w.complex_ else Fv, op=dist.ReduceOp.SUM)
    vec_step()
    barrier()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record()
    for _ in range(10):
        vec_step()
    v1.record()
    barrier()
    vt = torch.tensor([v0.elapsed_time(v1) / 10], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vt, op=dist.ReduceOp.MAX)
    vector = {"value": float(N) * Q / (float(vt.item()) * 1e-3), "unit": "(k,q) updates/s of m FMAs", "ms_per_step": float(vt.item()),
              "m": w.m, "note": "F[k][:] = sum_q w_q psi_k(z_q) f[q][:], all N basis rows, all q; f = first m columns of J"}

    # ---- live parity of this run's kernels against the CPU oracle: a 2x2 pair window over rank 0's first 32768 points ----
    parity = None
    if rank == 0:
        import numpy as np
        from oracle import harness as H
        o = H.CpuContainer(H.oracle_api(w.complex_), w.params, w.basis_type).initialize()
        qa, qb = q0, min(q1, q0 + 32768)
        ia, ja = N // 2, max(0, N - 2)
        got = c.projectMatrix(J[: qb - qa].contiguous(), ia, ia + 2, ja, ja + 2, q0=qa, q1=qb).cpu().numpy()
        ref = o.project_matrix_qrange(J[: qb - qa].cpu().numpy(), qa, qb, ia, ia + 2, ja, ja + 2, nthreads=2)
        parity = {"normwise_err": float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))), "tolerance": 1e-12,
                  "window": f"pairs i in [{ia},{ia + 2}) x j in [{ja},{ja + 2}), q in [{qa},{qb}), all {m2} columns, vs the oracle's (i,j,q) loop"}

    # ---- optional: the same full window with the i<->j symmetry exploited (reported separately) ----
    sym = None
    if wi == wj and not args.no_sym:
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
        tt = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sms = float(tt.item()) / args.steps
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
    e2e = None
    if not args.no_e2e:
        Jh = torch.empty((nq, m2), dtype=tdt, pin_memory=True)
        Jh.copy_(J)
        Ch = torch.empty((wi, wj, m2), dtype=tdt, pin_memory=True)
        del J, Jflat
        torch.cuda.empty_cache()

        sp = ShardedProjector(c, rank, world) if world > 1 else None

        def e2e_step():
            if world == 1:
                c.projectMatrixHost(Jh, 0, wi, 0, wj, out=Ch)             # pspace_cuda_project_matrix_host
            else:
                sp.projectMatrixHost(Jh, Ch, 0, wi, 0, wj, dev_out=C)     # hostin -> NCCL allreduce -> pinned host
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": updates * args.steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(Jh.numel() * Jh.element_size()), "d2h_bytes_per_step": int(Ch.numel() * Ch.element_size()),
               "ms_per_step": float(dt.item()) / args.steps * 1e3,
               "api": "pspace_cuda_project_matrix_host (C-ABI, pinned host J -> host C; J uploaded in slabs overlapped with the projection)" if world == 1 else
                      "ShardedProjector.projectMatrixHost: pspace_cuda_project_matrix_hostin (slab-pipelined upload of the pinned J shard) -> NCCL allreduce -> C to pinned host; bytes are per rank"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (rank 0's launch) ---------------------------------------------------
    kms = statistics.mean(kern_ms)
    flop_launch = float(wi) * wj * nq * flop_per_update
    achieved = flop_launch / (kms * 1e-3) * 1e-12
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(f"{w.name}@{world}")
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_project_ws (persistent stream-K, TMA + mbarrier, DMMA.8x8x4 on the tensor pipe's FP64 sub-pipe)",
                "peak_source": f"measured in this run: FP64 DMMA {peak_dmma:.2f} / DFMA {peak_dfma:.2f} TFLOP/s register-resident loops "
                               "(MEASURED_PEAKS.json has no FP64 figure; bf16 peak does not apply to an FP64 path)",
                "kernel_ms": kms, "flop_per_launch": flop_launch,
                "note": "event time spans the whole project call: 2 basis-panel kernels + k_project_ws + k_fixup; useful FLOP "
                        "only (fragment DMULs and padded tile rows are not counted)"}

    cpu = None
    if not args.no_cpu and world == 1:
        cpu, _ = cpu_sample(w, args.cpu_seconds)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128" if w.complex_ else "f64", "data": "synthetic",
            "config": {"workload": f"{w.name}: matrix (Jacobian) projection, all ordered pairs of a {wi}x{wj} basis window x all q",
                       "N": N, "Q": Q, "m": w.m, "params": w.nvars, "basis": "complete" if w.basis_type else "tensor",
                       "updates_per_step": updates, "flop_per_step": updates * flop_per_update,
                       "parallelism": f"q-shard x{world}" + ((" + fused peer-memory reduce-scatter/all-gather of C (NVLink pushes from the projection epilogue)"
                                                               if use_peer else " + NCCL allreduce of C") if world > 1 else ""),
                       "l2": f"inputs larger than L2: J shard {nq * m2 * width * 8 / 1e9:.2f} GB + C {wi * wj * m2 * width * 8 / 1e9:.2f} GB vs 126 MB L2 (no flush needed)",
                       "plan": plan, "panel_cache": 0},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline}
    if other:
        line["other_collective"] = other
    if fused_note:
        line["config"]["fused_note"] = fused_note
    line["vector_projection"] = vector
    line["parity"] = parity
    if sym:
        line["symmetric"] = sym
    if e2e:
        line["e2e"] = e2e
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
