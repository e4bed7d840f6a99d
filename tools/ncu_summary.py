#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch, --set full --import-source on) as markdown for profiles/.

usage: tools/ncu_summary.py gpurun_out/x.ncu-rep "title" [kernel-regex] > profiles/x.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock during capture"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (DMMA lives here), % of elapsed"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64 (DFMA/DMUL) pipe active, % of elapsed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]


KERNEL = []


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"] + KERNEL, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], sys.argv[2]
    if len(sys.argv) > 3:                       # optional: regex selecting one kernel of a multi-kernel report
        KERNEL.extend(["--kernel-name", "regex:" + sys.argv[3]])
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"# {title}\n")
    print(f"source: `{rep}` (ncu --set full --clock-control none --import-source on; one launch, ~40 replay passes —")
    print("durations under the profiler are not bench values)\n")
    print(f"kernel: `{d.get('Kernel Name', ('?',))[0]}`\n")
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in d:
            print(f"| {label} (`{k.split('.', 2)[-1] if k.startswith(('TPC', 'SM_')) else k}`) | {d[k][0]} {d[k][1]} |")
    print("\n## warp stall samples (all warps)\n\n| reason | samples | share |\n|---|---|---|")
    st = {h.split("stalled_")[1]: float(d[h][0] or 0) for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h}
    tot = sum(st.values()) or 1
    for k, v in sorted(st.items(), key=lambda x: -x[1]):
        if v > 0:
            print(f"| {k} | {int(v)} | {100 * v / tot:.1f} % |")
    src = page(rep, "source")
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    rows = src[2:]
    total = sum(int(r[ix["# Samples"]]) for r in rows) or 1
    agg = {}
    for r in rows:
        t = r[ix["Source"]].split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        a = agg.setdefault(op, [0, 0])
        a[0] += int(r[ix["# Samples"]])
        a[1] += int(r[ix["Instructions Executed"]])
    print("\n## samples by SASS opcode (top 12)\n\n| opcode | samples | share | warp-instructions executed |\n|---|---|---|---|")
    for op, (s, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:12]:
        print(f"| {op} | {s} | {100 * s / total:.1f} % | {n} |")
    cnt = {}
    for r in rows:
        for m in ("DMMA", "UTMALDG", "SYNCS", "USETMAXREG", "LDGSTS", "DMUL", "DFMA", "LDS", "STS", "STG", "LDG"):
            if m in r[ix["Source"]]:
                cnt[m] = cnt.get(m, 0) + 1
    print("\nstatic SASS instruction counts in this kernel: " + ", ".join(f"{k} {v}" for k, v in sorted(cnt.items())))


if __name__ == "__main__":
    main()
