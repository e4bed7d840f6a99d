#!/usr/bin/env python3
"""Per-kernel table from an ncu launch list taken with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X.csv <command>
best (shortest) launch of every kernel with its DRAM bytes:   python tools/launch_table.py X.csv  > profiles/X.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = next(r for r in rows if "Metric Name" in r)
ik, im, iv, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
bscale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
agg = collections.OrderedDict()
for r in rows:
    if r is h:
        continue
    name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
    v = float(r[iv].replace(",", ""))
    d = agg.setdefault(name, {"t": [], "rd": [], "wr": []})
    if r[im] == "gpu__time_duration.sum":
        d["t"].append(v * tscale.get(r[iu], 1e-6))
    elif r[im] == "dram__bytes_read.sum":
        d["rd"].append(v * bscale.get(r[iu], 1))
    elif r[im] == "dram__bytes_write.sum":
        d["wr"].append(v * bscale.get(r[iu], 1))
print("| kernel | launches | best duration (ms) | DRAM read (MB) | DRAM written (MB) | DRAM GB/s |\n|---|---|---|---|---|---|")
for k, d in agg.items():
    if not d["t"]:
        continue
    t = min(d["t"])
    i = d["t"].index(t)
    rd = d["rd"][i] if i < len(d["rd"]) else 0.0
    wr = d["wr"][i] if i < len(d["wr"]) else 0.0
    print(f"| `{k}` | {len(d['t'])} | {t:.4f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / t / 1e6:.0f} |")
