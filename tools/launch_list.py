#!/usr/bin/env python3
"""Per-kernel shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).

usage: tools/launch_list.py profiles/r01_launches.csv > table.md
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
    acc = OrderedDict()
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        name = re.sub(r"ProjParams.*", "", name)[:110]
        key = (name, r["Grid Size"], r["Block Size"])
        ns = float(r["Metric Value"].replace(",", ""))
        n, t = acc.get(key, (0, 0.0))
        acc[key] = (n + 1, t + ns)
    total = sum(t for _, t in acc.values())
    print("| kernel | launches | block | total ms | share |")
    print("|---|---|---|---|---|")
    for (name, grid, block), (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name} grid={grid}` | {n} | {block} | {t * 1e-6:.3f} | {100 * t / total:.2f} % |")


if __name__ == "__main__":
    main()
