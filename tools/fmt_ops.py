"""jsonl of tools/bench_ops.py -> markdown table rows:  python tools/fmt_ops.py < profiles/r02_ops_cfg5.jsonl"""
import json
import sys

print("| op (§8 row) | ms | bound | achieved | peak | frac | rate | CPU oracle rate (1 thread) |\n|---|---|---|---|---|---|---|---|")
for line in sys.stdin:
    r = json.loads(line)
    if r["bound"] == "hbm":
        ach, peak = f'{r["achieved_gbs"]:.0f} GB/s', f'{r["hbm_peak_gbs"]:.0f} GB/s'
    else:
        ach, peak = f'{r["achieved_tflops"]:.2f} TFLOP/s', f'{r["fp64_peak_tflops"]:.2f} TFLOP/s'
    cpu = f'{r["cpu_rate"]:.3e} ({r["cpu_note"]})' if r.get("cpu_rate") else "—"
    print(f'| {r["op"]} | {r["ms"]:.3f} | {r["bound"]} | {ach} | {peak} | {r["frac"]:.3f} | {r["rate"]:.3e} {r["unit"]} | {cpu} |')
