#!/usr/bin/env python3
"""DRAM traffic of the dominant kernel of bench.py's step, for `roofline.traffic`:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_project_ws -c 1 --launch-skip 1 \
        --csv --log-file gpurun_out/traffic_cfg5.csv python tools/prof_kernel.py 5 1048576
    python tools/capture_traffic.py gpurun_out/traffic_cfg5.csv cfg5_10param_complete_p3_m24@1 "r02: ncu ... (command above)"

merges {"dram_bytes_per_launch", "project_cu_sha256", "source"} into profiles/traffic.json under the key; bench.py reports
the number only while pspace_b200/csrc/project.cu still hashes to the recorded value (never a stale constant)."""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, key, source = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Metric Name" in r)
    i_name, i_val, i_unit = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = 0.0
    for r in rows:
        if r is hdr or len(r) <= i_val or not r[i_name].startswith("dram__bytes_"):
            continue
        v = float(r[i_val].replace(",", ""))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[r[i_unit]]
        tot += v * scale
    sha = hashlib.sha256(open(os.path.join(ROOT, "pspace_b200", "csrc", "project.cu"), "rb").read()).hexdigest()[:16]
    out = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(out))
    except Exception:
        d = {}
    d[key] = {"dram_bytes_per_launch": tot, "project_cu_sha256": sha, "source": source}
    json.dump(d, open(out, "w"), indent=1, sort_keys=True)
    print(key, f"{tot / 1e9:.2f} GB per launch", sha)


if __name__ == "__main__":
    main()
