#!/usr/bin/env python3
"""Static look at the DMMA loop of k_project_ws in the built library: decodes the SASS control words (stall count,
yield, scoreboard fields) and prints, per instantiation, the number of instructions and the sum of static stall cycles
of one chunk (8 k-steps).  ptxas produces one of two schedules for this loop depending on unrelated code in the kernel
(3222 vs 3249 static cycles per chunk and warp for the 128x96 real tile); the sum predicts the measured 1.2 % difference,
so a schedule regression can be caught here, without a GPU."""
import re
import subprocess
import sys
import os

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pspace_b200", "lib", "libpspace_cuda.so")


def parse(text):
    lines = text.split("\n")
    ins, i = [], 0
    while i < len(lines):
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
                ins.append((m.group(2).strip(), (w >> 105) & 0xF))
                i += 2
                continue
        i += 1
    return ins


def main():
    pats = sys.argv[1:] if len(sys.argv) > 1 else ["k_project_ws"]      # any of these substrings (one cuobjdump pass)
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if not any(pat in name for pat in pats):
            continue
        ins = parse(f)
        idx = [i for i, x in enumerate(ins) if "DMMA" in x[0]]
        if not idx:
            continue
        tw = [i for i, x in enumerate(ins[: idx[0]]) if "TRYWAIT" in x[0]]
        start = tw[-1] if tw else idx[0]
        # the RAG instantiations carry a second, rolled loop (the trimmed last chunk) after the main one: the main loop is
        # the first `main` DMMAs, the largest multiple of 64 (k-steps x 8x8 tiles of a 64-point stage)
        main = len(idx) if len(idx) % 64 == 0 else (len(idx) // 64) * 64
        if main == 0:
            main = len(idx)
        loop = ins[start: idx[main - 1] + 1]
        idx = idx[:main]
        short = re.sub(r"_ZN\d+_GLOBAL__N__[0-9a-f_]+project_cu_[0-9a-f_]+", "", name)[:70]
        print(f"{short:70s} instrs {len(ins):5d}  loop {len(loop):4d}  DMMA {len(idx):4d}  static stall cycles/chunk {sum(x[1] for x in loop):5d}")


if __name__ == "__main__":
    main()
