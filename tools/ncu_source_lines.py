#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name ...`:
warp-instructions executed and stall samples per CUDA source line (top N), to see where a
kernel's issue slots go. Rows whose Address column is '-' are the per-CUDA-line aggregates."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hdr_i]
    ie, samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = {}
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if len(r) <= ie or not r[0].isdigit() or r[2] != "-":
            continue
        key = (fname, int(r[0]))
        n, s = int(r[ie] or 0), int(r[samp] or 0)
        old = data.get(key, (0, 0, r[1].strip()))
        data[key] = (old[0] + n, old[1] + s, r[1].strip())
    tot_i = sum(d[0] for d in data.values()) or 1
    tot_s = sum(d[1] for d in data.values()) or 1
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    for (f, line), (n, s, text) in sorted(data.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f[:16]:16s}:{line:4d} {100 * n / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% samp  {text[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
