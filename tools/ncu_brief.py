#!/usr/bin/env python
"""Brief of an .ncu-rep: key raw metrics per kernel + the top stall sites of the SASS page.
    python tools/ncu_brief.py gpurun_out/X.ncu-rep [n_top]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed.avg.per_cycle_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("==", vals[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:70s} {vals[i]:>14s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    A, S, N, I = (hdr.index(x) for x in ("Address", "Source", "# Samples", "Instructions Executed"))
    stall = [(n, hdr.index(n)) for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    data = [r for r in rows[hi + 1:] if len(r) > I and r[A].startswith("0x") or (len(r) > I and r[A][:1].isdigit())]
    tot = sum(int(r[N] or 0) for r in data) or 1
    print(f"-- top stall sites ({tot} samples)")
    for r in sorted(data, key=lambda r: -int(r[N] or 0))[:top]:
        why = sorted(((int(r[i] or 0), n[6:]) for n, i in stall), reverse=True)[:2]
        print(f"   {r[A][-5:]} {100 * int(r[N]) / tot:5.1f}% x{r[I]:>8s} {why[0][1]:>14s} {r[S][:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
