#!/usr/bin/env python
"""Per source line of one kernel: stall samples, executed instructions, L2 sectors requested by its global loads/stores.

usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name K > K.csv
       python profiles/ncu_source_lines.py K.csv [min_sectors]
The csv interleaves one row per CUDA line (its totals) with the SASS rows that belong to it; only the former are kept.
"""
import csv
import sys


def main(path, min_sectors=1):
    rows = list(csv.reader(open(path)))
    file_name = ""
    hdr = None
    out = []
    for r in rows:
        if r and r[0] == "File Path":
            file_name = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].strip().isdigit():
            g = lambda name: float((r[hdr.index(name)] or "0").replace(",", "")) if name in hdr else 0.0
            out.append((file_name, int(r[0]), r[1].strip(), g("# Samples"), g("Instructions Executed"),
                        g("L2 Theoretical Sectors Global"), g("L2 Theoretical Sectors Global Ideal")))
    tot_s = sum(o[3] for o in out) or 1.0
    tot_sec = sum(o[5] for o in out) or 1.0
    print(f"total samples {tot_s:.0f}, total L2 sectors (theoretical) {tot_sec:.0f}")
    for f, ln, src, smp, ins, sec, ideal in out:
        if sec >= min_sectors or smp > 0.02 * tot_s:
            print(f"{f}:{ln:<4d} smp {100 * smp / tot_s:5.1f}%  sectors {sec:9.0f} ({100 * sec / tot_sec:4.1f}%)  | {src[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1)
