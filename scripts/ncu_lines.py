#!/usr/bin/env python
"""Per-source-line instruction / stall-sample table from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
cur, data = None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] not in ("", "Line No") and r[2] == "-":
        try:
            data.append((int(r[7]), int(r[6]), cur, r[0], r[1].strip()))
        except ValueError:
            pass
tot, tots = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print(f"kernel {rx}: warp instructions {tot}, stall samples {tots}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]:9d} {100*d[0]/tot:5.1f}%  samples {100*d[1]/tots:5.1f}%  {d[2]}:{d[3]}: {d[4][:105]}")
print("--- by stall samples")
for d in sorted(data, key=lambda d: -d[1])[:15]:
    print(f"{d[1]:6d} {100*d[1]/tots:5.1f}%  inst {100*d[0]/tot:5.1f}%  {d[2]}:{d[3]}: {d[4][:105]}")
