#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by source line: samples and instructions."""
import csv
import sys

path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []
cur_file = ""
with open(path) as f:
    for r in csv.reader(f):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0] == "" or r[2] != "-":
            continue  # sass rows
        try:
            rows.append((cur_file, int(r[0]), r[1].strip()[:110], int(r[4]), int(r[7])))
        except ValueError:
            pass
tot_s = sum(x[3] for x in rows) or 1
tot_i = sum(x[4] for x in rows) or 1
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
byfile = {}
for x in rows:
    a = byfile.setdefault(x[0], [0, 0]); a[0] += x[3]; a[1] += x[4]
for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:24s} samples {100*v[0]/tot_s:5.1f}%  instr {100*v[1]/tot_i:5.1f}%")
print("top lines by samples:")
for x in sorted(rows, key=lambda x: -x[3])[:topn]:
    print(f"{100*x[3]/tot_s:5.1f}% smp {100*x[4]/tot_i:5.1f}% ins  {x[0]}:{x[1]:<4d} {x[2]}")
