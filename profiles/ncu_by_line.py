#!/usr/bin/env python
"""Aggregate an ncu source page by CUDA source line: python profiles/ncu_by_line.py report.ncu-rep [kernel-index]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["--kernel-id", sys.argv[2]] if len(sys.argv) > 2 else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
cur_file = None
hdr = None
first_kernel_done = False
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] == "":  # sass row
        continue
    try:
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        inst = int(r[hdr.index("Instructions Executed")]); tinst = int(r[hdr.index("Thread Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
    except Exception:
        continue
    a = agg.setdefault(key, [0, 0, 0]); a[0] += inst; a[1] += tinst; a[2] += samp
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print(f"total warp-inst {tot}  samples {tots}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[2]/tots:5.1f}% smp  thr/inst {a[1]/max(a[0],1):5.1f}  {k[0]}:{k[1]}  {k[2]}")
