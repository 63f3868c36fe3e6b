#!/usr/bin/env python
"""Aggregate an ncu report's source page per CUDA source line: executed warp instructions and stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr and r[0].isdigit():
        try: agg.append((fname, int(r[0]), r[1].strip()[:110], int(r[iE] or 0), int(r[iS] or 0)))
        except ValueError: pass
totE = sum(a[3] for a in agg) or 1; totS = sum(a[4] for a in agg) or 1
print(f"total warp-instructions {totE}, samples {totS}")
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    print(f"{a[0]:22s}:{a[1]:4d} exec {a[3]/totE*100:5.1f}% samp {a[4]/totS*100:5.1f}%  {a[2]}")
