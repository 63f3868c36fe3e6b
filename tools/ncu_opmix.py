#!/usr/bin/env python
"""SASS opcode mix (executed warp instructions, stall samples) of an ncu report: python tools/ncu_opmix.py rep [top]"""
import csv, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS = hdr.index('Source'); iN = hdr.index('# Samples'); iE = hdr.index('Instructions Executed')
ops = collections.Counter(); samp = collections.Counter(); tot = totS = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    t = r[iS].split()
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.split('.')[0]
    e = int(r[iE] or 0); s = int(r[iN] or 0)
    ops[op] += e; samp[op] += s; tot += e; totS += s
print('total warp-inst', tot, 'samples', totS)
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f'{op:10s} exec {c/tot*100:5.1f}%  samples {samp[op]/totS*100:5.1f}%')
