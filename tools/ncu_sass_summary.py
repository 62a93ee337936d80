#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page source --csv --kernel-name K` (SASS view): executed warp instructions by opcode,
and the hottest instructions by stall samples. Usage: ncu_sass_summary.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ops = collections.Counter(); samp = collections.Counter()
tot = 0; tots = 0
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break      # next launch
    if len(r) > iI and r[iI].isdigit(): body.append(r)
for r in body:
    if len(r) <= iI: continue
    s = r[iS].strip()
    t = s.split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    op = op.split(".")[0]
    n = int(r[iI] or 0); k = int(r[iN] or 0)
    ops[op] += n; samp[op] += k; tot += n; tots += k
print("kernel:", rows[0][1][:100])
print("total warp instructions %d, samples %d" % (tot, tots))
for op, n in ops.most_common(18):
    print("  %-10s %6.2f%% inst  %6.2f%% samples" % (op, 100.0 * n / tot, 100.0 * samp[op] / max(tots, 1)))
print("hottest instructions by samples:")
order = sorted(range(len(body)), key=lambda i: -int(body[i][iN] or 0) if len(body[i]) > iI else 0)[:top]
for i in sorted(order):
    r = body[i]
    print("  #%4d %5.2f%% samp %5.2f%% inst  %s" % (i, 100.0 * int(r[iN]) / max(tots, 1), 100.0 * int(r[iI]) / tot, r[iS].strip()[:90]))
