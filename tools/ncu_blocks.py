#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export per kernel section: runs of SASS instructions with the same execution
count (~basic blocks) with their share of executed instructions, stall samples and average active lanes.
usage: tools/ncu_blocks.py <source.csv> [section index ...] [--min PCT] [--ops]"""
import collections
import csv
import sys

path = sys.argv[1]
args = sys.argv[2:]
min_pct = 1.0
if "--min" in args:
    i = args.index("--min"); min_pct = float(args[i + 1]); del args[i:i + 2]
ops = "--ops" in args
if ops:
    args.remove("--ops")
want = [int(a) for a in args]
sections, cur = [], None
for r in csv.reader(open(path)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif r and cur is not None and len(r) > 5:
        cur["rows"].append(r)
for si, kx in enumerate(sections):
    if want and si not in want:
        continue
    h = kx["hdr"]
    iI, iS, iT, iN = h.index("Instructions Executed"), h.index("Source"), h.index("Thread Instructions Executed"), h.index("# Samples")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    reg = [(int(r[iI]), int(r[iT]), int(r[iN]), r[iS].strip(), r) for r in kx["rows"]]
    tot = sum(x[0] for x in reg); tots = sum(x[2] for x in reg)
    print("== section %d: %s\n   SASS lines %d, warp instructions %d, samples %d" % (si, kx["name"][:90], len(reg), tot, tots))
    st = collections.Counter()
    for x in reg:
        for i, c in stall_cols:
            st[c] += int(x[4][i] or 0)
    print("   stalls:", ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / max(1, tots)) for c, v in st.most_common(7)))
    if ops:
        oc = collections.Counter()
        for x in reg:
            op = x[3].split()[1] if x[3].startswith("@") else x[3].split()[0]
            oc[op.split(".")[0]] += x[0]
        print("   opcodes:", ", ".join("%s %.1f%%" % (o, 100.0 * v / tot) for o, v in oc.most_common(14)))
    start = 0
    for i in range(1, len(reg) + 1):
        if i == len(reg) or abs(reg[i][0] - reg[start][0]) > 0.02 * max(1, reg[start][0]):
            n = sum(x[0] for x in reg[start:i]); s = sum(x[2] for x in reg[start:i]); t = sum(x[1] for x in reg[start:i])
            if n > min_pct / 100.0 * tot:
                print("   [%4d,%4d) len %3d  exec/inst %10d  inst %5.1f%%  samples %5.1f%%  lanes %4.1f   %s | %s"
                      % (start, i, i - start, reg[start][0], 100.0 * n / tot, 100.0 * s / max(1, tots), t / max(1, n), reg[start][3][:36], reg[i - 1][3][:36]))
            start = i
