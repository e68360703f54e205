#!/usr/bin/env python
"""Opcode histogram and stall totals of one kernel from `ncu --page source --csv`.
Usage: ncu -i rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_source_hist.py src.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
start = [n for n, r in enumerate(rows) if r and r[0] == "Address"]
for s in start[:int(sys.argv[2]) if len(sys.argv) > 2 else 1]:
    hdr = rows[s]
    data = []
    for r in rows[s + 1:]:
        if len(r) != len(hdr):
            break
        data.append(r)
    iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = sum(int(r[iS]) for r in data)
    ops, samp = Counter(), Counter()
    for r in data:
        parts = r[isrc].split()
        op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
        op = op.split(".")[0]
        ops[op] += int(r[iI]); samp[op] += int(r[iS])
    tote = sum(ops.values())
    print(rows[s - 1][1][:100] if s > 0 else "")
    print("SASS instructions: %d   executed warp-instr: %d   samples: %d" % (len(data), tote, tot))
    for op, c in ops.most_common(22):
        print("  %-10s exec %6.2f%%  samples %6.2f%%" % (op, 100 * c / tote, 100 * samp[op] / max(tot, 1)))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[hdr.index(h)]) for r in data) for h in stall_cols}
    print("  stalls:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
