#!/usr/bin/env python
"""Opcode histogram per kernel from the SASS of the shipped library: fp64-pipe instructions (DADD, DMUL, DFMA, DSETP),
TMA (UTMALDG / UTMAPF), mbarrier (SYNCS), shuffles, shared-memory loads / stores, global loads / stores, local-memory
(spill) traffic.  These are STATIC counts (every path of a kernel once); the executed counts per cell that bench.py
quotes come from the ncu capture (tools/ncu_traffic.py).
Usage: python tools/sass_count.py [lib.so] [name-filter]  > profiles/<round>_sass_histogram.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nyles_b200/libnyles_b200.so"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
GROUPS = [("fp64", ("DADD", "DMUL", "DFMA", "DSETP")), ("MUFU", ("MUFU",)), ("UTMALDG", ("UTMALDG",)), ("UTMAPF", ("UTMAPF",)),
          ("SYNCS", ("SYNCS",)), ("BAR", ("BAR",)), ("SHFL", ("SHFL",)), ("LDS", ("LDS",)), ("STS", ("STS",)),
          ("LDG", ("LDG",)), ("STG", ("STG",)), ("LDL/STL", ("LDL", "STL")), ("UTCxMMA", ("UTCHMMA", "UTCQMMA", "UTCIMMA"))]
cur, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        base = m.group(1).split(".")[0]
        counts[cur]["all"] += 1
        for g, ops in GROUPS:
            if base in ops:
                counts[cur][g] += 1
print("library: %s   code objects for: %s" % (lib, ", ".join(sorted(arch))))
print("%-78s %6s %s" % ("kernel", "all", " ".join("%7s" % g for g, _ in GROUPS)))
tot = collections.Counter()
for k, c in counts.items():
    if flt in k:
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)[:78]
        print("%-78s %6d %s" % (name, c["all"], " ".join("%7d" % c[g] for g, _ in GROUPS)))
        tot.update(c)
print("%-78s %6d %s" % ("TOTAL", tot["all"], " ".join("%7d" % tot[g] for g, _ in GROUPS)))
