#!/usr/bin/env python
"""Count fp64-pipe instructions (DADD, DMUL, DFMA, DSETP, MUFU.RCP64H) per kernel in the SASS of the library.
Usage: python tools/sass_count.py [lib.so] [name-filter]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nyles_b200/libnyles_b200.so"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        counts[cur]["all"] += 1
        if base in ("DADD", "DMUL", "DFMA", "DSETP") or op.startswith("MUFU.RCP64H"):
            counts[cur][base] += 1
for k, c in counts.items():
    if flt in k:
        dp = sum(v for n, v in c.items() if n != "all")
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:90]
        print("%-90s all %6d  fp64 %5d  %s" % (name, c["all"], dp, dict((n, v) for n, v in c.items() if n != "all")))
