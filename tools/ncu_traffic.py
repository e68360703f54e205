#!/usr/bin/env python
"""Per-launch facts of the hot kernels from `ncu --set full` reports, as the JSON map that bench.py reads
(profiles/ncu_traffic.json):

  families.<family>       dram__bytes_read.sum + dram__bytes_write.sum per launch group  -> roofline.traffic
  fp64.<family>           fp64-pipe warp instructions (DADD + DMUL + DFMA + DSETP, *executed*, from the SASS page of the
                          report) per launch group and per cell, and ncu's own sm__pipe_fp64_cycles_active percentage
  stamp                   sha256 of the kernel sources the capture was taken from: bench.py refuses to quote the numbers
                          next to a library built from other sources (it then reports them as stale)

Usage: python tools/ncu_traffic.py out.json cells_per_launch report1.ncu-rep [report2.ncu-rep ...]"""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
DP_OPS = ("DADD", "DMUL", "DFMA", "DSETP")


def source_stamp():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "nyles_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    it = hdr.index("gpu__time_duration.sum")
    ip = hdr.index("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    for r in data:
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        yield r[ik], b, float(r[it]), float(r[ip])


def dp_instructions(rep):
    """{kernel key: executed fp64 warp instructions of its first launch in the report} from the SASS page."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res, name = {}, None
    n = 0
    while n < len(rows):
        r = rows[n]
        if r and r[0] == "Kernel Name":
            name = r[1]
        if r and r[0] == "Address" and name is not None:
            hdr = r
            ie, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
            total = 0
            n += 1
            while n < len(rows) and len(rows[n]) == len(hdr) and rows[n][0] != "Address":
                parts = rows[n][isrc].split()
                op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "")
                if op.split(".")[0] in DP_OPS:
                    total += int(rows[n][ie])
                n += 1
            res.setdefault(key_of(name), total)
            name = None
            continue
        n += 1
    return res


def key_of(name):
    m = re.search(r"(k_\w+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")).replace("(bool)", "").replace("(int)", "").replace(" ", "")


def main():
    cells = float(sys.argv[2])
    per, pipe, dp = {}, {}, {}
    for rep in sys.argv[3:]:
        for name, b, t, p in launches(rep):
            per.setdefault(key_of(name), []).append(b)
            pipe.setdefault(key_of(name), []).append(p)
        dp.update(dp_instructions(rep))
    fam, fp64 = {}, {}
    mean = lambda v: sum(v) / len(v)                      # noqa: E731
    big = lambda k: max(per[k]) if k in per else 0.0      # noqa: E731  the finest-level launch of a leg instance

    def add_fp64(family, keys):
        tot = sum(dp.get(k, 0) for k in keys)
        if tot:
            fp64[family] = {"dp_warp_instr_per_launch_group": tot, "dp_instr_per_cell": tot * 32.0 / cells,
                            "ncu_pipe_fp64_pct": max(max(pipe[k]) for k in keys if k in pipe), "kernels": keys}
    mom = [k for k in per if k.startswith("k_mom3")]
    momf = [k for k in per if k.startswith("k_momentum")]
    if mom or momf:
        # the interior launch plus the frame launches (six per group; all of them are in the report when it was
        # taken with --launch-count covering one group)
        fam["rhs_momentum"] = sum(max(per[k]) for k in mom) + sum(sum(per[k]) / max(1, len(per[k]) // 6 if mom else len(per[k])) for k in momf)
        add_fp64("rhs_momentum", mom + momf)
    up = [k for k in per if k.startswith("k_upwind") or k.startswith("k_up3")]
    if up:
        fam["rhs_tracer"] = sum(mean(per[k]) for k in up)
        add_fp64("rhs_tracer", up)
    if any(k.startswith("k_vleg") for k in per):
        fam["mg_down_fine"] = big("k_vleg<0,1,1>") + big("k_vleg<0,1,0>")
        fam["mg_up_fine"] = big("k_vleg<1,2,1>") + big("k_vleg<1,2,0>")
    json.dump({"stamp": source_stamp(), "cells_per_launch": cells, "families": fam, "fp64": fp64,
               "kernels": {k: {"launches": len(v), "mean_bytes": mean(v), "max_bytes": max(v)} for k, v in per.items()}},
              open(sys.argv[1], "w"), indent=1)
    print(json.dumps({"families": fam, "fp64": {k: v["dp_instr_per_cell"] for k, v in fp64.items()}}))


if __name__ == "__main__":
    main()
