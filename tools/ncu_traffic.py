#!/usr/bin/env python
"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the hot kernels from `ncu --set full`
reports, as a JSON map  kernel family -> bytes per launch group, which bench.py puts into roofline.traffic.
Usage: python tools/ncu_traffic.py out.json report1.ncu-rep [report2.ncu-rep ...]"""
import csv
import json
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    it = hdr.index("gpu__time_duration.sum")
    for r in data:
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        yield r[ik], b, float(r[it])


def main():
    per = {}
    for rep in sys.argv[2:]:
        for name, b, t in launches(rep):
            m = re.search(r"(k_\w+)(<[^>]*>)?", name)
            key = (m.group(1) + (m.group(2) or "")).replace("(bool)", "").replace("(int)", "").replace(" ", "")
            per.setdefault(key, []).append(b)
    fam = {}
    mean = lambda v: sum(v) / len(v)                      # noqa: E731
    big = lambda k: max(per[k]) if k in per else 0.0      # noqa: E731  the finest-level launch of a leg instance
    for k, v in per.items():
        if k.startswith("k_momentum"):
            fam["rhs_momentum"] = mean(v)
        if k.startswith("k_upwind2"):
            fam["rhs_tracer"] = mean(v)
    if any(k.startswith("k_vleg") for k in per):
        fam["mg_down_fine"] = big("k_vleg<0,1,1>") + big("k_vleg<0,1,0>")
        fam["mg_up_fine"] = big("k_vleg<1,2,1>") + big("k_vleg<1,2,0>")
    json.dump({"families": fam, "kernels": {k: {"launches": len(v), "mean_bytes": mean(v), "max_bytes": max(v)} for k, v in per.items()}},
              open(sys.argv[1], "w"), indent=1)
    print(json.dumps(fam))


if __name__ == "__main__":
    main()
