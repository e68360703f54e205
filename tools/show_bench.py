#!/usr/bin/env python
"""Pretty-print the JSON line(s) of bench.py: python tools/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print("%s: %.2f ms/step  %.3e cell-updates/s  step-roofline %.3f of measured  e2e %.3e  launches %s  vc/step %s" % (
        path, d["ms_per_step"], d["value"], d["roofline_step"]["frac_of_measured"], d["e2e"]["value"],
        d.get("gpu_launches"), d["config"].get("vcycles_per_step")))
    r = d.get("roofline") or {}
    print("   dominant: %s  %.0f GB/s  frac %.3f  avg %.3f ms" % (r.get("kernel"), r.get("achieved", 0), r.get("frac", 0), r.get("avg_ms", 0)))
    for k, v in d["roofline_step"]["kernel_time_share"].items():
        print("   %-20s %.4f  %6.2f ms/step" % (k, v, v * d["ms_per_step"]))
