#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and share of device time.  Usage: python tools/summarize_launches.py launches.csv [skip_first_N]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = [l for l in open(path) if l.startswith('"')]
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in list(csv.DictReader(rows))[skip:]:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"<unnamed>::", "", name)
        ns = float(r["Metric Value"])
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        tot[name] += ns
        cnt[name] += 1
    total = sum(tot.values())
    print("%-72s %7s %12s %9s %7s" % ("kernel", "count", "total_ms", "avg_us", "share"))
    for k in sorted(tot, key=lambda k: -tot[k]):
        print("%-72s %7d %12.3f %9.1f %6.1f%%" % (k[:72], cnt[k], tot[k] / 1e6, tot[k] / cnt[k] / 1e3, 100 * tot[k] / total))
    print("%-72s %7d %12.3f" % ("TOTAL", sum(cnt.values()), total / 1e6))


if __name__ == "__main__":
    main()
