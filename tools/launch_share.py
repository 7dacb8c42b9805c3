#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
   python tools/launch_share.py gpurun_out/xxx_launches.csv [skip_first_n]"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            val = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
            rows.append((r["Kernel Name"], ns, r["Grid Size"], r["Block Size"]))
    rows = rows[skip:]
    tot = sum(r[1] for r in rows)
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        short = name.split("(")[0][-70:]
        a = agg.setdefault(short, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ns
    print("%d launches, %.1f us total (cold-cache, serialised under ncu: compare shares, not absolutes)" % (len(rows), tot / 1e3))
    print("%-72s %6s %10s %10s %7s  %s" % ("kernel", "count", "total_us", "avg_us", "share", "grid x block"))
    for k, (c, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %10.1f %10.2f %6.1f%%  %s x %s" % (k, c, ns / 1e3, ns / 1e3 / c, 100 * ns / tot, grid, block))


if __name__ == "__main__":
    main()
