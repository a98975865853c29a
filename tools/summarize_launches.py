#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys


def main(path, first=None, count=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if first is not None:
        rows = rows[int(first): int(first) + int(count)]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        name = r["Kernel Name"]
        m = re.search(r"(\w+)(<[^>]*>)?\(", name)
        short = (m.group(1) + (m.group(2) or "")) if m else name[:40]
        key = (short, r["Grid Size"], r["Block Size"])
        us = float(r["Metric Value"].replace(",", "")) / 1e3
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    print(f"{len(rows)} launches, {total/1e3:.3f} ms total (ncu per-launch times: cold-cache, serialised)")
    print(f"{'kernel':44s} {'grid':>16s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
    for (short, grid, block), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{short:44s} {grid:>16s} {n:4d} {us:10.1f} {us/n:9.1f} {100*us/total:5.1f}%")


if __name__ == "__main__":
    main(*sys.argv[1:])
