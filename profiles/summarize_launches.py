#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    unit = None
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        unit = r[ui]
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, total {tot * scale:.1f} us (serialised, cold-cache)")
    print(f"{'kernel':48s} {'n':>5s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:48s} {v[0]:5d} {v[1] * scale:10.1f} {v[1] * scale / v[0]:8.1f} {v[1] / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
