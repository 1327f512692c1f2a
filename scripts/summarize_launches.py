#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: summarize_launches.py launches.csv [top_n]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total device time (serialised, cold-cache: compare shares)")
    print("| kernel | launches | total ms | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {n[:90]} | {c} | {ns / 1e6:.3f} | {ns / c / 1e3:.1f} | {100 * ns / tot:.1f}% |")


if __name__ == "__main__":
    main()
