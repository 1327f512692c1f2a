#!/usr/bin/env python
"""Development aid: hottest SASS instructions (warp-stall samples) of an ncu report captured with --import-source on.
    python scripts/ncu_hot.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out = []
hdr = None
for r in rows:
    if len(r) > 3 and 'Source' in r and 'Address' in r:
        if hdr is not None: break      # first kernel launch only
        hdr = r; cols = {c: i for i, c in enumerate(r)}; continue
    if hdr is None or len(r) != len(hdr): continue
    try: n = int(r[cols['# Samples']] or 0)
    except ValueError: continue
    out.append((n, r[cols['Source']]))
tot = sum(n for n, _ in out) or 1
print("total samples", tot, "instructions", len(out))
for i, (n, s) in enumerate(out):
    if 100.0 * n / tot >= minp: print(f"{i:5d} {n:7d} {100.0 * n / tot:5.1f}%  {s[:120]}")
