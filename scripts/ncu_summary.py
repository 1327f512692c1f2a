#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, without a GPU) into profiles/<out>.json.

    python scripts/ncu_summary.py gpurun_out/attn_decode_r2.ncu-rep profiles/r2_ncu_attn_decode.json \
        --position 511 --algorithmic-bytes 168.7e6

Per captured launch: duration, DRAM bytes read+written, DRAM / tensor-pipe / SM throughput %, registers, achieved occupancy,
top stall reasons; the JSON carries the mean over the launches (what bench.py's roofline.traffic cites) plus the list."""
import argparse
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
UNIT = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--position", type=int, default=None)
    ap.add_argument("--algorithmic-bytes", type=float, default=None)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if not hdr:
        sys.exit("no launches in " + a.rep)
    names, units, data = rows[hdr[0]], rows[hdr[0] + 1], rows[hdr[0] + 2:]
    launches = []
    for r in data:
        if len(r) != len(names):
            continue
        d = {"kernel": r[names.index("Kernel Name")]}
        stalls = {}
        for n, u, v in zip(names, units, r):
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if n in WANT:
                d[WANT[n]] = x * UNIT.get(u, 1.0)
            if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
                stalls[n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = x
        d["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:4])
        launches.append(d)
    n = len(launches)
    mean = lambda k: sum(l.get(k, 0.0) for l in launches) / n
    out = {"report": a.rep, "kernel": launches[0]["kernel"][:120], "launches": n, "duration_us": mean("duration"),
           "dram_bytes_per_launch": mean("dram_read") + mean("dram_write"), "dram_pct": mean("dram_pct"),
           "tensor_pipe_pct": mean("tensor_pipe_pct"), "sm_pct": mean("sm_pct"), "l2_hit_pct": mean("l2_hit_pct"),
           "registers": mean("registers"), "achieved_occupancy_pct": mean("achieved_occupancy_pct"),
           "position": a.position, "algorithmic_bytes": a.algorithmic_bytes, "note": a.note, "per_launch": launches}
    if a.algorithmic_bytes:
        out["achieved_gbs_under_ncu"] = a.algorithmic_bytes / (out["duration_us"] * 1e-6) / 1e9
        out["traffic_over_algorithmic"] = out["dram_bytes_per_launch"] / a.algorithmic_bytes
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "per_launch"}, indent=1))


if __name__ == "__main__":
    main()
