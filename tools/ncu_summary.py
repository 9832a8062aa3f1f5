#!/usr/bin/env python
"""Summarise Nsight Compute captures on the CPU side (no GPU needed).

  launch list : ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
                    --csv --log-file launches.csv python tools/profile_step.py --steps 2
                tools/ncu_summary.py launches launches.csv [--skip-first N]      -> per-kernel share table (csv on stdout)
  full report : ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o rep <cmd>
                tools/ncu_summary.py report rep.ncu-rep                          -> key metrics + top warp-stall SASS sites

Numbers under a profiler are never bench values: compare SHARES, not absolutes (launches are serialised, cold-cache).
"""
import argparse
import collections
import csv
import io
import re
import subprocess
import sys

KEY_METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("bv2::", "").replace("tc::", "")


def launches(path, skip_first=0):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(lambda: collections.defaultdict(float))  # launch id -> metric -> value
    names = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        lid = int(r[idx["ID"]])
        names[lid] = short(r[idx["Kernel Name"]])
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        m = r[idx["Metric Name"]]
        if m.startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> us
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)  # -> MB
        per[lid][m] = v
    ids = sorted(per)[skip_first:]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for lid in ids:
        a = agg[names[lid]]
        a[0] += 1
        a[1] += per[lid].get("gpu__time_duration.sum", 0.0)
        a[2] += per[lid].get("dram__bytes_read.sum", 0.0) + per[lid].get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"# total {tot / 1e3:.2f} ms, {len(ids)} launches")
    print("kernel,launches,sum_us,share_pct,avg_us,dram_MB")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k},{a[0]},{a[1]:.1f},{100 * a[1] / tot:.1f},{a[1] / a[0]:.1f},{a[2]:.1f}")


def _ncu_csv(rep, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def report(rep, top=12):
    rows = _ncu_csv(rep, "--page", "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr) if i < len(r)}
        print("==", short(d["Kernel Name"][0]))
        for m in KEY_METRICS:
            if m in d:
                print(f"  {m:70s} {d[m][0]} {d[m][1]}")
    sass = _ncu_csv(rep, "--page", "source", "--print-source", "sass")
    hi = next((i for i, r in enumerate(sass) if r and r[0] == "Address"), None)
    if hi is None:
        return
    h = sass[hi]
    isamp, isrc = h.index("# Samples"), h.index("Source")
    data = []
    for r in sass[hi + 1:]:
        try:
            data.append((int(r[isamp]), r[isrc]))
        except (ValueError, IndexError):
            pass
    tot = sum(n for n, _ in data) or 1
    print(f"-- top warp-stall sampling sites ({tot} samples)")
    for n, s in sorted(data, reverse=True)[:top]:
        print(f"  {100 * n / tot:5.1f}%  {s[:100]}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("launches"); a.add_argument("csv"); a.add_argument("--skip-first", type=int, default=0)
    b = sub.add_parser("report"); b.add_argument("rep"); b.add_argument("--top", type=int, default=12)
    args = ap.parse_args()
    if args.cmd == "launches":
        launches(args.csv, args.skip_first)
    else:
        report(args.rep, args.top)
