#!/usr/bin/env python
"""Print the per-launch sequence (id, stream, kernel, grid, block, us, dram MB) of an ncu launch-list csv; --from/--to select ids."""
import argparse, collections, csv, re
ap = argparse.ArgumentParser(); ap.add_argument("path"); ap.add_argument("--lo", type=int, default=0); ap.add_argument("--hi", type=int, default=1 << 30)
a = ap.parse_args()
rows = [r for r in csv.reader(l for l in open(a.path) if l.startswith('"'))]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr): continue
    lid = int(r[idx["ID"]])
    d = per.setdefault(lid, {"name": re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", r[idx["Kernel Name"]])).replace("bv2::", ""), "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]], "stream": r[idx["Stream"]], "us": 0.0, "mb": 0.0})
    v = float(r[idx["Metric Value"]].replace(",", "")); u = r[idx["Metric Unit"]]; m = r[idx["Metric Name"]]
    if m.startswith("gpu__time"): d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    else: d["mb"] += v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
for lid, d in per.items():
    if a.lo <= lid < a.hi: print(f"{lid:4d} s{d['stream']:>3} {d['name']:<44} {d['grid']:<16} {d['block']:<12} {d['us']:8.1f} us {d['mb']:8.1f} MB")
