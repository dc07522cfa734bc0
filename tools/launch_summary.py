#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py gpurun_out/launches.csv "<command line that was profiled>" > profiles/x.txt"""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path, errors="replace")))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[i0]; col = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(lambda: [0.0, 0])
for r in rows[i0 + 1:]:
    if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    t = tot[r[col["Kernel Name"]]]
    t[0] += us; t[1] += 1
total = sum(t[0] for t in tot.values()); n = sum(t[1] for t in tot.values())
ours = sum(t[0] for k, t in tot.items() if "gigs::" in k or k.startswith("gigs") or "rs_" in k or "blend_" in k or "deferred" in k)
print(f"ncu --metrics gpu__time_duration.sum --clock-control none : {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("(cold-cache, serialised per-launch times: compare SHARES, not absolutes)")
print(f"total {total:.1f} us over {n} launches")
print(f"our kernels {ours:.1f} us = {100 * ours / total:.1f}% ; framework (torch) kernels {total - ours:.1f} us = {100 * (total - ours) / total:.1f}%\n")
for k, (us, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} us {100 * us / total:5.1f}%  n={c:4d}  {k[:120]}")
