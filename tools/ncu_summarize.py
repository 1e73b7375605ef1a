#!/usr/bin/env python
"""Collapse an `ncu --csv` metrics log (one row per kernel launch and metric) into a per-kernel table (markdown)."""
import collections
import csv
import sys

src, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    key = (r["ID"], r["Kernel Name"])
    per.setdefault(key, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(r["Metric Unit"], 1.0) \
        if r["Metric Name"] == "gpu__time_duration.sum" else (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
agg = collections.OrderedDict()
for (i, name), m in per.items():
    short = name.split("(")[0].replace("nrpn::", "").replace("void ", "")
    a = agg.setdefault(short, dict(n=0, us=0.0, rd=0.0, wr=0.0, tens=0.0, lts=0.0, l1=0.0))
    us = m.get("gpu__time_duration.sum", 0.0)
    def byt(k):
        v = m.get(k, (0.0, "byte"))
        return v[0] * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(v[1], 1.0)
    a["n"] += 1; a["us"] += us; a["rd"] += byt("dram__bytes_read.sum"); a["wr"] += byt("dram__bytes_write.sum")
    a["tens"] += us * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", (0.0, ""))[0]
    a["lts"] += us * m.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", (0.0, ""))[0]
    a["l1"] += us * m.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", (0.0, ""))[0]
tot = sum(a["us"] for a in agg.values())
print(f"# {title}\n")
print("ncu per-launch durations are cold-cache and serialised: use the SHARE of the step, not the absolute time.\n")
print("| kernel | launches | time us | share | DRAM GB/s (r+w) | tensor pipe active % | L2 thr % | L1/smem thr % |")
print("|---|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    if a["us"] <= 0:
        continue
    print(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {(a['rd'] + a['wr']) / a['us'] / 1e3:.0f} | "
          f"{a['tens'] / a['us']:.1f} | {a['lts'] / a['us']:.1f} | {a['l1'] / a['us']:.1f} |")
print(f"\ntotal {tot / 1e3:.3f} ms over {sum(a['n'] for a in agg.values())} launches")
