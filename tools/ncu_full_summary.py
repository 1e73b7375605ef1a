#!/usr/bin/env python
"""Summarise `ncu --set full` reports (raw page CSV) of the captured kernels into one JSON list (profiles/*_ncu_full_summary.json).
usage: ncu_full_summary.py out.json capture_name=report.ncu-rep [...]  (run where `ncu` is installed)"""
import csv
import io
import json
import subprocess
import sys

KEYS = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read_B", "dram__bytes_write.sum": "dram_write_B",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_smem_throughput_pct",
        "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}

out = []
for spec in sys.argv[2:]:
    name, path = spec.split("=", 1)
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = {"capture": name, "kernel": vals[hdr.index("Kernel Name")].split("(")[0]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS and v != "":
                rec[KEYS[h]] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
        if "time_us" in rec and rec["time_us"] > 0:
            rec["dram_GBps"] = (rec.get("dram_read_B", 0.0) + rec.get("dram_write_B", 0.0)) / rec["time_us"] / 1e3
        out.append(rec)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
