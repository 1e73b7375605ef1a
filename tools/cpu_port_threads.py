#!/usr/bin/env python
"""Thread-count sweep of the CPU port (oracle/net.py) on the bench box's host: picks the intra-op thread count bench.py uses."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
for th in (8, 16, 32, 64, 128):
    if th > (os.cpu_count() or 1):
        continue
    bench.cpu_port_run(32, th, y_extent=64)
    t = bench.cpu_port_run(32, th, y_extent=128)[0]
    print(f"threads {th:3d}: 32x128x256 block {t:.2f} s", flush=True)
for th in (16, 32, 64):
    if th > (os.cpu_count() or 1):
        continue
    t = bench.cpu_port_run(160, th)[0]
    print(f"threads {th:3d}: full scene {t:.2f} s", flush=True)
