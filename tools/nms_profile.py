#!/usr/bin/env python
"""Target for `ncu --nvtx --nvtx-include "target/" --metrics gpu__time_duration.sum`: ONE oriented NMS over n boxes of the config-5 distribution
(tools/nms_sweep.py), after one untimed run.  argv: n [groups]."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_rpn_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ops.set_nms_cull_mode(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
g = torch.Generator().manual_seed(0)
c = torch.rand(n, 3, generator=g) * torch.tensor([256.0, 256.0, 160.0])
s = torch.rand(n, 3, generator=g) * 44 + 4
th = (torch.rand(n, 1, generator=g) - 0.5) * math.pi
boxes = torch.cat([c, s, th], 1).cuda().contiguous()
scores = torch.rand(n, generator=g).cuda()
grp = torch.randint(0, groups, (n,), generator=g).int().cuda() if groups > 1 else None
keep, nk = ops.nms_device(boxes, scores, grp, 0.3)
torch.cuda.synchronize()
import ctypes
from nerf_rpn_b200._lib import lib
st = (ctypes.c_ulonglong * 16)()
lib().nrpn_nms_cells_stats(st, 1)
torch.cuda.nvtx.range_push("target")
keep, nk = ops.nms_device(boxes, scores, grp, 0.3)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
lib().nrpn_nms_cells_stats(st, 0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); keep, nk = ops.nms_device(boxes, scores, grp, 0.3); b.record(); torch.cuda.synchronize()
print("kept", int(nk.item()), "ms", round(a.elapsed_time(b), 2), "cull mode", int(lib().nrpn_get_nms_cull_mode()))
names = ("records streamed", "pair tests", "exact IoU evaluations", "hits", "work items")
for mode, off in (("cross", 0), ("adjacency", 8)):
    print(mode, {k: int(st[off + i]) for i, k in enumerate(names)})
