"""CPU: the C-ABI library loads without a GPU, exports every entry point include/nerf_rpn_b200.h declares, and the
ctypes mirror of its structs has the C layout.  No compute calls here."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

from nerf_rpn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nerf_rpn_b200.h")


def test_library_loads_and_exports_header_symbols():
    L = _lib.lib()
    src = open(HEADER).read()
    declared = sorted(set(re.findall(r"\b(nrpn_[a-z0-9_]+)\s*\(", src)))
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.exported_symbols()) == declared            # the ctypes table covers exactly the header
    assert L.nrpn_version() >= 100
    assert L.nrpn_status_string(0) == b"ok" and L.nrpn_status_string(-3) == b"workspace too small"


def test_host_only_queries():
    L = _lib.lib()
    assert [L.nrpn_conv3d_block_n(c) for c in (8, 64, 72, 128, 256, 2048)] == [64, 64, 128, 128, 256, 256]
    assert L.nrpn_nms_max_boxes() >= 10000
    assert L.nrpn_nms_workspace_bytes(10000) > 10000 * 157 * 8
    assert L.nrpn_nms_workspace_bytes(0) > 0


def test_struct_layouts_match_c():
    code = r'''
#include <stdio.h>
#include <stddef.h>
#include "nerf_rpn_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(nrpn_conv_desc, workspace), sizeof(nrpn_conv_level), sizeof(nrpn_conv_desc), offsetof(nrpn_conv_desc, stride),
         offsetof(nrpn_conv_desc, w), offsetof(nrpn_conv_desc, level), sizeof(nrpn_rpn_level), sizeof(nrpn_rpn_desc),
         offsetof(nrpn_rpn_desc, cell_anchors), offsetof(nrpn_rpn_desc, nms_thresh), offsetof(nrpn_rpn_desc, valid));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    C, R = _lib.ConvDesc, _lib.RpnDesc
    mine = [C.workspace.offset, ctypes.sizeof(_lib.ConvLevel), ctypes.sizeof(C), C.stride.offset, C.w.offset, C.level.offset,
            ctypes.sizeof(_lib.RpnLevel), ctypes.sizeof(R), R.cell_anchors.offset, R.nms_thresh.offset, R.valid.offset]
    assert mine == vals


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libnerf_rpn_b200.so")
    with pytest.raises(_lib.NativeLibraryError, match="no CPU / PyTorch fallback"):
        _lib.lib()


def test_ops_reject_cpu_tensors():
    import torch
    from nerf_rpn_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.iou3d_pairs(torch.zeros(2, 7), torch.zeros(2, 7))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.pack_stem_input(torch.zeros(1, 4, 4, 4, 4))
