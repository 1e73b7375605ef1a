#!/usr/bin/env python
"""GPU-box probe: which fp32 operation orders does the reference's torch-CUDA IoU chain use on this device?

The north star asks for bit-identical NMS keep sets against the reference RUN ON THE SAME GPU.  The reference's
chain (oriented_iou_loss.py:6-107, box_intersection_2d.py:11-176) is ~40 ATen kernels; three of its steps have a device-
dependent rounding order that the fused kernel must reproduce: sin/cos (ATen calls CUDA sinf/cosf), the 4x2 * 2x2 torch.bmm
(cuBLAS), and two torch.sum reductions (24 masked vertices -> mean, 8 shoelace terms -> area).  This script runs the staged
reference (oracle/_ref, real K1) on random boxes and checks candidate formulas for bit equality, per step, so the choice made
in csrc/box_iou.cuh is measured, not guessed.  Output: gpurun_out/ref_gpu_probe.json (copied to profiles/).
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_gpu  # noqa: E402


def rand_obb(n, g, extent=30.0, smin=1.0, smax=11.0):
    return torch.cat([torch.rand(n, 3, generator=g) * extent, torch.rand(n, 3, generator=g) * (smax - smin) + smin,
                      (torch.rand(n, 1, generator=g) - 0.5) * math.pi], 1)


def fma32(a, b, c):
    """round32(a*b + c) with a single rounding (a*b is exact in fp64; the fp64 add rounds to 53 bits first: double rounding is
    possible but needs a 29-bit tie pattern, negligible for a statistic)."""
    return (a.double() * b.double() + c.double()).float()


def biteq(a, b):
    return float((a.view(torch.int32) == b.view(torch.int32)).float().mean().item())


def sum_candidates(t):
    """t (..., n) fp32 -> dict of candidate summation orders over the last dim."""
    n = t.shape[-1]
    out = {}
    s = torch.zeros_like(t[..., 0])
    for i in range(n):
        s = s + t[..., i]
    out["sequential"] = s
    acc = [torch.zeros_like(t[..., 0]) for _ in range(4)]
    for i in range(n):
        acc[i % 4] = acc[i % 4] + t[..., i]
    out["4acc_interleaved_then_seq"] = ((acc[0] + acc[1]) + acc[2]) + acc[3]
    out["4acc_interleaved_then_tree"] = (acc[0] + acc[2]) + (acc[1] + acc[3])
    out["4acc_interleaved_then_tree2"] = (acc[0] + acc[1]) + (acc[2] + acc[3])
    # contiguous chunks of n/4
    if n % 4 == 0:
        q = n // 4
        ch = []
        for j in range(4):
            c = torch.zeros_like(t[..., 0])
            for i in range(q):
                c = c + t[..., j * q + i]
            ch.append(c)
        out["4chunks_then_seq"] = ((ch[0] + ch[1]) + ch[2]) + ch[3]
    # shuffle-down tree on a power-of-two padded vector
    m = 1
    while m < n:
        m *= 2
    v = [t[..., i] if i < n else torch.zeros_like(t[..., 0]) for i in range(m)]
    off = m // 2
    while off >= 1:
        v = [v[i] + v[i + off] if i + off < len(v) else v[i] for i in range(off)] + []
        off //= 2
    out["shfl_down_tree"] = v[0]
    # vector-of-4 loads per thread then tree across threads (inner-dim reduce with vec4)
    if n % 4 == 0:
        thr = []
        for j in range(n // 4):
            a = [t[..., 4 * j + k] for k in range(4)]
            thr.append(((a[0] + a[1]) + a[2]) + a[3])
        while len(thr) > 1:
            half = len(thr) // 2
            thr = [thr[i] + thr[i + half] for i in range(half)] + (thr[2 * half:] if len(thr) % 2 else [])
        out["vec4_seq_then_tree"] = thr[0]
        # vec4 with per-lane accumulators: acc_k = sum_j x[4j+k]; then combine seq
        acc = [torch.zeros_like(t[..., 0]) for _ in range(4)]
        for j in range(n // 4):
            for k in range(4):
                acc[k] = acc[k] + t[..., 4 * j + k]
        out["vec4_lane_acc_then_seq"] = ((acc[0] + acc[1]) + acc[2]) + acc[3]
    return out


def main():
    assert torch.cuda.is_available()
    ref = ref_gpu.load()
    b2d, oil = ref.box_intersection_2d, ref.oriented_iou_loss
    from nerf_rpn_b200 import ops
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__}
    g = torch.Generator().manual_seed(5)
    n = 400_000
    a, b = rand_obb(n, g, extent=14.0).cuda(), rand_obb(n, g, extent=14.0).cuda()
    for shape_name, view in (("B1xN", lambda t: t[None]), ("Nx1", lambda t: t[:, None])):
        A, B = view(a), view(b)
        iou_ref = oil.cal_iou_3d(A, B).reshape(-1)
        ours = ops.iou3d_pairs(a.contiguous(), b.contiguous())
        nz = iou_ref > 0
        res[f"iou_pairs_{shape_name}"] = {"n": n, "nonzero": int(nz.sum()), "bit_equal_all": biteq(iou_ref, ours),
                                          "bit_equal_nonzero": biteq(iou_ref[nz], ours[nz]),
                                          "max_abs_diff": float((iou_ref - ours).abs().max())}
    # ---- step 1: sin / cos
    alpha = a[:, 6]
    res["sin"] = {"torch_vs_fp64_rounded": biteq(torch.sin(alpha), alpha.double().sin().float()),
                  "cos_torch_vs_fp64_rounded": biteq(torch.cos(alpha), alpha.double().cos().float())}
    # ---- step 2: bmm of corners (B,N,4,2) x rot_T (2,2)
    box = a[None][..., [0, 1, 3, 4, 6]]
    corners_ref = oil.box2corners_th(box)[0]                       # (n,4,2)
    x, y, w, h, al = (box[0, :, i:i + 1] for i in range(5))
    x4 = torch.tensor([0.5, -0.5, -0.5, 0.5], device="cuda")[None] * w
    y4 = torch.tensor([0.5, 0.5, -0.5, -0.5], device="cuda")[None] * h
    s, c = torch.sin(al), torch.cos(al)
    ns = -s
    cand = {
        "mul_mul_add": (x4 * c + y4 * ns, x4 * s + y4 * c),
        "fma(y4,r1, x4*r0)": (fma32(y4, ns, x4 * c), fma32(y4, c, x4 * s)),
        "fma(x4,r0, y4*r1)": (fma32(x4, c, y4 * ns), fma32(x4, s, y4 * c)),
    }
    res["bmm"] = {}
    for k, (rx, ry) in cand.items():
        cx, cy = rx + x, ry + y
        res["bmm"][k] = {"x": biteq(cx, corners_ref[..., 0]), "y": biteq(cy, corners_ref[..., 1])}
    # ---- step 3: torch.sum orders on the real intermediate tensors
    A, B = a[None], b[None]
    c1 = oil.box2corners_th(A[..., [0, 1, 3, 4, 6]])
    c2 = oil.box2corners_th(B[..., [0, 1, 3, 4, 6]])
    inters, mask_inter = b2d.box_intersection_th(c1, c2)
    c12, c21 = b2d.box_in_box_th(c1, c2)
    vertices, mask = b2d.build_vertices(c1, c2, c12, c21, inters, mask_inter)
    num_valid = torch.sum(mask.int(), dim=2).int()
    masked = vertices * mask.float().unsqueeze(-1)                  # (1,n,24,2)
    sum_ref = torch.sum(masked, dim=2, keepdim=True)[0, :, 0]       # (n,2)
    keep = (num_valid[0] >= 3)
    res["sum24"] = {}
    for k, v in sum_candidates(masked[0].permute(0, 2, 1)).items():     # (n,2,24)
        res["sum24"][k] = biteq(v[keep], sum_ref[keep])
    mean = torch.sum(masked, dim=2, keepdim=True) / num_valid.unsqueeze(-1).unsqueeze(-1)
    vn = (vertices - mean).float()
    idx = ref.sort_vertices.sort_vertices_forward(vn.contiguous(), mask.contiguous(), num_valid.contiguous())
    ours_idx = ops.sort_vertices_forward(vn.contiguous(), mask.contiguous(), num_valid.contiguous())
    ok = (num_valid[0] <= 8)
    res["k1_vs_nrpn_sort_vertices"] = {"rows": int(ok.sum()), "rows_equal": float((idx[0][ok] == ours_idx[0][ok]).all(dim=-1).float().mean())}
    idx_ext = idx.long().unsqueeze(-1).repeat([1, 1, 1, 2])
    sel = torch.gather(vertices, 2, idx_ext)
    terms = sel[:, :, 0:-1, 0] * sel[:, :, 1:, 1] - sel[:, :, 0:-1, 1] * sel[:, :, 1:, 0]     # (1,n,8)
    tot_ref = torch.sum(terms, dim=2)[0]
    res["sum8"] = {k: biteq(v[keep], tot_ref[keep]) for k, v in sum_candidates(terms[0]).items()}
    # the shoelace terms themselves: a*b - c*d as separate ops or fused?
    t_sep = terms[0]
    t_fma = fma32(sel[0, :, 0:-1, 0], sel[0, :, 1:, 1], -(sel[0, :, 0:-1, 1] * sel[0, :, 1:, 0]))
    res["shoelace_term"] = {"separate_ops(ref by construction)": 1.0, "fma_variant_equal_to_ref": biteq(t_fma[keep], t_sep[keep])}
    # ---- NMS keep sets, reference loop vs nrpn_nms
    nms = {}
    for tag, nb, groups in (("2500x4_levels", 10000, 4), ("10000_single", 10000, 1), ("3000_dense", 3000, 1)):
        gg = torch.Generator().manual_seed(77 + nb + groups)
        boxes = rand_obb(nb, gg, extent=60.0 if tag != "3000_dense" else 25.0, smin=2.0, smax=14.0)
        scores = torch.rand(nb, generator=gg)
        lv = torch.randint(0, groups, (nb,), generator=gg)
        import time
        t0 = time.perf_counter()
        if groups == 1:
            k_ref = ref.utils.nms(boxes, scores, 0.3)
        else:
            k_ref = ref.utils.batched_nms(boxes, scores, lv, 0.3)
        t_ref = time.perf_counter() - t0
        t0 = time.perf_counter()
        k_our, n_our = ops.nms_device(boxes.cuda(), scores.cuda(), lv.to(torch.int32).cuda() if groups > 1 else None, 0.3)
        k_our = k_our[: int(n_our.item())].cpu()
        t_our = time.perf_counter() - t0
        same = k_ref.numel() == k_our.numel() and bool((k_ref == k_our).all())
        sym = len(set(k_ref.tolist()) ^ set(k_our.tolist()))
        nms[tag] = {"kept_ref": int(k_ref.numel()), "kept_ours": int(k_our.numel()), "identical": same, "symmetric_difference": sym,
                    "reference_seconds": round(t_ref, 3), "ours_seconds_incl_first_call": round(t_our, 4)}
    res["nms"] = nms
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_gpu_probe.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
