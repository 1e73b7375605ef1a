#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/nerf_rpn) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tools/make_golden.py

The reference has no tests or fixtures of its own, so these vectors -- outputs of
the reference's own Python run here -- are what pins the oracle and the CUDA path.  Two shims are needed
to import it on a GPU-less box: a CPU stand-in for the `sort_vertices` pybind
module (tools/ref_stub/sort_vertices.py) and a no-op Tensor.cuda().
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "ref_stub"))
sys.path.insert(0, "/root/reference/nerf_rpn")
torch.Tensor.cuda = lambda self, *a, **k: self
_is_avail = torch.cuda.is_available
torch.cuda.is_available = lambda: False

from model.anchor import AnchorGenerator3D, RPNHead  # noqa: E402
from model.coder import AABBCoder, MidpointOffsetCoder  # noqa: E402
from model.feature_extractor import Bottleneck, ResNet_FPN_256, SwinTransformer_FPN, VGG_FPN  # noqa: E402
from model.nerf_rpn import NeRFRegionProposalNetwork  # noqa: E402
from model.rotated_iou.box_intersection_2d import (box_in_box_th, box_intersection_th,  # noqa: E402
                                                   build_vertices)
from model.rotated_iou.oriented_iou_loss import box2corners_th, cal_iou_3d  # noqa: E402
from model.utils import batched_nms, box_iou_3d, nms  # noqa: E402
import sort_vertices as sv_stub  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

ANCHOR_SIZES = ((8,), (16,), (32,), (64,),)
ASPECT = (((1., 1., 1.), (1., 1., 2.), (1., 2., 2.), (1., 1., 3.), (1., 3., 3.)),) * 4


def rand_obb(n, gen, extent=30.0, smin=1.0, smax=11.0):
    return torch.cat([torch.rand(n, 3, generator=gen) * extent,
                      torch.rand(n, 3, generator=gen) * (smax - smin) + smin,
                      (torch.rand(n, 1, generator=gen) - 0.5) * math.pi], 1)


def rand_aabb(n, gen, extent=30.0, smin=0.5, smax=12.0):
    lo = torch.rand(n, 3, generator=gen) * extent
    return torch.cat([lo, lo + torch.rand(n, 3, generator=gen) * (smax - smin) + smin], 1)


def gen_iou():
    A = [0, 0, 0, 3, 3, 3, 0]
    B = [1, 1, 1, 2, 2, 2, math.pi / 3]
    C = [16.704862594604492, 4.111976623535156, 11.86344051361084, 7.168941974639893, 2.519232749938965,
         5.041347503662109, -0.7416284084320068]
    kat = [(A, B), (A, A), ([0, 0, 0, 2, 2, 2, 0], [1, 0, 0, 2, 2, 2, 0]),
           ([0, 0, 0, 2, 2, 2, 0], [0, 0, 0, 2, 2, 2, math.pi / 4]),
           ([0, 0, 0, 2, 2, 2, 0], [0, 0, 1, 2, 2, 2, math.pi / 4]), (C, C),
           ([0, 0, 0, 2, 2, 2, 0], [5, 5, 5, 1, 1, 1, 0.3]),            # disjoint
           ([0, 0, 0, 4, 4, 4, 0.2], [0.1, 0.1, 0, 1, 1, 1, 1.0]),      # contained
           ([0, 0, 0, 2, 2, 2, 0], [2, 0, 0, 2, 2, 2, 0])]              # touching faces
    a = torch.tensor([p[0] for p in kat], dtype=torch.float32)
    b = torch.tensor([p[1] for p in kat], dtype=torch.float32)
    kat_iou = cal_iou_3d(a[None], b[None])[0]
    g = torch.Generator().manual_seed(11)
    ra = rand_obb(4000, g, extent=12.0)
    rb = rand_obb(4000, g, extent=12.0)
    r_iou = cal_iou_3d(ra[None], rb[None])[0]
    g = torch.Generator().manual_seed(12)
    m = rand_obb(160, g, extent=20.0)
    mat = box_iou_3d(m, m)
    ga = rand_aabb(300, g)
    gmat = box_iou_3d(ga, ga)
    np.savez_compressed(os.path.join(OUT, "iou.npz"), kat_a=a.numpy(), kat_b=b.numpy(), kat_iou=kat_iou.numpy(),
                        rand_a=ra.numpy(), rand_b=rb.numpy(), rand_iou=r_iou.numpy(),
                        mat_boxes=m.numpy(), mat_iou=mat.numpy(), aabb_boxes=ga.numpy(), aabb_iou=gmat.numpy())
    print("iou.npz: kat", kat_iou.tolist(), "rand nonzero", int((r_iou > 0).sum()))


def gen_sort_vertices():
    """Inputs exactly as box_intersection_2d.py:161-174 would hand them to the native op."""
    g = torch.Generator().manual_seed(21)
    a = rand_obb(1500, g, extent=10.0)
    b = rand_obb(1500, g, extent=10.0)
    a[:50] = b[:50]                      # identical boxes -> the num_valid==8 special case
    b[50:100, 6] = a[50:100, 6]          # parallel edges
    c1 = box2corners_th(a[None][..., [0, 1, 3, 4, 6]])
    c2 = box2corners_th(b[None][..., [0, 1, 3, 4, 6]])
    inters, mask_inter = box_intersection_th(c1, c2)
    c12, c21 = box_in_box_th(c1, c2)
    vertices, mask = build_vertices(c1, c2, c12, c21, inters, mask_inter)
    num_valid = torch.sum(mask.int(), dim=2).int()
    mean = torch.sum(vertices * mask.float().unsqueeze(-1), dim=2, keepdim=True) / num_valid.unsqueeze(-1).unsqueeze(-1)
    vn = (vertices - mean).float()
    idx = sv_stub.sort_vertices_forward(vn, mask, num_valid)
    np.savez_compressed(os.path.join(OUT, "sort_vertices.npz"), vertices=vn.numpy(), mask=mask.numpy(),
                        num_valid=num_valid.numpy(), idx=idx.numpy())
    print("sort_vertices.npz: num_valid hist", np.bincount(num_valid.numpy().ravel()).tolist())


def gen_nms():
    out = {}
    torch.manual_seed(0)   # SURVEY.md section 8c vector
    boxes = torch.cat([torch.rand(64, 3) * 20, torch.rand(64, 3) * 6 + 2, (torch.rand(64, 1) - .5) * math.pi], 1)
    scores = torch.rand(64)
    lv = torch.randint(0, 4, (64,))
    out["s64_boxes"], out["s64_scores"], out["s64_levels"] = boxes.numpy(), scores.numpy(), lv.numpy()
    out["s64_keep"] = nms(boxes, scores, 0.3).numpy()
    out["s64_bkeep"] = batched_nms(boxes, scores, lv, 0.3).numpy()
    g = torch.Generator().manual_seed(31)
    boxes = rand_obb(700, g, extent=24.0, smin=2.0, smax=12.0)
    scores = torch.rand(700, generator=g)
    lv = torch.randint(0, 4, (700,), generator=g)
    out["o700_boxes"], out["o700_scores"], out["o700_levels"] = boxes.numpy(), scores.numpy(), lv.numpy()
    out["o700_keep"] = nms(boxes, scores, 0.3).numpy()
    out["o700_bkeep"] = batched_nms(boxes, scores, lv, 0.3).numpy()
    out["o700_keep_t5"] = nms(boxes, scores, 0.5).numpy()
    boxes = rand_aabb(1500, g, extent=40.0, smin=2.0, smax=16.0)
    scores = torch.rand(1500, generator=g)
    lv = torch.randint(0, 4, (1500,), generator=g)
    out["a1500_boxes"], out["a1500_scores"], out["a1500_levels"] = boxes.numpy(), scores.numpy(), lv.numpy()
    out["a1500_keep"] = nms(boxes, scores, 0.3).numpy()
    out["a1500_bkeep"] = batched_nms(boxes, scores, lv, 0.3).numpy()
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **out)
    print("nms.npz:", {k: len(v) for k, v in out.items() if "keep" in k})


def gen_decode_anchors():
    g = torch.Generator().manual_seed(41)
    ag = AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
    mesh = torch.zeros(1, 4, 32, 48, 40)
    feats = [torch.zeros(1, 256, 8, 12, 10), torch.zeros(1, 256, 4, 6, 5), torch.zeros(1, 256, 2, 3, 3),
             torch.zeros(1, 256, 1, 2, 2)]
    anchors, _ = ag(mesh, feats)
    cell = [c.numpy() for c in ag.cell_anchors]
    anc = anchors[0]
    n = anc.shape[0]
    d6 = torch.randn(n, 6, generator=g) * 0.5
    d6[::97, 3:] = 9.0       # exercise the exp clamp
    dec6 = AABBCoder().decode_single(d6, anc)
    d8 = torch.randn(n, 8, generator=g) * 0.5
    d8[::89, 3:6] = 6.0
    d8[::83, 6:] = 0.9
    dec7 = MidpointOffsetCoder().decode_single(d8, anc)
    np.savez_compressed(os.path.join(OUT, "decode.npz"), anchors=anc.numpy(), cell0=cell[0], cell1=cell[1],
                        cell2=cell[2], cell3=cell[3], d6=d6.numpy(), dec6=dec6.numpy(), d8=d8.numpy(),
                        dec7=dec7.numpy())
    print("decode.npz: anchors", tuple(anc.shape), "cell0", cell[0][:, 3:].tolist())


def gen_rpn_small():
    """Full reference forward (ResNet50-FPN + anchor head + RPN post-processing) on a small grid."""
    for rotated in (False, True):
        torch.manual_seed(0)
        backbone = ResNet_FPN_256(Bottleneck, [3, 4, 6, 3], input_dim=4, is_max_pool=True)
        ag = AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
        head = RPNHead(256, ag.num_anchors_per_location()[0], 4, rotate=rotated)
        # spread the objectness so top-k / NMS see a non-degenerate load
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            head.cls_logits.weight.copy_(torch.randn(head.cls_logits.weight.shape, generator=g) * 0.2)
            head.bbox_pred.weight.copy_(torch.randn(head.bbox_pred.weight.shape, generator=g) * 0.05)
            for m in backbone.modules():      # non-trivial BN statistics so the folding is exercised
                if isinstance(m, torch.nn.BatchNorm3d):
                    m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                    m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
                    m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        # calibrate the two predictors so logits have std ~2 and deltas std ~0.3 on this input
        gi = torch.Generator().manual_seed(1000)
        grid = torch.rand(32, 48, 40, 4, generator=gi)           # channels-last as on disk
        x = grid.permute(3, 0, 1, 2).contiguous()
        backbone.eval(); head.eval()
        with torch.no_grad():
            lg, dl = head(list(backbone(x[None])))
            s_l = torch.cat([t.flatten() for t in lg]).std().item()
            s_d = torch.cat([t.flatten() for t in dl]).std().item()
            head.cls_logits.weight.mul_(2.0 / s_l)
            head.bbox_pred.weight.mul_(0.3 / s_d)
        model = NeRFRegionProposalNetwork(backbone, ag, head, rpn_pre_nms_top_n_test=2500, rpn_post_nms_top_n_test=2500,
                                          rpn_nms_thresh=0.3, rpn_fg_iou_thresh=0.35, rpn_bg_iou_thresh=0.2,
                                          rpn_score_thresh=0.0, rotated_bbox=rotated)
        model.eval()
        with torch.no_grad():
            (features, proposals, level_index), _, scores = model([x])
            logits, deltas = head(features)
        chk = {k: float(v.double().sum()) for k, v in list(backbone.state_dict().items())[:3]}
        out = dict(grid=grid.numpy(), proposals=proposals[0].numpy(), scores=scores[0].numpy(),
                   level_index=level_index[0].numpy(),
                   conv1_sum=np.float64(backbone.conv1.weight.double().sum().item()),
                   head_sum=np.float64(head.conv[0].weight.double().sum().item()),
                   cls_w=head.cls_logits.weight.detach().numpy().reshape(13, 256),
                   bbox_w=head.bbox_pred.weight.detach().numpy().reshape(-1, 256))
        for i in range(4):
            out[f"feat{i}"] = features[i][0].numpy().astype(np.float16)
            out[f"logits{i}"] = logits[i][0].numpy()
            out[f"deltas{i}"] = deltas[i][0].numpy()
        name = "rpn_small_obb.npz" if rotated else "rpn_small_aabb.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print(name, "proposals", tuple(proposals[0].shape), "score range", float(scores[0].min()), float(scores[0].max()), chk)


def gen_vgg_small():
    """BASELINE config 1: one 32x32x32 grid, VGG19 ("EF") + FPN + anchor head (AABB), reference forward on CPU."""
    torch.manual_seed(0)
    backbone = VGG_FPN("EF", 4, True, 32)
    ag = AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
    head = RPNHead(256, ag.num_anchors_per_location()[0], 4, rotate=False)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        head.cls_logits.weight.copy_(torch.randn(head.cls_logits.weight.shape, generator=g) * 0.2)
        head.bbox_pred.weight.copy_(torch.randn(head.bbox_pred.weight.shape, generator=g) * 0.05)
        for m in backbone.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    gi = torch.Generator().manual_seed(1000)
    grid = torch.rand(32, 32, 32, 4, generator=gi)
    x = grid.permute(3, 0, 1, 2).contiguous()
    backbone.eval(); head.eval()
    with torch.no_grad():
        lg, dl = head(list(backbone(x[None])))
        head.cls_logits.weight.mul_(2.0 / torch.cat([t.flatten() for t in lg]).std().item())
        head.bbox_pred.weight.mul_(0.3 / torch.cat([t.flatten() for t in dl]).std().item())
    model = NeRFRegionProposalNetwork(backbone, ag, head, rpn_pre_nms_top_n_test=2500, rpn_post_nms_top_n_test=2500,
                                      rpn_nms_thresh=0.3, rpn_score_thresh=0.0, rotated_bbox=False)
    model.eval()
    with torch.no_grad():
        (features, proposals, level_index), _, scores = model([x])
        logits, deltas = head(features)
    out = dict(grid=grid.numpy(), proposals=proposals[0].numpy(), scores=scores[0].numpy(), level_index=level_index[0].numpy(),
               stem_sum=np.float64(list(backbone.layers.children())[0].weight.double().sum().item()),
               fpn_sum=np.float64(backbone.fpn_neck.fpn_convs[3].weight.double().sum().item()),
               head_sum=np.float64(head.conv[0].weight.double().sum().item()),
               cls_w=head.cls_logits.weight.detach().numpy().reshape(13, 256),
               bbox_w=head.bbox_pred.weight.detach().numpy().reshape(-1, 256))
    # sub-sampled to keep the fixture small: stride 4 on the 32^3 level, 2 on the 16^3 level (features); logits stride 2 / 1
    fstride, lstride = (4, 2, 1, 1), (2, 1, 1, 1)
    for i in range(4):
        fs, ls = fstride[i], lstride[i]
        out[f"feat{i}"] = features[i][0][:, ::fs, ::fs, ::fs].numpy().astype(np.float16)
        out[f"logits{i}"] = logits[i][0][:, ::ls, ::ls, ::ls].numpy()
    out["fstride"], out["lstride"] = np.array(fstride), np.array(lstride)
    np.savez_compressed(os.path.join(OUT, "vgg_small_aabb.npz"), **out)
    print("vgg_small_aabb.npz proposals", tuple(proposals[0].shape), [tuple(f.shape) for f in features], len(backbone.state_dict()))


def gen_fcos_small():
    """Anchor-free head (run_fcos.py / test_fcos.sh flags: --norm_reg_targets --centerness_on_reg): reference FCOSOverNeRF
    forward on CPU, ResNet50-FPN backbone, 32x48x40 grid; AABB and OBB, default and tight top-n settings."""
    import argparse
    from model.fcos.fcos import FCOSOverNeRF
    for rotated in (False, True):
        for tag, pre_n, post_n in (("", 2500, 2500), ("_tight", 300, 150)):
            args = argparse.Namespace(num_convs=4, norm_reg_targets=True, centerness_on_reg=True, rotated_bbox=rotated,
                                      pre_nms_thresh=0.0, pre_nms_top_n=pre_n, nms_thresh=0.3, fpn_post_nms_top_n=post_n,
                                      min_size=0.0, center_sampling_radius=1.5, iou_loss_type="iou",
                                      use_additional_l1_loss=False, proj2d_loss_weight=0.0)
            torch.manual_seed(0)
            backbone = ResNet_FPN_256(Bottleneck, [3, 4, 6, 3], input_dim=4, is_max_pool=True)
            model = FCOSOverNeRF(args, backbone, [4, 8, 16, 32])
            head = model.fcos_module.head
            g = torch.Generator().manual_seed(7)
            with torch.no_grad():
                for m in model.modules():
                    if isinstance(m, (torch.nn.BatchNorm3d, torch.nn.GroupNorm)):
                        if isinstance(m, torch.nn.BatchNorm3d):
                            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
                        m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                        m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                for i, sc in enumerate(head.scales):
                    sc.scale.fill_(1.0 + 0.1 * i)
            gi = torch.Generator().manual_seed(1000)
            grid = torch.rand(32, 48, 40, 4, generator=gi)
            x = grid.permute(3, 0, 1, 2).contiguous()
            model.eval()
            with torch.no_grad():
                feats = list(backbone(x[None]))
                # calibrate the three predictors on the towers' outputs: logits std 2, distances ~ a few voxels, centerness std 1
                ct = [head.cls_tower(f) for f in feats]; bt = [head.bbox_tower(f) for f in feats]
                s_c = torch.cat([head.cls_logits(t).flatten() for t in ct]).std().item()
                s_b = torch.cat([head.bbox_pred(t).flatten() for t in bt]).std().item()
                s_t = torch.cat([head.centerness(t).flatten() for t in bt]).std().item()
                head.cls_logits.weight.mul_(2.0 / s_c); head.cls_logits.bias.fill_(-1.0)
                head.bbox_pred.weight.mul_(1.5 / s_b); head.bbox_pred.bias.fill_(1.0)
                head.centerness.weight.mul_(1.0 / s_t)
                boxes, losses, scores = model([x])
                logits, bbox_reg, ctr = head(feats)
            out = dict(boxes=boxes[0].numpy(), scores=scores[0].numpy(),      # grid = torch.rand(32,48,40,4, seed 1000), not stored
                       conv1_sum=np.float64(backbone.conv1.weight.double().sum().item()),
                       tower_sum=np.float64(head.cls_tower[0].weight.double().sum().item()),
                       cls_w=head.cls_logits.weight.detach().numpy(), bbox_w=head.bbox_pred.weight.detach().numpy(),
                       ctr_w=head.centerness.weight.detach().numpy())
            for i in range(4):
                out[f"logits{i}"] = logits[i][0].numpy(); out[f"reg{i}"] = bbox_reg[i][0].numpy(); out[f"ctr{i}"] = ctr[i][0].numpy()
            name = f"fcos_small_{'obb' if rotated else 'aabb'}{tag}.npz"
            np.savez_compressed(os.path.join(OUT, name), **out)
            print(name, "boxes", tuple(boxes[0].shape), "scores", float(scores[0].min()), float(scores[0].max()))


SWIN_S = dict(patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=[4, 4, 4],
              stochastic_depth_prob=0.1, expand_dim=True)


def gen_swin_small():
    """BASELINE config 3 at a small size: Swin-S 3-D window-attention backbone + FPN + FCOS head (OBB), reference forward on CPU.
    Grid 40x52x34 -> token grid 10x13x8 (padded to 12x16x8 inside the attention: exercises padding, shift and masks)."""
    import argparse
    from model.fcos.fcos import FCOSOverNeRF
    args = argparse.Namespace(num_convs=4, norm_reg_targets=True, centerness_on_reg=True, rotated_bbox=True, pre_nms_thresh=0.0,
                              pre_nms_top_n=2500, nms_thresh=0.3, fpn_post_nms_top_n=2500, min_size=0.0, center_sampling_radius=1.5,
                              iou_loss_type="iou", use_additional_l1_loss=False, proj2d_loss_weight=0.0)
    torch.manual_seed(0)
    backbone = SwinTransformer_FPN(**SWIN_S)
    model = FCOSOverNeRF(args, backbone, [4, 8, 16, 32])
    head = model.fcos_module.head
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (torch.nn.LayerNorm, torch.nn.GroupNorm)):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        for name, p in backbone.named_parameters():
            if name.endswith("relative_position_bias_table"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            if name.endswith("qkv.bias") or name.endswith("proj.bias") or name.endswith("mlp.0.bias") or name.endswith("mlp.3.bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        for i, sc in enumerate(head.scales):
            sc.scale.fill_(1.0 + 0.1 * i)
    gi = torch.Generator().manual_seed(1000)
    grid = torch.rand(40, 52, 34, 4, generator=gi)
    x = grid.permute(3, 0, 1, 2).contiguous()
    model.eval()
    with torch.no_grad():
        feats = list(backbone(x[None]))
        ct = [head.cls_tower(f) for f in feats]; bt = [head.bbox_tower(f) for f in feats]
        s_c = torch.cat([head.cls_logits(t).flatten() for t in ct]).std().item()
        s_b = torch.cat([head.bbox_pred(t).flatten() for t in bt]).std().item()
        s_t = torch.cat([head.centerness(t).flatten() for t in bt]).std().item()
        head.cls_logits.weight.mul_(2.0 / s_c); head.cls_logits.bias.fill_(-1.0)
        head.bbox_pred.weight.mul_(1.5 / s_b); head.bbox_pred.bias.fill_(1.0)
        head.centerness.weight.mul_(1.0 / s_t)
        boxes, _, scores = model([x])
        logits, bbox_reg, ctr = head(feats)
    out = dict(boxes=boxes[0].numpy(), scores=scores[0].numpy(),
               pe_sum=np.float64(backbone.patch_partition[0].weight.double().sum().item()),
               qkv_sum=np.float64(backbone.stages[2][5].attn.qkv.weight.double().sum().item()),
               tower_sum=np.float64(head.cls_tower[0].weight.double().sum().item()),
               cls_w=head.cls_logits.weight.detach().numpy(), bbox_w=head.bbox_pred.weight.detach().numpy(),
               ctr_w=head.centerness.weight.detach().numpy())
    for i in range(4):
        out[f"feat{i}"] = feats[i][0].numpy().astype(np.float16)
        out[f"logits{i}"] = logits[i][0].numpy()
    np.savez_compressed(os.path.join(OUT, "swin_small_fcos_obb.npz"), **out)
    print("swin_small_fcos_obb.npz boxes", tuple(boxes[0].shape), [tuple(f.shape) for f in feats], len(backbone.state_dict()))


def recall_scene(i, dims=(32, 48, 40), n_gt=24):
    """Scene i of the recall fixture: U[0,1) grid (seed 2000+i) and n_gt random OBB ground-truth boxes inside it."""
    gi = torch.Generator().manual_seed(2000 + i)
    grid = torch.rand(*dims, 4, generator=gi)
    d = torch.tensor(dims, dtype=torch.float32)
    size = torch.rand(n_gt, 3, generator=gi) * 16.0 + 4.0
    ctr = torch.rand(n_gt, 3, generator=gi) * (d - 4.0) + 2.0
    theta = (torch.rand(n_gt, 1, generator=gi) - 0.5) * math.pi
    return grid.permute(3, 0, 1, 2).contiguous(), torch.cat([ctr, size, theta], 1)


def gen_recall(n_scenes=12):
    """recall@{0.25,0.5} (eval.py:14-81, limits 300/1000/2500) of the reference's OBB proposals on 12 small scenes with
    planted ground truth; weights = the rpn_small_obb model. Also pins the greedy matching itself on random IoU matrices."""
    import eval as ref_eval
    import types
    from tests import recipes
    g = np.load(os.path.join(OUT, "rpn_small_obb.npz"))
    ns = types.SimpleNamespace(ResNet_FPN_256=ResNet_FPN_256, Bottleneck=Bottleneck, AnchorGenerator3D=AnchorGenerator3D, RPNHead=RPNHead)
    backbone, ag, head = recipes.build_small_model(ns, True, g)
    model = NeRFRegionProposalNetwork(backbone, ag, head, rpn_pre_nms_top_n_test=2500, rpn_post_nms_top_n_test=2500,
                                      rpn_nms_thresh=0.3, rpn_fg_iou_thresh=0.35, rpn_bg_iou_thresh=0.2,
                                      rpn_score_thresh=0.0, rotated_bbox=True)
    model.eval()
    props, scores, gts = [], [], []
    for i in range(n_scenes):
        x, gt = recall_scene(i)
        with torch.no_grad():
            (_, proposals, _), _, sc = model([x])
        props.append(proposals[0]); scores.append(sc[0]); gts.append(gt)
        print("recall scene", i, tuple(proposals[0].shape))
    out = dict(n_scenes=np.int64(n_scenes), gt=torch.stack(gts).numpy(), n_props=np.array([p.shape[0] for p in props]))
    thr = torch.tensor([0.25, 0.5])
    for limit in (300, 1000, 2500):
        r = ref_eval.evaluate_box_proposals_recall(props, scores, gts, thresholds=thr, limit=limit)
        out[f"recalls_{limit}"] = r["recalls"].numpy()
        out[f"gt_overlaps_{limit}"] = r["gt_overlaps"].numpy()
        print("limit", limit, "recall@0.25/0.5", r["recalls"].tolist(), "num_pos", r["num_pos"])
    # VOC-style AP (eval.py:319-395) on the same proposals; the proposals themselves are stored so that the device metric can be
    # checked on identical inputs
    for thr_iou in (0.25, 0.5):
        for top_k in (None, 300):
            r = ref_eval.evaluate_box_proposals_ap(props, scores, gts, iou_thresh=thr_iou, top_k=top_k)
            out[f"ap_{int(thr_iou * 100)}_{top_k or 0}"] = np.float64(float(r["ap"]))
            print("AP", thr_iou, top_k, float(r["ap"]))
    out["ref_props"] = np.concatenate([p.numpy() for p in props]).astype(np.float32)
    out["ref_scores"] = np.concatenate([s_.numpy() for s_ in scores]).astype(np.float32)
    # the matching loop alone, on random matrices (ties included), through the reference function with a patched IoU
    gen = torch.Generator().manual_seed(5)
    mats, outs = [], []
    orig = ref_eval.box_iou_3d
    try:
        for (p, q) in ((40, 7), (5, 9), (300, 24), (1, 1)):
            m = torch.rand(p, q, generator=gen)
            m[m < 0.3] = 0.0
            m = (m * 8).round() / 8                    # many exact ties
            ref_eval.box_iou_3d = lambda a, b, m=m: m.clone()
            r = ref_eval.evaluate_box_proposals_recall([torch.zeros(p, 7)], [torch.arange(p, 0, -1).float()], [torch.zeros(q, 7)],
                                                       thresholds=thr, limit=None)
            mats.append(m.numpy()); outs.append(r["gt_overlaps"].numpy())
    finally:
        ref_eval.box_iou_3d = orig
    for k, (m, o) in enumerate(zip(mats, outs)):
        out[f"match_m{k}"] = m; out[f"match_o{k}"] = o
    np.savez_compressed(os.path.join(OUT, "recall_small_obb.npz"), **out)


def gen_recall_large(n_scenes=32, n_gt=64):
    """The recall fixture at a size that resolves north_star's 0.5 pt: 32 scenes x 64 planted OBBs = 2048 ground-truth boxes
    (one box = 0.05 pt), same model / scene recipe as gen_recall (scene seeds 2000+i, n_gt boxes each)."""
    import eval as ref_eval
    import types
    from tests import recipes
    g = np.load(os.path.join(OUT, "rpn_small_obb.npz"))
    ns = types.SimpleNamespace(ResNet_FPN_256=ResNet_FPN_256, Bottleneck=Bottleneck, AnchorGenerator3D=AnchorGenerator3D, RPNHead=RPNHead)
    backbone, ag, head = recipes.build_small_model(ns, True, g)
    model = NeRFRegionProposalNetwork(backbone, ag, head, rpn_pre_nms_top_n_test=2500, rpn_post_nms_top_n_test=2500,
                                      rpn_nms_thresh=0.3, rpn_fg_iou_thresh=0.35, rpn_bg_iou_thresh=0.2,
                                      rpn_score_thresh=0.0, rotated_bbox=True)
    model.eval()
    props, scores, gts = [], [], []
    for i in range(n_scenes):
        x, gt = recall_scene(i, n_gt=n_gt)
        with torch.no_grad():
            (_, proposals, _), _, sc = model([x])
        props.append(proposals[0]); scores.append(sc[0]); gts.append(gt)
        print("recall-large scene", i, tuple(proposals[0].shape), flush=True)
    out = dict(n_scenes=np.int64(n_scenes), n_gt=np.int64(n_gt), n_props=np.array([p.shape[0] for p in props]))
    thr = torch.tensor([0.25, 0.5])
    for limit in (300, 1000, 2500):
        r = ref_eval.evaluate_box_proposals_recall(props, scores, gts, thresholds=thr, limit=limit)
        out[f"recalls_{limit}"] = r["recalls"].numpy()
        print("limit", limit, "recall@0.25/0.5", r["recalls"].tolist(), "num_pos", r["num_pos"], flush=True)
    np.savez_compressed(os.path.join(OUT, "recall_large_obb.npz"), **out)


def gen_targets():
    """assign_targets_to_anchors (rpn.py:240-290) pieces run through the reference's own functions on CPU: obb2hbb_3d,
    batched_box_iou (chunks of 16 GT), Matcher(0.35, 0.2, allow_low_quality_matches=True) and the label mapping."""
    from model.coder.misc import obb2hbb_3d
    from model.utils import Matcher, batched_box_iou
    dims = (32, 48, 40)
    ag = AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
    feats = [torch.zeros(1, 1, *[(d + s - 1) // s for d in dims]) for s in (4, 8, 16, 32)]
    anchors = ag(torch.zeros(1, 4, *dims), feats)[0][0]
    matcher = Matcher(0.35, 0.2, allow_low_quality_matches=True)
    out = dict(anchors_sum=np.float64(anchors.double().sum().item()), n_anchors=np.int64(anchors.shape[0]))
    gen = torch.Generator().manual_seed(11)
    mask = torch.rand(anchors.shape[0], generator=gen) > 0.1            # stands in for get_padding_masks (anchor.py:124-152)
    for tag, n_gt in (("a", 24), ("b", 3)):
        _, gt = recall_scene(7 if tag == "a" else 8, dims, n_gt)
        if tag == "b":
            gt[2, :3] = torch.tensor([500.0, 500.0, 500.0])              # a ground-truth box no anchor touches: the IoU == max == 0 tie quirk
        for kind in ("obb", "aabb"):
            g = gt if kind == "obb" else obb2hbb_3d(gt)
            gq = obb2hbb_3d(g) if kind == "obb" else g
            for use_mask in (False, True):
                m = batched_box_iou(gq, anchors, 16)
                if use_mask:
                    m[:, ~mask] = -1.0
                idx = matcher(m)
                labels = (idx >= 0).float()
                labels[idx == -1] = 0.0
                labels[idx == -2] = -1.0
                if use_mask:
                    labels[~mask] = -1.0
                key = f"{tag}_{kind}_{int(use_mask)}"
                out["gt_" + key] = g.numpy()
                out["labels_" + key] = labels.numpy().astype(np.int8)
                out["matched_" + key] = idx.numpy().astype(np.int16)
                print(key, "pos", int((labels == 1).sum()), "neg", int((labels == 0).sum()), "ignored", int((labels == -1).sum()))
    out["mask"] = mask.numpy()
    np.savez_compressed(os.path.join(OUT, "targets_small.npz"), **out)


def gen_losses():
    """Loss side of the RPN training step through the reference's own functions (CPU): sampler with recorded randperm draws,
    AABB / midpoint-offset encoders on the matched pairs, smooth-L1 + BCE as combined in compute_loss (rpn.py:394-417)."""
    import torch.nn.functional as F
    from model.coder.AABB_coder import encode_boxes_3d
    from model.coder.midpoint_offset_coder import bbox2delta_sp
    from model.coder.misc import obb2hbb_3d
    from model.utils import BalancedPositiveNegativeSampler
    t = np.load(os.path.join(OUT, "targets_small.npz"))
    dims = (32, 48, 40)
    ag = AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
    feats = [torch.zeros(1, 1, *[(d + s - 1) // s for d in dims]) for s in (4, 8, 16, 32)]
    anchors = ag(torch.zeros(1, 4, *dims), feats)[0][0]
    out = {}
    gen = torch.Generator().manual_seed(21)
    for kind, code in (("aabb", 6), ("obb", 8)):
        gt = torch.from_numpy(t[f"gt_a_{kind}_0"])
        labels = torch.from_numpy(t[f"labels_a_{kind}_0"].astype(np.float32))
        matched = torch.from_numpy(t[f"matched_a_{kind}_0"].astype(np.int64))
        mgt = gt[matched.clamp(min=0)]
        targets = encode_boxes_3d(mgt, anchors) if kind == "aabb" else bbox2delta_sp(anchors, mgt)
        # the sampler draws two randperms from the global RNG: record them
        torch.manual_seed(5)
        n_pos, n_neg = int((labels >= 1).sum()), int((labels == 0).sum())
        perm_pos, perm_neg = torch.randperm(n_pos), torch.randperm(n_neg)
        torch.manual_seed(5)
        pos_m, neg_m = BalancedPositiveNegativeSampler(256, 0.5)([labels])
        pos_m, neg_m = pos_m[0].bool(), neg_m[0].bool()
        objectness = torch.randn(anchors.shape[0], generator=gen)
        deltas = torch.randn(anchors.shape[0], code, generator=gen) * 0.3
        pos, neg = torch.where(pos_m)[0], torch.where(neg_m)[0]
        sampled = torch.cat([pos, neg])
        box = F.smooth_l1_loss(deltas[pos], targets[pos], beta=1 / 9, reduction="sum") / sampled.numel()
        obj = F.binary_cross_entropy_with_logits(objectness[sampled], labels[sampled])
        finite = torch.isfinite(targets).all(dim=1)
        out.update({f"{kind}_perm_pos": perm_pos.numpy(), f"{kind}_perm_neg": perm_neg.numpy(), f"{kind}_pos_mask": pos_m.numpy(),
                    f"{kind}_neg_mask": neg_m.numpy(), f"{kind}_targets_pos": targets[pos].numpy(), f"{kind}_pos_idx": pos.numpy(),
                    f"{kind}_objectness_sampled": objectness[sampled].numpy(), f"{kind}_deltas_pos": deltas[pos].numpy(),
                    f"{kind}_neg_idx": neg.numpy(),
                    f"{kind}_loss_obj": np.float64(float(obj)), f"{kind}_loss_box": np.float64(float(box))})
        print(kind, "sampled", int(pos_m.sum()), int(neg_m.sum()), "losses", float(obj), float(box), "finite targets", int(finite.sum()))
    np.savez_compressed(os.path.join(OUT, "loss_small.npz"), **out)


FCOS_LOSS_CASES = (  # name, rotated, iou_loss_type, center_sampling_radius, additional_l1, batch
    ("aabb_iou", False, "iou", 1.5, False, 2), ("aabb_giou", False, "giou", 1.5, False, 2), ("aabb_linear", False, "linear_iou", 0.0, False, 1),
    ("aabb_sl1", False, "smooth_l1", 1.5, False, 2), ("obb_sl1", True, "smooth_l1", 1.5, False, 2), ("obb_iou_l1", True, "iou", 1.5, True, 2),
    ("obb_nocs", True, "smooth_l1", 0.0, False, 1), ("aabb_empty", False, "iou", 1.5, False, 2),
    ("obb_sl1_p2d", True, "smooth_l1", 1.5, False, 2, 0.5), ("obb_iou_p2d", True, "iou", 1.5, True, 1, 0.25))


def fcos_loss_inputs(name, rotated, batch, seed):
    """Synthetic head outputs + ground truth of one case (shared by tools/make_golden.py and the tests through the stored arrays)."""
    g = torch.Generator().manual_seed(seed)
    mesh = (64, 80, 48)
    strides = [4, 8, 16, 32]
    grids = [tuple(int(math.ceil(m / s)) for m in mesh) for s in strides]
    D = 8 if rotated else 6
    cls = [torch.randn(batch, 1, *gr, generator=g) * 2 - 2 for gr in grids]
    reg = [torch.cat([torch.rand(batch, 6, *gr, generator=g) * 3 + 0.1] + ([torch.randn(batch, 2, *gr, generator=g) * 0.3] if rotated else []), 1)
           for gr in grids]
    ctr = [torch.randn(batch, 1, *gr, generator=g) for gr in grids]
    sizes = [mesh, (52, 80, 40)][:batch]
    gts = []
    for b in range(batch):
        n = 0 if name.endswith("empty") and b == 1 else 9
        sz = torch.tensor(sizes[b], dtype=torch.float32)
        ext = torch.rand(n, 3, generator=g) * torch.tensor([60.0, 60.0, 40.0]) + 5.0                  # sizes 5 .. 65: every level's range is hit
        ctrs = torch.rand(n, 3, generator=g) * sz
        if rotated:
            gts.append(torch.cat([ctrs, ext, (torch.rand(n, 1, generator=g) - 0.5) * math.pi], 1))
        else:
            gts.append(torch.cat([ctrs - ext / 2, ctrs + ext / 2], 1))
    if rotated and batch > 0 and len(gts[0]) > 1:
        gts[0][0, 6] = 0.0                                                                            # the "theta too small" branch of encode_fcos_obb
        gts[0][1, 6] = 1e-4
    return strides, grids, sizes, cls, reg, ctr, gts


def gen_fcos_loss():
    """FCOSLossComputation of the unmodified reference (fcos/loss.py) on CPU: targets (level first), the three losses and their gradients w.r.t. the
    head outputs, AABB and OBB heads, every loss type that runs on a GPU-less box (the rotated IoU goes through the stub vertex sort)."""
    import argparse
    from model.fcos.fcos import FCOSModule
    out = {}
    for ci, case in enumerate(FCOS_LOSS_CASES):
        name, rotated, loss_type, radius, add_l1, batch = case[:6]
        proj2d = case[6] if len(case) > 6 else 0.0
        strides, grids, sizes, cls, reg, ctr, gts = fcos_loss_inputs(name, rotated, batch, 100 + ci)
        args = argparse.Namespace(num_convs=1, norm_reg_targets=True, centerness_on_reg=True, rotated_bbox=rotated, pre_nms_thresh=0.0,
                                  pre_nms_top_n=100, nms_thresh=0.3, fpn_post_nms_top_n=100, min_size=0.0, center_sampling_radius=radius,
                                  iou_loss_type=loss_type, use_additional_l1_loss=add_l1, proj2d_loss_weight=proj2d)
        mod = FCOSModule(args, 256, strides)
        locations = mod.compute_locations(cls)
        masks = mod.compute_padding_masks(locations, sizes) if batch > 1 else None
        for t in cls + reg + ctr:
            t.requires_grad_(True)
        ev = mod.loss_evaluator
        labels, reg_targets = ev.prepare_targets(locations, [t.clone() for t in gts])
        l_cls, l_reg, l_ctr = ev(locations, cls, reg, ctr, gts, masks)
        (l_cls + 2.0 * l_reg + 3.0 * l_ctr).backward()
        out[f"{name}/losses"] = np.array([l_cls.item(), l_reg.item(), l_ctr.item()], np.float64)
        for l in range(4):
            out[f"{name}/cls{l}"] = cls[l].detach().numpy(); out[f"{name}/reg{l}"] = reg[l].detach().numpy(); out[f"{name}/ctr{l}"] = ctr[l].detach().numpy()
            out[f"{name}/dcls{l}"] = cls[l].grad.numpy(); out[f"{name}/dreg{l}"] = reg[l].grad.numpy(); out[f"{name}/dctr{l}"] = ctr[l].grad.numpy()
            out[f"{name}/labels{l}"] = labels[l].numpy(); out[f"{name}/reg_targets{l}"] = reg_targets[l].numpy()
            if masks is not None:
                out[f"{name}/mask{l}"] = masks[l].numpy()
        for b in range(batch):
            out[f"{name}/gt{b}"] = gts[b].numpy()
        out[f"{name}/sizes"] = np.array(sizes, np.int64)
        print(name, "losses", out[f"{name}/losses"], "positives", int(sum((lb > 0).sum() for lb in labels)))
    from model.fcos.utils import decode_fcos_obb
    g = torch.Generator().manual_seed(77)
    loc = torch.rand(3000, 3, generator=g) * 100
    reg = torch.cat([torch.rand(3000, 6, generator=g) * 12 + 0.05, torch.randn(3000, 2, generator=g) * 0.35], 1)
    reg[:40, 6:] = 0.0                                                       # vertices at the edge midpoints: theta = pi/4 exactly-ish
    reg[40:80, 6] = 0.5; reg[40:80, 7] = -0.5                                # clamped to the AABB corner: an axis-aligned box
    out["decode/loc"], out["decode/reg"], out["decode/boxes"] = loc.numpy(), reg.numpy(), decode_fcos_obb(loc, reg).numpy()
    np.savez_compressed(os.path.join(OUT, "fcos_loss.npz"), **out)


def gen_proj2d():
    """loss_rpn_box_reg_2d through the reference's own RegionProposalNetwork.compute_loss (rpn.py:372-456) on CPU, both heads: value and gradient w.r.t. the
    decoded boxes (the path the 2-D projection loss back-propagates through, rpn.py:516-518), plus the coders' decode of the deltas."""
    from model.coder import AABBCoder, MidpointOffsetCoder
    from model.rpn import RegionProposalNetwork
    out = {}
    for kind, rotated, res in (("aabb", False, 200), ("obb", True, 160)):
        g = torch.Generator().manual_seed(31 + int(rotated))
        n, n_pos, n_neg = 600, 48, 300
        ag = AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
        rpn = RegionProposalNetwork(ag, RPNHead(256, 13, 1, rotate=rotated), 0.35, 0.2, 256, 0.5, dict(training=100, testing=100),
                                    dict(training=100, testing=100), 0.3, rotated_bbox=rotated)
        c, half = torch.rand(n, 3, generator=g) * res, torch.rand(n, 3, generator=g) * 20 + 2
        anchors = torch.cat([c - half, c + half], 1)
        deltas = (torch.randn(n, 8 if rotated else 6, generator=g) * 0.3).requires_grad_(True)
        pred = rpn.box_coder.decode_single(deltas, anchors)
        pred_leaf = pred.detach().clone().requires_grad_(True)
        if rotated:
            tgt = torch.cat([pred.detach()[:, :3] + torch.randn(n, 3, generator=g) * 3, pred.detach()[:, 3:6] * (torch.rand(n, 3, generator=g) + 0.5),
                             (torch.rand(n, 1, generator=g) - 0.5) * math.pi], 1)
        else:
            tgt = pred.detach() + torch.randn(n, 6, generator=g) * 3
        labels = torch.full((n,), -1.0)
        labels[:n_pos] = 1.0; labels[n_pos:n_pos + n_neg] = 0.0
        torch.manual_seed(3)
        l_obj, l_3d, l_2d = rpn.compute_loss(torch.randn(n, 1, generator=g), deltas.detach(), [labels], [torch.zeros(n, deltas.shape[1])], pred_leaf, [tgt], res)
        l_2d.backward()
        out.update({f"{kind}_anchors": anchors.numpy(), f"{kind}_deltas": deltas.detach().numpy(), f"{kind}_decoded": pred.detach().numpy(),
                    f"{kind}_target": tgt.numpy(), f"{kind}_loss_2d": np.float64(l_2d.item()), f"{kind}_dpred": pred_leaf.grad.numpy(),
                    f"{kind}_n_pos": np.int64(n_pos), f"{kind}_res": np.int64(res)})
        print("proj2d", kind, "loss_2d", l_2d.item(), "grad norm", float(pred_leaf.grad.norm()))
    np.savez_compressed(os.path.join(OUT, "proj2d.npz"), **out)


if __name__ == "__main__":
    if "--proj2d-only" in sys.argv:
        gen_proj2d(); sys.exit(0)
    if "--fcos-loss-only" in sys.argv:
        gen_fcos_loss(); sys.exit(0)
    if "--losses-only" in sys.argv:
        sys.path.insert(0, ROOT); gen_losses(); sys.exit(0)
    if "--targets-only" in sys.argv:
        sys.path.insert(0, ROOT); gen_targets(); sys.exit(0)
    if "--recall-large-only" in sys.argv:
        sys.path.insert(0, ROOT); gen_recall_large(); sys.exit(0)
    if "--recall-only" in sys.argv:
        sys.path.insert(0, ROOT); gen_recall(); sys.exit(0)
    if "--swin-only" in sys.argv:
        gen_swin_small(); sys.exit(0)
    if "--fcos-only" in sys.argv:
        gen_fcos_small(); sys.exit(0)
    gen_vgg_small()
    sys.exit(0) if "--vgg-only" in sys.argv else None
    gen_iou()
    gen_sort_vertices()
    gen_nms()
    gen_decode_anchors()
    gen_rpn_small()
    gen_fcos_small()
    gen_swin_small()
    sys.path.insert(0, ROOT); gen_recall()
    gen_targets()
    gen_losses()
    gen_fcos_loss()
    gen_proj2d()
