"""Shared model-construction recipe (TEST helper): rebuilds, from seeds only, exactly the weights that
tools/make_golden.py gave the reference when it produced tests/golden/rpn_small_*.npz."""
import numpy as np
import torch

ANCHOR_SIZES = ((8,), (16,), (32,), (64,),)
ASPECT = (((1., 1., 1.), (1., 1., 2.), (1., 2., 2.), (1., 1., 3.), (1., 3., 3.)),) * 4


def build_small_model(ns, rotated, golden):
    """ns: namespace providing ResNet_FPN_256, Bottleneck, AnchorGenerator3D, RPNHead (reference or ours)."""
    torch.manual_seed(0)
    backbone = ns.ResNet_FPN_256(ns.Bottleneck, [3, 4, 6, 3], input_dim=4, is_max_pool=True)
    ag = ns.AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
    head = ns.RPNHead(256, ag.num_anchors_per_location()[0], 4, rotate=rotated)
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
        # the generator calibrated both predictors on the reference's own forward; take the stored result
        head.cls_logits.weight.copy_(torch.from_numpy(golden["cls_w"]).view_as(head.cls_logits.weight))
        head.bbox_pred.weight.copy_(torch.from_numpy(golden["bbox_w"]).view_as(head.bbox_pred.weight))
    assert abs(backbone.conv1.weight.double().sum().item() - float(golden["conv1_sum"])) < 1e-9, "seeded weights differ"
    assert abs(head.conv[0].weight.double().sum().item() - float(golden["head_sum"])) < 1e-9, "seeded head weights differ"
    return backbone, ag, head


def golden_input(golden):
    return torch.from_numpy(golden["grid"]).permute(3, 0, 1, 2).contiguous()       # (4,W,L,H) like datasets.py:55-57


def build_vgg_small(ns, golden):
    """Config-1 model (VGG19 'EF' + FPN + anchor head, 32^3 grid) with the weights tools/make_golden.py gave the reference."""
    torch.manual_seed(0)
    backbone = ns.VGG_FPN("EF", 4, True, 32)
    ag = ns.AnchorGenerator3D(ANCHOR_SIZES, ASPECT)
    head = ns.RPNHead(256, ag.num_anchors_per_location()[0], 4, rotate=False)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        torch.randn(head.cls_logits.weight.shape, generator=g); torch.randn(head.bbox_pred.weight.shape, generator=g)   # same draws
        for m in backbone.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        head.cls_logits.weight.copy_(torch.from_numpy(golden["cls_w"]).view_as(head.cls_logits.weight))
        head.bbox_pred.weight.copy_(torch.from_numpy(golden["bbox_w"]).view_as(head.bbox_pred.weight))
    stem = list(backbone.layers.children())[0]
    assert abs(stem.weight.double().sum().item() - float(golden["stem_sum"])) < 1e-9, "seeded VGG weights differ"
    assert abs(backbone.fpn_neck.fpn_convs[3].weight.double().sum().item() - float(golden["fpn_sum"])) < 1e-9
    assert abs(head.conv[0].weight.double().sum().item() - float(golden["head_sum"])) < 1e-9
    return backbone, ag, head


def fcos_args(rotated, pre_n=2500, post_n=2500):
    import argparse
    return argparse.Namespace(num_convs=4, norm_reg_targets=True, centerness_on_reg=True, rotated_bbox=rotated, pre_nms_thresh=0.0,
                              pre_nms_top_n=pre_n, nms_thresh=0.3, fpn_post_nms_top_n=post_n, min_size=0.0,
                              center_sampling_radius=1.5, iou_loss_type="iou", use_additional_l1_loss=False, proj2d_loss_weight=0.0)


def build_fcos_small(ns, rotated, golden, pre_n=2500, post_n=2500):
    """FCOSOverNeRF(ResNet50-FPN) with the weights tools/make_golden.py:gen_fcos_small gave the reference."""
    torch.manual_seed(0)
    backbone = ns.ResNet_FPN_256(ns.Bottleneck, [3, 4, 6, 3], input_dim=4, is_max_pool=True)
    model = ns.FCOSOverNeRF(fcos_args(rotated, pre_n, post_n), backbone, [4, 8, 16, 32])
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
        head.cls_logits.weight.copy_(torch.from_numpy(golden["cls_w"])); head.cls_logits.bias.fill_(-1.0)
        head.bbox_pred.weight.copy_(torch.from_numpy(golden["bbox_w"])); head.bbox_pred.bias.fill_(1.0)
        head.centerness.weight.copy_(torch.from_numpy(golden["ctr_w"]))
    assert abs(backbone.conv1.weight.double().sum().item() - float(golden["conv1_sum"])) < 1e-9
    assert abs(head.cls_tower[0].weight.double().sum().item() - float(golden["tower_sum"])) < 1e-9, "seeded FCOS tower weights differ"
    return model


def seed1000_input(dims=(32, 48, 40)):
    gi = torch.Generator().manual_seed(1000)
    return torch.rand(*dims, 4, generator=gi).permute(3, 0, 1, 2).contiguous()


SWIN_S = dict(patch_size=[4, 4, 4], embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=[4, 4, 4],
              stochastic_depth_prob=0.1, expand_dim=True)


def build_swin_fcos_small(ns, golden):
    """FCOSOverNeRF(Swin-S + FPN), OBB, with the weights tools/make_golden.py:gen_swin_small gave the reference."""
    torch.manual_seed(0)
    backbone = ns.SwinTransformer_FPN(**SWIN_S)
    model = ns.FCOSOverNeRF(fcos_args(True), backbone, [4, 8, 16, 32])
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
        head.cls_logits.weight.copy_(torch.from_numpy(golden["cls_w"])); head.cls_logits.bias.fill_(-1.0)
        head.bbox_pred.weight.copy_(torch.from_numpy(golden["bbox_w"])); head.bbox_pred.bias.fill_(1.0)
        head.centerness.weight.copy_(torch.from_numpy(golden["ctr_w"]))
    assert abs(backbone.patch_partition[0].weight.double().sum().item() - float(golden["pe_sum"])) < 1e-9
    assert abs(backbone.stages[2][5].attn.qkv.weight.double().sum().item() - float(golden["qkv_sum"])) < 1e-9, "seeded Swin weights differ"
    assert abs(head.cls_tower[0].weight.double().sum().item() - float(golden["tower_sum"])) < 1e-9
    return model
