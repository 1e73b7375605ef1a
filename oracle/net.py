"""oracle/net.py -- TEST INFRASTRUCTURE ONLY.

fp32 CPU restatement (torch.nn.functional, NCDHW) of the reference network forward for the hot path, driven by a
state_dict with the reference's key names:
  ResNet_FPN_256.forward / Bottleneck.forward   nerf_rpn/model/feature_extractor.py:48-68,215-235
  RPNHead.forward                               nerf_rpn/model/anchor.py:206-213
  concat_box_prediction_layers ordering         nerf_rpn/model/rpn.py:20-27,105-130
Pinned by tests/golden/rpn_small_*.npz (feature maps, logits, deltas and proposals of the unmodified reference).
It is also the "port" CPU baseline timed by bench.py (cpu_baseline / --impl reference).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import rpn_post as rp


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)


def _stage_blocks(sd):
    n = {}
    for k in sd:
        if k.startswith("layers."):
            s, b = int(k.split(".")[1]), int(k.split(".")[2])
            n[s] = max(n.get(s, 0), b + 1)
    return [n[s] for s in sorted(n)]


@torch.no_grad()
def resnet_fpn_forward(sd, x):
    """x (N,4,W,L,H) fp32 -> [P2,P3,P4,P5] (N,256,...) fp32."""
    c1 = F.relu(_bn(F.conv3d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1"))
    c1 = F.max_pool3d(c1, kernel_size=3, stride=2, padding=1)
    c_out = [c1]
    for s, nb in enumerate(_stage_blocks(sd)):
        h = c_out[-1]
        for b in range(nb):
            p = f"layers.{s}.{b}"
            stride = 2 if (b == 0 and s > 0) else 1
            out = F.relu(_bn(F.conv3d(h, sd[p + ".conv1.weight"], stride=stride), sd, p + ".bn1"))
            out = F.relu(_bn(F.conv3d(out, sd[p + ".conv2.weight"], padding=1), sd, p + ".bn2"))
            out = _bn(F.conv3d(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            res = h
            if p + ".downsample.0.weight" in sd:
                res = _bn(F.conv3d(h, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
            h = F.relu(out + res)
        c_out.append(h)
    nl = len(c_out) - 1
    p_out = [F.conv3d(c_out[-1], sd["latlayers.0.weight"], sd["latlayers.0.bias"])]
    for i in range(nl - 1):
        lat = F.conv3d(c_out[-2 - i], sd[f"latlayers.{i + 1}.weight"], sd[f"latlayers.{i + 1}.bias"])
        p_out.append(F.interpolate(p_out[i], size=lat.shape[-3:], mode="nearest") + lat)
    for i in range(nl - 1):
        p_out[i + 1] = F.conv3d(p_out[i + 1], sd[f"smooths.{i}.weight"], sd[f"smooths.{i}.bias"], padding=1)
    p_out.reverse()
    return p_out


@torch.no_grad()
def head_forward(hsd, feats):
    depth = len([k for k in hsd if k.startswith("conv.") and k.endswith(".weight")])
    logits, deltas = [], []
    for f in feats:
        t = f
        for i in range(depth):
            t = F.relu(F.conv3d(t, hsd[f"conv.{2 * i}.weight"], hsd[f"conv.{2 * i}.bias"], padding=1))
        logits.append(F.conv3d(t, hsd["cls_logits.weight"], hsd["cls_logits.bias"]))
        deltas.append(F.conv3d(t, hsd["bbox_pred.weight"], hsd["bbox_pred.bias"]))
    return logits, deltas


def flatten_predictions(logits, deltas, A, code, n=0):
    """(N,A,X,Y,Z) / (N,A*code,X,Y,Z) -> per level flat (V*A,), (V*A, code) in the reference's anchor order."""
    lg = [np.transpose(l[n].numpy(), (1, 2, 3, 0)).reshape(-1) for l in logits]
    dl = [np.transpose(d[n].numpy().reshape(A, code, *d.shape[2:]), (2, 3, 4, 0, 1)).reshape(-1, code) for d in deltas]
    return lg, dl


@torch.no_grad()
def full_forward(sd, hsd, x, cells, rotated, pre_nms_top_n=2500, post_nms_top_n=2500, nms_thresh=0.3, score_thresh=0.0):
    """One scene (1,4,W,L,H): features, and (boxes, scores, levels) as NeRFRegionProposalNetwork.forward returns them."""
    feats = resnet_fpn_forward(sd, x)
    logits, deltas = head_forward(hsd, feats)
    A = cells[0].shape[0]
    code = 8 if rotated else 6
    lg, dl = flatten_predictions(logits, deltas, A, code)
    mesh = tuple(x.shape[-3:])
    grids = [tuple(f.shape[-3:]) for f in feats]
    strides = [tuple(mesh[i] // g[i] for i in range(3)) for g in grids]
    b, s, lv = rp.rpn_proposals(lg, dl, grids, strides, cells, mesh, rotated, pre_nms_top_n, post_nms_top_n, nms_thresh, score_thresh)
    return feats, (b, s, lv)


@torch.no_grad()
def vgg_fpn_forward(sd, x):
    """VGG_FPN.forward (nerf_rpn/model/feature_extractor.py:362-377, make_layers :331-360) + FPN.forward (fpn.py:134-161),
    functional fp32, keyed by the reference's state_dict names (layers.{i}[.{j}].*, fpn_neck.{lateral,fpn}_convs.{i}.*)."""
    strided = "layers.3.0.weight" not in sd and any(k.startswith("layers.4.") for k in sd)       # stem has a max-pool at index 3
    h = F.conv3d(x, sd["layers.0.weight"], sd["layers.0.bias"], stride=2 if strided else 1, padding=3)
    h = F.relu(_bn(h, sd, "layers.1"))
    if strided:
        h = F.max_pool3d(h, kernel_size=3, stride=2, padding=1)
    first = 4 if strided else 3
    groups = sorted({int(k.split(".")[1]) for k in sd if k.startswith("layers.") and int(k.split(".")[1]) >= first})
    feats = []
    for gi in groups:
        idxs = sorted({int(k.split(".")[2]) for k in sd if k.startswith(f"layers.{gi}.")})
        convs = [j for j in idxs if f"layers.{gi}.{j}.weight" in sd and sd[f"layers.{gi}.{j}.weight"].dim() == 5]
        for j in convs:
            h = F.conv3d(h, sd[f"layers.{gi}.{j}.weight"], sd[f"layers.{gi}.{j}.bias"], padding=1)
            if f"layers.{gi}.{j + 1}.running_mean" in sd:
                h = _bn(h, sd, f"layers.{gi}.{j + 1}")
            h = F.relu(h)
        if gi != groups[0]:                          # every stage but the first ends with MaxPool3d(2, 2, ceil_mode=True) in the *F cfgs
            h = F.max_pool3d(h, kernel_size=2, stride=2, ceil_mode=True)
        feats.append(h)
    lat = [F.conv3d(f, sd[f"fpn_neck.lateral_convs.{i}.weight"], sd[f"fpn_neck.lateral_convs.{i}.bias"]) for i, f in enumerate(feats)]
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    return [F.conv3d(l, sd[f"fpn_neck.fpn_convs.{i}.weight"], sd[f"fpn_neck.fpn_convs.{i}.bias"], padding=1) for i, l in enumerate(lat)]
