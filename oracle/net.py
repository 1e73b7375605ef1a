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
