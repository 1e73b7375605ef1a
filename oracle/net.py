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


# ------------------------------------------------------------------------------------------------ Swin
def _window_attention(x, sd, p, heads, shift, win=4):
    """shifted_window_attention (nerf_rpn/model/feature_extractor.py:382-497) for window = 4^3, restated:
    zero-pad to multiples of the window AFTER norm1 (padded tokens take part), cyclic shift by -shift when shifted,
    per-window multi-head attention with q scaled by head_dim^-0.5, relative-position bias, -100 mask between the 27
    shift regions, softmax, projection, un-shift, crop."""
    B, H, W, D, C = x.shape
    ph, pw, pd = (-H) % win, (-W) % win, (-D) % win
    xp = F.pad(x, (0, 0, 0, pd, 0, pw, 0, ph))
    PH, PW, PD = H + ph, W + pw, D + pd
    sh = [0 if win >= e else shift for e in (PH, PW, PD)]
    if sum(sh) > 0:
        xp = torch.roll(xp, shifts=(-sh[0], -sh[1], -sh[2]), dims=(1, 2, 3))
    nh, nw, nd = PH // win, PW // win, PD // win
    t = xp.view(B, nh, win, nw, win, nd, win, C).permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(B * nh * nw * nd, win ** 3, C)
    qkv = F.linear(t, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"]).reshape(t.shape[0], win ** 3, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + ".relative_position_bias_table"][sd[p + ".relative_position_index"]].view(win ** 3, win ** 3, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if sum(sh) > 0:
        region = x.new_zeros((PH, PW, PD))
        cnt = 0
        for hs in ((0, -win), (-win, -sh[0]), (-sh[0], None)):
            for ws in ((0, -win), (-win, -sh[1]), (-sh[1], None)):
                for ds in ((0, -win), (-win, -sh[2]), (-sh[2], None)):
                    region[hs[0]:hs[1], ws[0]:ws[1], ds[0]:ds[1]] = cnt
                    cnt += 1
        region = region.view(nh, win, nw, win, nd, win).permute(0, 2, 4, 1, 3, 5).reshape(nh * nw * nd, win ** 3)
        diff = region.unsqueeze(1) - region.unsqueeze(2)
        mask = torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))
        attn = (attn.view(B, nh * nw * nd, heads, win ** 3, win ** 3) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, win ** 3, win ** 3)
    attn = F.softmax(attn, dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(t.shape[0], win ** 3, C)
    o = F.linear(o, sd[p + ".proj.weight"], sd[p + ".proj.bias"])
    o = o.view(B, nh, nw, nd, win, win, win, C).permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(B, PH, PW, PD, C)
    if sum(sh) > 0:
        o = torch.roll(o, shifts=(sh[0], sh[1], sh[2]), dims=(1, 2, 3))
    return o[:, :H, :W, :D, :].contiguous()


def _ln(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


@torch.no_grad()
def swin_fpn_forward(sd, x, depths, num_heads):
    """SwinTransformer_FPN.forward (feature_extractor.py:781-789): patch embed (Conv3d k4 s4 + LN), stages of
    [PatchMerging] + SwinTransformerBlocks (:593-686), FPN neck (fpn.py:134-161). Channels-last (B,H,W,D,C) inside."""
    h = F.conv3d(x, sd["patch_partition.0.weight"], sd["patch_partition.0.bias"], stride=4).permute(0, 2, 3, 4, 1)
    h = _ln(h, sd, "patch_partition.2")
    feats = []
    for s, depth in enumerate(depths):
        idx = 0
        if s > 0:
            p = f"stages.{s}.0"
            H, W, D = h.shape[1:4]
            hp = F.pad(h, (0, 0, 0, D % 2, 0, W % 2, 0, H % 2))
            parts = [hp[:, i::2, j::2, k::2, :] for (i, j, k) in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1))]
            h = F.linear(_ln(torch.cat(parts, -1), sd, p + ".norm"), sd[p + ".reduction.weight"])
            idx = 1
        for b in range(depth):
            p = f"stages.{s}.{idx + b}"
            h = h + _window_attention(_ln(h, sd, p + ".norm1"), sd, p + ".attn", num_heads[s], 0 if b % 2 == 0 else 2)
            m = F.linear(F.gelu(F.linear(_ln(h, sd, p + ".norm2"), sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"])),
                         sd[p + ".mlp.3.weight"], sd[p + ".mlp.3.bias"])
            h = h + m
        feats.append(h.permute(0, 4, 1, 2, 3).contiguous())
    lat = [F.conv3d(f, sd[f"fpn_neck.lateral_convs.{i}.weight"], sd[f"fpn_neck.lateral_convs.{i}.bias"]) for i, f in enumerate(feats)]
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    return [F.conv3d(l, sd[f"fpn_neck.fpn_convs.{i}.weight"], sd[f"fpn_neck.fpn_convs.{i}.bias"], padding=1) for i, l in enumerate(lat)]
