"""Stand-alone forward() of the sub-modules the reference exposes: Bottleneck.forward (feature_extractor.py:48-68),
FPN.forward (fpn.py:134-161), RPNHead.forward (anchor.py:206-213), FCOSHead.forward (fcos/fcos.py:104-130).

Inside NeRFRegionProposalNetwork / FCOSOverNeRF these layers run as part of one captured launch sequence (engine.py).  A caller that
uses a sub-module on its own (the reference's detector.py does, a user script may) gets the same kernels, launched eagerly: NCDHW fp32
tensors in and out like the reference, channels-last 16-bit inside (the layout conversions are torch copies -- plumbing --; every
convolution / normalisation is a libnerf_rpn_b200 launch).  Eval-mode semantics (BatchNorm running statistics); in training mode
the whole-model training engine (nerf_rpn_b200/train.py) is the supported path and these functions raise.
"""
from typing import List, Sequence, Tuple

import torch

from .. import ops
from ..engine import _Conv
from ..precision import resolve as _resolve_precision

_CACHE = {}


def _version(module) -> tuple:
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def _packed(module, precision, build):
    """Pre-packed weights of `module`, rebuilt when any parameter / buffer changed (storage or version)."""
    key = (id(module), precision)
    ver = _version(module)
    hit = _CACHE.get(key)
    if hit is None or hit[0] != ver:
        old = _Conv.dtype
        _Conv.dtype = torch.bfloat16 if precision == "bf16" else torch.float16
        try:
            hit = (ver, build())
        finally:
            _Conv.dtype = old
        _CACHE[key] = hit
    return hit[1]


def _check(x: torch.Tensor, training: bool, what: str):
    if training:
        raise NotImplementedError(f"nerf_rpn_b200: {what}.forward in training mode is not a stand-alone path; train through "
                                  "NeRFRegionProposalNetwork (nerf_rpn_b200/train.py)")
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"nerf_rpn_b200: {what}.forward needs CUDA tensors (no CPU path)")


def _to_cl(x: torch.Tensor, dtype) -> torch.Tensor:
    return x.permute(0, 2, 3, 4, 1).to(dtype).contiguous()          # (N,C,X,Y,Z) fp32 -> (N,X,Y,Z,C) 16-bit


def _from_cl(y: torch.Tensor, c: int = None) -> torch.Tensor:
    y = y if c is None else y[..., :c]
    return y.permute(0, 4, 1, 2, 3).float()                           # NCDHW view semantics (channels_last_3d strides) like the engine


def _conv(layer: _Conv, xs: Sequence[torch.Tensor], out_dims, res=None, out_fp32=False):
    """One launch over the levels in xs (shared weights); returns the outputs (N, *out_dims[i], layer.cout)."""
    n = xs[0].shape[0]
    dt = torch.float32 if out_fp32 else xs[0].dtype
    ys = [torch.empty((n, *od, layer.cout), dtype=dt, device=xs[0].device) for od in out_dims]
    args = []
    for i, x in enumerate(xs):
        r = None if res is None else res[i]
        args.append(ops.ConvLevelArgs(x, ys[i], n, x.shape[1:4], out_dims[i], layer.cout, res=r, res_dims=None if r is None else r.shape[1:4],
                                      ldr=0 if r is None else r.shape[-1]))
    ops.conv3d_fprop(args, layer.w, layer.shift, layer.cin, layer.cout, layer.taps, stride=layer.stride, relu=layer.relu, out_fp32=out_fp32)
    return ys


def _down(d):
    return tuple((v - 1) // 2 + 1 for v in d)


def _dtype(precision):
    return torch.bfloat16 if precision == "bf16" else torch.float16


# ---------------------------------------------------------------------------------------------- Bottleneck
def bottleneck_forward(blk, x: torch.Tensor, precision=None) -> torch.Tensor:
    """feature_extractor.py:48-68: relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1(x)))))))) + [downsample](x)), BatchNorm folded."""
    _check(x, blk.training, "Bottleneck")
    precision = _resolve_precision(precision)
    ws = precision == "fp16_w2"

    def build():
        dev = x.device
        s = blk.stride
        L = dict(c1=_Conv(blk.conv1.weight, None, blk.bn1, stride=s, relu=True, device=dev, wsplit=ws),
                 c2=_Conv(blk.conv2.weight, None, blk.bn2, relu=True, device=dev, wsplit=ws and blk.conv2.weight.shape[1] != 64),
                 c3=_Conv(blk.conv3.weight, None, blk.bn3, relu=True, device=dev, wsplit=ws), ds=None)
        if blk.downsample is not None:
            L["ds"] = _Conv(blk.downsample[0].weight, None, blk.downsample[1], stride=s, relu=False, device=dev, wsplit=ws)
        return L
    L = _packed(blk, precision, build)
    xc = _to_cl(x, _dtype(precision))
    xd = tuple(xc.shape[1:4])
    od = _down(xd) if blk.stride == 2 else xd
    a = _conv(L["c1"], [xc], [od])[0]
    b = _conv(L["c2"], [a], [od])[0]
    r = xc if L["ds"] is None else _conv(L["ds"], [xc], [od])[0]
    o = _conv(L["c3"], [b], [od], res=[r])[0]
    return _from_cl(o)


# ---------------------------------------------------------------------------------------------- FPN
def fpn_forward(fpn, inputs: Sequence[torch.Tensor], precision=None) -> Tuple[torch.Tensor, ...]:
    """fpn.py:134-161: laterals, in-place top-down nearest-upsample accumulation, one 3^3 conv per level."""
    assert len(inputs) == len(fpn.in_channels)
    _check(inputs[0], fpn.training, "FPN")
    precision = _resolve_precision(precision)
    ws = precision == "fp16_w2"

    def build():
        dev = inputs[0].device
        return dict(lat=[_Conv(m.weight, m.bias, device=dev, wsplit=ws) for m in fpn.lateral_convs],
                    out=[_Conv(m.weight, m.bias, device=dev, wsplit=ws and i > 0) for i, m in enumerate(fpn.fpn_convs)])
    L = _packed(fpn, precision, build)
    xs = [_to_cl(t, _dtype(precision)) for t in inputs]
    n_l = len(xs)
    lat = [None] * n_l
    for i in range(n_l - 1, -1, -1):
        d = tuple(xs[i].shape[1:4])
        lat[i] = _conv(L["lat"][i], [xs[i]], [d], res=None if i == n_l - 1 else [lat[i + 1]])[0]
    return tuple(_from_cl(_conv(L["out"][i], [lat[i]], [tuple(lat[i].shape[1:4])])[0]) for i in range(n_l))


# ---------------------------------------------------------------------------------------------- RPNHead
def rpn_head_pred(head, feats: Sequence[torch.Tensor], precision=None):
    """The head's fused output: per level fp32 (N, w, l, h, 128) rows [A logits | A*code deltas | zero pad] -- the layout the
    post-processing kernel (nrpn_rpn_proposals) reads."""
    _check(feats[0], head.training, "RPNHead")
    precision = _resolve_precision(precision)
    if precision == "fp16_w2":
        precision_w = "fp16"                               # the head keeps single halves (precision.py)
    else:
        precision_w = precision

    def build():
        dev = feats[0].device
        convs = [m for m in head.conv if isinstance(m, torch.nn.Conv3d)]
        w = torch.cat([head.cls_logits.weight, head.bbox_pred.weight], 0)
        b = torch.cat([head.cls_logits.bias, head.bbox_pred.bias], 0)
        if w.shape[0] > 128:
            raise ValueError("fused predictor supports at most 128 output channels")
        return dict(conv=[_Conv(m.weight, m.bias, relu=True, device=dev) for m in convs], pred=_Conv(w, b, device=dev, cout_pad_to=128))
    L = _packed(head, precision_w, build)
    cur = [_to_cl(t, _dtype(precision_w)) for t in feats]
    dims = [tuple(t.shape[1:4]) for t in cur]
    for layer in L["conv"]:
        cur = _conv(layer, cur, dims)
    return _conv(L["pred"], cur, dims, out_fp32=True)


def rpn_head_forward(head, feats: Sequence[torch.Tensor], precision=None):
    """anchor.py:206-213: per level conv stack (shared weights, all levels per launch) -> (logits (N,A,w,l,h), bbox_reg (N,A*code,w,l,h))."""
    A = head.cls_logits.weight.shape[0]
    nb = head.bbox_pred.weight.shape[0]
    pred = rpn_head_pred(head, feats, precision)
    logits = [p[..., :A].permute(0, 4, 1, 2, 3).contiguous() for p in pred]
    bbox = [p[..., A:A + nb].permute(0, 4, 1, 2, 3).contiguous() for p in pred]
    return logits, bbox


# ---------------------------------------------------------------------------------------------- FCOSHead
def fcos_head_forward(head, feats: Sequence[torch.Tensor], precision=None):
    """fcos/fcos.py:104-130 (eval): towers of conv + GroupNorm(32) + ReLU shared over levels, 3^3 predictors, per-level Scale, ReLU on the
    first six regressors and x stride -> (logits, bbox_reg, centerness) lists of (N,C,w,l,h) fp32."""
    _check(feats[0], head.training, "FCOSHead")
    precision = _resolve_precision(precision)
    precision_w = "fp16" if precision == "fp16_w2" else precision
    code = head.bbox_pred.weight.shape[0]

    def build():
        dev = feats[0].device

        def tower(seq):
            mods, out = list(seq.children()), []
            for i in range(0, len(mods), 3):
                conv, gn = mods[i], mods[i + 1]
                out.append((_Conv(conv.weight, conv.bias, relu=False, device=dev), gn.weight.detach().float().to(dev).contiguous(),
                            gn.bias.detach().float().to(dev).contiguous(), float(gn.eps)))
            return out
        w = torch.cat([head.bbox_pred.weight, head.centerness.weight], 0)
        b = torch.cat([head.bbox_pred.bias, head.centerness.bias], 0)
        return dict(cls=tower(head.cls_tower), box=tower(head.bbox_tower), cls_pred=_Conv(head.cls_logits.weight, head.cls_logits.bias, device=dev),
                    reg_pred=_Conv(w, b, device=dev), scales=[float(s.scale.detach().item()) for s in head.scales])
    L = _packed(head, precision_w, build)
    x0 = [_to_cl(t, _dtype(precision_w)) for t in feats]
    dims = [tuple(t.shape[1:4]) for t in x0]
    ends = {}
    for name in ("cls", "box"):
        cur = x0
        for (layer, gamma, beta, eps) in L[name]:
            cur = _conv(layer, cur, dims)
            ops.groupnorm_relu_(cur, gamma, beta, eps, True, 32)
        ends[name] = cur
    cls = _conv(L["cls_pred"], ends["cls"], dims, out_fp32=True)
    reg = _conv(L["reg_pred"], ends["box"], dims, out_fp32=True)
    logits = [c[..., :1].permute(0, 4, 1, 2, 3).contiguous() for c in cls]
    bbox_reg, ctr = [], []
    for l, r in enumerate(reg):
        bp = r[..., :code].permute(0, 4, 1, 2, 3).contiguous() * L["scales"][l]
        bp[:, :6] = torch.relu(bp[:, :6]) * head.fpn_strides[l]
        bbox_reg.append(bp)
        ctr.append(r[..., code:code + 1].permute(0, 4, 1, 2, 3).contiguous())
    return logits, bbox_reg, ctr
