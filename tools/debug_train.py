#!/usr/bin/env python
"""GPU-box diagnostic: where does the training step's gradient error against the reference's fp32 autograd come from?
Compares, tensor by tensor, the forward activations (features, head layers, predictor) and the backward activations (d loss / d head layer,
d loss / d feature) of nerf_rpn_b200.train with the UNMODIFIED reference (oracle/_ref) run in fp32 AND under torch.autocast (the
mixed-precision baseline a PyTorch user would get), on the same scene, weights and sampled anchors."""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_gpu  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def cl(t):            # ours (n,X,Y,Z,C) -> reference layout (n,C,X,Y,Z)
    return t.permute(0, 4, 1, 2, 3).float()


def run_reference(rotated, grid, gt, autocast_dtype=None, seed_w=0):
    m = ref_gpu.build_reference_model(rotated=rotated, seed=seed_w, rpn_fg_iou_thresh=0.35, rpn_bg_iou_thresh=0.2).cuda().train()
    acts, grads = {}, {}
    convs = [c for c in m.rpn.head.conv if isinstance(c, torch.nn.ReLU)]

    def keep(name):
        def hook(mod, inp, out):
            out.retain_grad(); acts[name] = out
        return hook
    for k, c in enumerate(convs):
        c.register_forward_hook(keep(f"h{k}"))
    feats_holder = {}
    orig_head_fwd = m.rpn.head.forward

    def head_fwd(x):
        for i, f in enumerate(x):
            f.retain_grad(); feats_holder[i] = f
        return orig_head_fwd(x)
    m.rpn.head.forward = head_fwd
    rec = {}
    orig = m.rpn.fg_bg_sampler

    def sampler(labels):
        p, n = orig(labels)
        rec["pos"] = [torch.where(a)[0] for a in p]; rec["neg"] = [torch.where(a)[0] for a in n]
        return p, n
    m.rpn.fg_bg_sampler = sampler
    torch.manual_seed(123)
    ctx = torch.autocast("cuda", dtype=autocast_dtype) if autocast_dtype is not None else torch.autocast("cuda", enabled=False)
    with ctx:
        _, losses, _ = m([grid], [gt])
        loss = losses["loss_objectness"] + 5.0 * losses["loss_rpn_box_reg"]
    loss.backward()
    params = list(m.backbone.parameters()) + list(m.rpn.head.parameters())
    out = dict(losses=(losses["loss_objectness"].item(), losses["loss_rpn_box_reg"].item()), grads=[p.grad.detach().float().clone() for p in params],
               names=[n for n, _ in m.backbone.named_parameters()] + ["head." + n for n, _ in m.rpn.head.named_parameters()],
               feats=[feats_holder[i].detach().float() for i in range(4)], dfeats=[feats_holder[i].grad.detach().float() for i in range(4)],
               hs=[], dhs=[], rec=rec, sd=({k: v.detach().clone() for k, v in m.backbone.state_dict().items()}, {k: v.detach().clone() for k, v in m.rpn.head.state_dict().items()}))
    # head activations: the hooks fire once per level per ReLU: acts keeps the LAST level only; good enough for a relative error per layer
    for k in range(len(convs)):
        out["hs"].append(acts[f"h{k}"].detach().float()); out["dhs"].append(acts[f"h{k}"].grad.detach().float())
    del m
    torch.cuda.empty_cache()
    return out


def main():
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    from nerf_rpn_b200.model.anchor import AnchorGenerator3D, RPNHead
    from nerf_rpn_b200.model.feature_extractor import Bottleneck, ResNet_FPN_256
    from nerf_rpn_b200.model.nerf_rpn import NeRFRegionProposalNetwork
    from nerf_rpn_b200.train import RPNTrainEngine
    rotated = False
    dims = (64, 96, 80)
    g = torch.Generator().manual_seed(11)
    grid = torch.rand(*dims, 4, generator=g).permute(3, 0, 1, 2).contiguous().cuda()
    d = torch.tensor(dims, dtype=torch.float32)
    size = torch.rand(12, 3, generator=g) * 20.0 + 6.0
    ctr = torch.rand(12, 3, generator=g) * (d - 8.0) + 4.0
    gt = torch.cat([ctr - size / 2, ctr + size / 2], 1).cuda()
    r32 = run_reference(rotated, grid, gt, None)
    for tag, dt in (("autocast-bf16", torch.bfloat16), ("autocast-fp16", torch.float16)):
        ra = run_reference(rotated, grid, gt, dt)
        fg, fr = torch.cat([x.reshape(-1) for x in ra["grads"]]), torch.cat([x.reshape(-1) for x in r32["grads"]])
        print(f"[reference {tag} vs its fp32] losses {ra['losses']} vs {r32['losses']}; gradient cosine {F.cosine_similarity(fg, fr, dim=0).item():.4f} rel {rel(fg, fr):.3f}; "
              f"feats {[round(rel(a, b), 4) for a, b in zip(ra['feats'], r32['feats'])]} dfeats {[round(rel(a, b), 3) for a, b in zip(ra['dfeats'], r32['dfeats'])]} "
              f"head dh (last level) {[round(rel(a, b), 3) for a, b in zip(ra['dhs'], r32['dhs'])]}", flush=True)
        sel = [("head.cls_logits.weight", -4), ("head.conv.0.weight", -12)]
        print("   per-tensor:", {n: round(rel(a, b), 3) for n, a, b in zip(ra["names"], ra["grads"], r32["grads"]) if n in ("head.cls_logits.weight", "head.conv.0.weight", "latlayers.0.weight", "layers.3.2.conv3.weight", "layers.0.0.conv1.weight", "conv1.weight")})
    for precision in ("fp16", "bf16"):
        backbone = ResNet_FPN_256(Bottleneck, [3, 4, 6, 3], input_dim=4, is_max_pool=True)
        head = RPNHead(256, 13, 4, rotate=rotated)
        backbone.load_state_dict(r32["sd"][0]); head.load_state_dict(r32["sd"][1])
        model = NeRFRegionProposalNetwork(backbone, AnchorGenerator3D(ref_gpu.ANCHOR_SIZES, ref_gpu.ASPECT), head, rpn_fg_iou_thresh=0.35, rpn_bg_iou_thresh=0.2,
                                          rotated_bbox=rotated).cuda().train()
        eng = RPNTrainEngine(model, precision=precision)
        plan = eng.plan(1, dims)
        plan.forced_samples = [(r32["rec"]["pos"][0], r32["rec"]["neg"][0])]
        eng.forward_backward(grid[None], [gt])
        torch.cuda.synchronize()
        inv = 1.0 / eng.loss_scale
        D = plan.dbg
        feats = [cl(f) for f in plan.features]
        print(f"[ours {precision} vs reference fp32] losses {eng.losses.tolist()} vs {r32['losses']}")
        print("   forward : feats", [round(rel(a, b), 4) for a, b in zip(feats, r32["feats"])],
              "head h_k (last level)", [round(rel(cl(D['hs'][k][3]), r32['hs'][k]), 4) for k in range(4)])
        print("   backward: dfeats", [round(rel(cl(a) * inv, b), 3) for a, b in zip(D["dfeats"], r32["dfeats"])],
              "head dh_k (last level, masked by ReLU in ours)", [round(rel(cl(D['dhs'][k][3]) * inv, r32['dhs'][k] * (r32['hs'][k] > 0)), 3) for k in range(4)])
        params = list(backbone.parameters()) + list(head.parameters())
        gg = [eng.grad_of(p).view(p.shape).clone() * inv for p in params]
        fg, fr = torch.cat([x.reshape(-1) for x in gg]), torch.cat([x.reshape(-1) for x in r32["grads"]])
        print(f"   gradient cosine {F.cosine_similarity(fg, fr, dim=0).item():.4f} rel {rel(fg, fr):.3f}")
        print("   per-tensor:", {n: round(rel(a, b), 3) for n, a, b in zip(r32["names"], gg, r32["grads"]) if n in ("head.cls_logits.weight", "head.conv.0.weight", "latlayers.0.weight", "layers.3.2.conv3.weight", "layers.0.0.conv1.weight", "conv1.weight")})
        del eng, model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
