"""CPU: the FCOS training loss (fcos/loss.py of the reference).
  * oracle/fcos_loss_oracle.py against tests/golden/fcos_loss.npz = outputs of the UNMODIFIED reference (tools/make_golden.py gen_fcos_loss):
    targets bit for bit (AABB) / to 1e-5 (OBB corners), losses and autograd gradients to 1e-5;
  * the device functions of csrc/fcos_loss.cuh compiled for the host (tests/host_shim/fcos_loss_host.cpp) against the same vectors:
    what the CUDA kernels execute per element, checked without a GPU.
The rotated-IoU term of the OBB head needs the reference's CUDA vertex sort: tests/test_gpu_fcos_loss.py checks it on the GPU box."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import fcos_loss_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {  # name: (rotated, iou_loss_type, center_sampling_radius, additional_l1, batch, proj2d_loss_weight)   == tools/make_golden.py FCOS_LOSS_CASES
    "aabb_iou": (False, "iou", 1.5, False, 2, 0.0), "aabb_giou": (False, "giou", 1.5, False, 2, 0.0), "aabb_linear": (False, "linear_iou", 0.0, False, 1, 0.0),
    "aabb_sl1": (False, "smooth_l1", 1.5, False, 2, 0.0), "obb_sl1": (True, "smooth_l1", 1.5, False, 2, 0.0), "obb_iou_l1": (True, "iou", 1.5, True, 2, 0.0),
    "obb_nocs": (True, "smooth_l1", 0.0, False, 1, 0.0), "aabb_empty": (False, "iou", 1.5, False, 2, 0.0),
    "obb_sl1_p2d": (True, "smooth_l1", 1.5, False, 2, 0.5), "obb_iou_p2d": (True, "iou", 1.5, True, 1, 0.25)}
STRIDES = [4, 8, 16, 32]
LOSS_TYPE = {"smooth_l1": 0, "iou": 1, "linear_iou": 2, "giou": 3}
WEIGHTS = (1.0, 2.0, 3.0)          # the golden gradients are those of loss_cls + 2 loss_reg + 3 loss_centerness


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "fcos_loss.npz"))


def load_case(g, name):
    rotated, loss_type, radius, add_l1, batch, proj2d = CASES[name]
    c = dict(rotated=rotated, loss_type=loss_type, radius=radius, add_l1=add_l1, batch=batch, proj2d=proj2d)
    for k in ("cls", "reg", "ctr", "dcls", "dreg", "dctr", "labels", "reg_targets"):
        c[k] = [g[f"{name}/{k}{l}"] for l in range(4)]
    c["mask"] = [g[f"{name}/mask{l}"] for l in range(4)] if batch > 1 else None
    c["gt"] = [g[f"{name}/gt{b}"] for b in range(batch)]
    c["losses"] = g[f"{name}/losses"]
    c["grids"] = [t.shape[2:] for t in c["cls"]]
    c["n_per"] = [int(np.prod(s)) for s in c["grids"]]
    return c


def per_scene(level_first, n_per, batch):
    """golden level-first lists [(N * P_l, ...)] -> per scene [(P, ...)] (levels concatenated)."""
    return [np.concatenate([level_first[l][n * pl:(n + 1) * pl] for l, pl in enumerate(n_per)]) for n in range(batch)]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_targets_match_reference(golden, name):
    c = load_case(golden, name)
    locs = O.compute_locations(c["grids"], STRIDES)
    want_l, want_r = per_scene(c["labels"], c["n_per"], c["batch"]), per_scene(c["reg_targets"], c["n_per"], c["batch"])
    for n in range(c["batch"]):
        lab, rt = O.targets(locs, STRIDES, c["gt"][n], c["radius"], True)
        np.testing.assert_array_equal(lab, want_l[n])
        if c["rotated"]:
            np.testing.assert_allclose(rt, want_r[n], rtol=1e-5, atol=1e-5)
        else:
            np.testing.assert_array_equal(rt, want_r[n])
    if name != "aabb_empty":
        assert sum(int((l > 0).sum()) for l in want_l) > 50


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_loss_and_gradients_match_reference(golden, name):
    c = load_case(golden, name)
    cls, reg, ctr = ([torch.tensor(a, requires_grad=True) for a in c[k]] for k in ("cls", "reg", "ctr"))
    lab, rt = per_scene(c["labels"], c["n_per"], c["batch"]), per_scene(c["reg_targets"], c["n_per"], c["batch"])
    l_cls, l_reg, l_ctr, _ = O.loss(cls, reg, ctr, lab, rt, c["mask"], c["loss_type"], c["rotated"], c["add_l1"], c["proj2d"])
    rotated_iou = c["rotated"] and c["loss_type"] != "smooth_l1"
    np.testing.assert_allclose(l_cls.item(), c["losses"][0], rtol=1e-5)
    np.testing.assert_allclose(l_ctr.item(), c["losses"][2], rtol=1e-5)
    if not rotated_iou:
        np.testing.assert_allclose(l_reg.item(), c["losses"][1], rtol=1e-5)
    (WEIGHTS[0] * l_cls + WEIGHTS[1] * l_reg + WEIGHTS[2] * l_ctr).backward()
    for l in range(4):
        np.testing.assert_allclose(cls[l].grad.numpy(), c["dcls"][l], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(ctr[l].grad.numpy(), c["dctr"][l], rtol=1e-4, atol=1e-7)
        if not rotated_iou:
            np.testing.assert_allclose(reg[l].grad.numpy(), c["dreg"][l], rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------------------------------------ the device functions, host build
@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libfcos_loss_shim.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host_shim", "fcos_loss_host.cpp")])
    return ctypes.CDLL(out)


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def shim_targets(shim, c, n):
    locs = np.ascontiguousarray(np.concatenate(O.compute_locations(c["grids"], STRIDES)))
    begin = np.concatenate([[0], np.cumsum(c["n_per"])]).astype(np.int32)
    rs = np.array([np.float32(s * c["radius"]) if c["radius"] > 0 else 0.0 for s in STRIDES], np.float32)
    soi = np.array(O.SIZES_OF_INTEREST, np.float32)
    lo, hi = np.ascontiguousarray(soi[:, 0]), np.ascontiguousarray(soi[:, 1])
    nd = np.array(STRIDES, np.float32)
    gt = np.ascontiguousarray(c["gt"][n], np.float32)
    D = 8 if c["rotated"] else 6
    labels = np.empty(begin[-1], np.float32); rt = np.empty((begin[-1], D), np.float32)
    shim.shim_fcos_targets(_fp(locs), _fp(begin), 4, _fp(rs), _fp(lo), _fp(hi), 1, _fp(nd), _fp(gt), gt.shape[0], 7 if c["rotated"] else 6,
                           _fp(labels), _fp(rt))
    return labels, rt


@pytest.mark.parametrize("name", list(CASES))
def test_device_functions_targets(shim, golden, name):
    c = load_case(golden, name)
    want_l, want_r = per_scene(c["labels"], c["n_per"], c["batch"]), per_scene(c["reg_targets"], c["n_per"], c["batch"])
    for n in range(c["batch"]):
        lab, rt = shim_targets(shim, c, n)
        np.testing.assert_array_equal(lab, want_l[n])
        if c["rotated"]:
            np.testing.assert_allclose(rt, want_r[n], rtol=1e-5, atol=1e-5)
        else:
            np.testing.assert_array_equal(rt, want_r[n])


@pytest.mark.parametrize("name", list(CASES))
def test_device_functions_loss_and_gradients(shim, golden, name):
    c = load_case(golden, name)
    N, D = c["batch"], 8 if c["rotated"] else 6
    lab = np.ascontiguousarray(np.stack(per_scene(c["labels"], c["n_per"], N)))
    rt = np.ascontiguousarray(np.stack(per_scene(c["reg_targets"], c["n_per"], N)))
    mask = None if c["mask"] is None else np.ascontiguousarray(np.concatenate([m.reshape(N, -1) for m in c["mask"]], 1).astype(np.uint8))
    arrs = {k: [np.ascontiguousarray(a) for a in c[k]] for k in ("cls", "reg", "ctr")}
    grads = {k: [np.full_like(a, np.nan) for a in arrs[k]] for k in ("cls", "reg", "ctr")}
    PP = ctypes.c_void_p * 4

    def pp(lst):
        return PP(*[a.ctypes.data for a in lst])
    n_per = np.array(c["n_per"], np.int32)
    ct = np.empty_like(lab); sums = np.zeros(6, np.float64)
    shim.shim_fcos_loss(4, _fp(n_per), N, int(c["rotated"]), LOSS_TYPE[c["loss_type"]], int(c["add_l1"]), pp(arrs["cls"]), pp(arrs["reg"]),
                        pp(arrs["ctr"]), pp(grads["cls"]), pp(grads["reg"]), pp(grads["ctr"]), _fp(lab), _fp(rt),
                        ctypes.c_void_p(0) if mask is None else _fp(mask), _fp(ct), _fp(sums))
    focal, n_pos, sum_ct, reg_raw, bce, add = sums
    kept = lab > 0 if mask is None else (lab > 0) & (mask != 0)
    assert n_pos == kept.sum() and n_pos > 0
    np.testing.assert_allclose(ct[kept].sum(dtype=np.float64), sum_ct, rtol=1e-6)
    assert (ct[~kept] == 0).all()
    rotated_iou = (c["rotated"] and c["loss_type"] != "smooth_l1") or c["proj2d"] > 0     # terms of the gathered positives: not the kernel's
    np.testing.assert_allclose(focal / max(n_pos, 1.0), c["losses"][0], rtol=1e-5)
    np.testing.assert_allclose(bce / max(n_pos, 1.0), c["losses"][2], rtol=1e-5)
    if not rotated_iou:
        assert add == 0
        np.testing.assert_allclose(reg_raw / sum_ct, c["losses"][1], rtol=1e-5)
    for l in range(4):
        np.testing.assert_allclose(WEIGHTS[0] * grads["cls"][l] / max(n_pos, 1.0), c["dcls"][l], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(WEIGHTS[2] * grads["ctr"][l] / max(n_pos, 1.0), c["dctr"][l], rtol=2e-4, atol=1e-7)
        if not rotated_iou:
            np.testing.assert_allclose(WEIGHTS[1] * grads["reg"][l] / sum_ct, c["dreg"][l], rtol=2e-4, atol=1e-7)
        elif c["loss_type"] != "smooth_l1":         # only the alpha / beta smooth-L1 is the kernel's: channels 0..5 wait for the rotated-IoU term
            assert (grads["reg"][l][:, :6] == 0).all() and np.abs(grads["reg"][l][:, 6:]).sum() > 0 or l == 3


def test_torch_fcos_decode_matches_reference(golden):
    """nerf_rpn_b200/model/coder_torch.py decode_fcos_obb (the differentiable decode under the rotated-IoU loss) == fcos/utils.py:12-61."""
    from nerf_rpn_b200.model.coder_torch import decode_fcos_obb
    reg = torch.tensor(golden["decode/reg"], requires_grad=True)
    got = decode_fcos_obb(torch.tensor(golden["decode/loc"]), reg)
    np.testing.assert_allclose(got.detach().numpy(), golden["decode/boxes"], rtol=1e-5, atol=2e-5)
    got.sum().backward()
    assert torch.isfinite(reg.grad).all() and reg.grad.abs().sum() > 0


# ------------------------------------------------------------------------------------------------ host logic of nerf_rpn_b200/model/fcos/loss.py
# The module's kernels cannot run here; its HOST side (target layout, mask concatenation, the autograd node, the gather / scatter of the
# rotated-IoU term, the normalisers) can: the two ops are replaced by the host build of the same device functions (tests only).
def make_host_ops(shim):
    """(fcos_targets, fcos_loss_sums) with the signatures of nerf_rpn_b200.ops, running the host build of the device functions on CPU tensors."""
    def fcos_targets(locations, n_points, strides, gt, radius, norm=True):
        loc = np.ascontiguousarray(locations.numpy())
        begin = np.concatenate([[0], np.cumsum(n_points)]).astype(np.int32)
        rs = np.array([np.float32(s * radius) if radius > 0 else 0.0 for s in strides], np.float32)
        soi = np.array(O.SIZES_OF_INTEREST, np.float32)
        lo, hi = np.ascontiguousarray(soi[:, 0]), np.ascontiguousarray(soi[:, 1])
        nd = np.array(strides, np.float32)
        g = np.ascontiguousarray(gt.numpy(), np.float32)
        D = 8 if gt.shape[1] == 7 else 6
        labels = np.empty(begin[-1], np.float32); rt = np.empty((begin[-1], D), np.float32)
        shim.shim_fcos_targets(_fp(loc), _fp(begin), len(n_points), _fp(rs), _fp(lo), _fp(hi), int(norm), _fp(nd), _fp(g), g.shape[0], gt.shape[1],
                               _fp(labels), _fp(rt))
        return torch.from_numpy(labels), torch.from_numpy(rt)

    def fcos_loss_sums(box_cls, box_reg, ctr, labels, reg_targets, mask, loss_type, use_obb, add_l1, want_grad=True):
        arrs = [[np.ascontiguousarray(t.numpy()) for t in lst] for lst in (box_cls, box_reg, ctr)]
        grads = [[np.zeros_like(a) for a in lst] for lst in arrs]
        PP = ctypes.c_void_p * len(box_cls)
        pp = lambda lst: PP(*[a.ctypes.data for a in lst])
        n_per = np.array([a[0, 0].size for a in arrs[0]], np.int32)
        lab, rt = np.ascontiguousarray(labels.numpy()), np.ascontiguousarray(reg_targets.numpy())
        m = None if mask is None else np.ascontiguousarray(mask.numpy())
        ct = np.empty_like(lab); sums = np.zeros(8, np.float64)
        null = ctypes.c_void_p(0)
        shim.shim_fcos_loss(len(box_cls), _fp(n_per), lab.shape[0], int(use_obb), LOSS_TYPE[loss_type], int(add_l1), pp(arrs[0]), pp(arrs[1]), pp(arrs[2]),
                            pp(grads[0]) if want_grad else null, pp(grads[1]) if want_grad else null, pp(grads[2]) if want_grad else null,
                            _fp(lab), _fp(rt), null if m is None else _fp(m), _fp(ct), _fp(sums))
        g = tuple([torch.from_numpy(a) for a in lst] for lst in grads) if want_grad else None
        return torch.from_numpy(sums), torch.from_numpy(ct), g
    return fcos_targets, fcos_loss_sums


def patch_host_ops(shim, setattr_):
    """Route the module's two ops to the host build and let CPU tensors pass its CUDA checks (tests only)."""
    from nerf_rpn_b200 import ops
    fcos_targets, fcos_loss_sums = make_host_ops(shim)
    setattr_(ops, "fcos_targets", fcos_targets)
    setattr_(ops, "fcos_loss_sums", fcos_loss_sums)
    setattr_(torch.Tensor, "is_cuda", property(lambda self: True))
    _to = torch.Tensor.to
    setattr_(torch.Tensor, "to", lambda self, *a, **k: _to(self, *a, **{kk: ("cpu" if kk == "device" else v) for kk, v in k.items()}))
    return ops


@pytest.fixture(scope="module")
def shim_path(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim_mp") / "libfcos_loss_shim.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host_shim", "fcos_loss_host.cpp")])
    return out


@pytest.fixture()
def host_ops(shim, monkeypatch):
    return patch_host_ops(shim, monkeypatch.setattr)


def _module(c, world_size=1):
    import argparse
    from nerf_rpn_b200.model.fcos.fcos import FCOSModule
    args = argparse.Namespace(num_convs=1, norm_reg_targets=True, centerness_on_reg=True, rotated_bbox=c["rotated"], pre_nms_thresh=0.0, pre_nms_top_n=100,
                              nms_thresh=0.3, fpn_post_nms_top_n=100, min_size=0.0, center_sampling_radius=c["radius"], iou_loss_type=c["loss_type"],
                              use_additional_l1_loss=c["add_l1"], proj2d_loss_weight=c["proj2d"])
    return FCOSModule(args, 256, STRIDES, world_size=world_size)


@pytest.mark.parametrize("name", list(CASES))
def test_module_host_logic_against_reference(host_ops, golden, name):
    """FCOSModule.compute_locations / compute_padding_masks / loss_evaluator.prepare_targets / _forward_train of the mirror == the reference's
    outputs (level-first targets, masks, the three losses, the gradients of the head outputs); kernels replaced by their host build."""
    c = load_case(golden, name)
    mod = _module(c)
    cls, reg, ctr = ([torch.tensor(a, requires_grad=True) for a in c[k]] for k in ("cls", "reg", "ctr"))
    locs = mod.compute_locations(cls)
    for a, b in zip(locs, O.compute_locations(c["grids"], STRIDES)):
        np.testing.assert_array_equal(a.numpy(), b)
    sizes = golden[f"{name}/sizes"]
    masks = mod.compute_padding_masks(locs, [tuple(int(v) for v in s) for s in sizes]) if c["batch"] > 1 else None
    if masks is not None:
        for a, b in zip(masks, c["mask"]):
            np.testing.assert_array_equal(a.numpy(), b)
    gts = [torch.tensor(g) for g in c["gt"]]
    lab, rt = mod.loss_evaluator.prepare_targets(locs, gts)
    for l in range(4):
        np.testing.assert_array_equal(lab[l].numpy(), c["labels"][l])
        np.testing.assert_allclose(rt[l].numpy(), c["reg_targets"][l], rtol=1e-5, atol=1e-5)
    rotated_iou = c["rotated"] and c["loss_type"] != "smooth_l1"
    if rotated_iou:                      # the rotated IoU itself needs the GPU: a differentiable stand-in checks the gather / weight / scatter plumbing
        import nerf_rpn_b200.model.fcos.loss as L
        seen = {}

        def fake(pred, tgt, loss_type):
            seen["pred"], seen["tgt"] = pred.detach().clone(), tgt.clone()
            return ((pred - tgt) ** 2).sum(1)
        L_rot, L.rotated_iou_losses = L.rotated_iou_losses, fake
    try:
        _, _, losses = mod._forward_train(locs, cls, reg, ctr, gts, masks)
    finally:
        if rotated_iou:
            L.rotated_iou_losses = L_rot
    np.testing.assert_allclose(losses["loss_cls"].item(), c["losses"][0], rtol=1e-5)
    np.testing.assert_allclose(losses["loss_centerness"].item(), c["losses"][2], rtol=1e-5)
    (WEIGHTS[0] * losses["loss_cls"] + WEIGHTS[1] * losses["loss_reg"] + WEIGHTS[2] * losses["loss_centerness"]).backward()
    for l in range(4):
        np.testing.assert_allclose(cls[l].grad.numpy(), c["dcls"][l], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(ctr[l].grad.numpy(), c["dctr"][l], rtol=2e-4, atol=1e-7)
    if not rotated_iou:                  # incl. obb_sl1_p2d: decode + projection are torch ops and run here as they do on the GPU
        np.testing.assert_allclose(losses["loss_reg"].item(), c["losses"][1], rtol=1e-5)
        for l in range(4):
            np.testing.assert_allclose(reg[l].grad.numpy(), c["dreg"][l], rtol=2e-4, atol=2e-7)
    elif c["proj2d"] == 0:                                # stand-in: sum_i ct_i |pred_i - tgt_i|^2 / sum ct (+ the alpha / beta smooth-L1), gradient 2 ct (pred - tgt) / sum ct
        N = c["batch"]
        labs = np.stack(per_scene(c["labels"], c["n_per"], N)); rts = np.stack(per_scene(c["reg_targets"], c["n_per"], N))
        m = np.ones_like(labs, bool) if c["mask"] is None else np.concatenate([mm.reshape(N, -1) for mm in c["mask"]], 1)
        ct = O.centerness_targets(torch.tensor(rts.reshape(-1, 8))).numpy().reshape(N, -1)
        off, want, sum_ct = 0, 0.0, ct[(labs > 0) & m].sum(dtype=np.float64)
        assert seen["pred"].shape[0] == ((labs > 0) & m).sum()
        for l, pl in enumerate(c["n_per"]):
            p = c["reg"][l].reshape(N, 8, pl).transpose(0, 2, 1)
            sel = (labs[:, off:off + pl] > 0) & m[:, off:off + pl]
            d = p - rts[:, off:off + pl]
            w = np.where(sel, ct[:, off:off + pl], 0.0)[..., None]
            sl1 = np.where(np.abs(d) < 1, 0.5 * d * d, np.abs(d) - 0.5); dsl1 = np.where(np.abs(d) < 1, d, np.sign(d))
            want += (w * d * d).sum(dtype=np.float64) + (w * sl1)[..., 6:].sum(dtype=np.float64)
            g = 2 * w * d
            g[..., 6:] += (w * dsl1)[..., 6:]
            np.testing.assert_allclose(reg[l].grad.numpy().reshape(N, 8, pl).transpose(0, 2, 1), WEIGHTS[1] * g / sum_ct, rtol=2e-4, atol=1e-6)
            off += pl
        np.testing.assert_allclose(losses["loss_reg"].item(), want / sum_ct, rtol=1e-5)


def _normaliser_worker(rank, world, port, q):
    import torch.distributed as dist
    from nerf_rpn_b200.model.fcos.loss import normalisers
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    local = torch.tensor([[3.0, 1.25], [0.0, 0.0]][rank], dtype=torch.float64)          # rank 1 has no positives (loss.py:583-586 still all-reduces)
    n_pos, s_ct = normalisers(local, world)
    none_pos, none_ct = normalisers(torch.zeros(2, dtype=torch.float64), world)
    q.put((rank, n_pos.item(), s_ct.item(), none_pos.item(), none_ct.item()))
    dist.destroy_process_group()


def test_normalisers_all_reduce_two_ranks_gloo():
    """8(e) C3: the FCOS loss' two scalar all-reduces (num_pos, sum of centerness targets: loss.py:541-556), done as one two-element all-reduce."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_normaliser_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for rank, n_pos, s_ct, none_pos, none_ct in out:
        assert n_pos == 1.5 and s_ct == 0.625            # totals (3, 1.25) / 2 ranks on BOTH ranks
        assert none_pos == 1.0 and none_ct == 0.0        # max(0 / 2, 1) = 1


# ------------------------------------------------------------------------------------------------ randomised: device functions vs the oracle
FUZZ = [(False, 40, 1.5), (False, 40, 0.0), (False, 7, 2.5), (False, 1, 1.0), (False, 0, 1.5), (True, 40, 1.5), (True, 7, 0.0), (True, 0, 1.0),
        (False, 300, 1.5), (True, 300, 0.0)]


@pytest.mark.parametrize("seed", range(len(FUZZ)))
def test_device_functions_fuzz_against_oracle(shim, seed):
    """Random level layouts, strides, centre-sampling radii, normalisation on / off, integer-aligned boxes (exact ties of volume and of the > 0 tests),
    boxes outside the grid, G in {0, 1, many}: targets of the host-built device functions == the oracle; AABB losses and gradients == autograd."""
    rng = np.random.default_rng(100 + seed)
    n_levels = int(rng.integers(1, 5))
    strides = [4, 8, 16, 32][:n_levels]
    mesh = rng.integers(20, 70, 3)
    grids = [tuple(int(np.ceil(m / s)) for m in mesh) for s in strides]
    locs = O.compute_locations(grids, strides)
    n_per = [len(p) for p in locs]
    rotated, G, radius = FUZZ[seed]
    norm = bool(rng.integers(0, 2))
    ext = rng.random((G, 3)) * 50 + 3
    ctrs = rng.random((G, 3)) * (mesh + 20) - 10
    if G > 4:                                                  # integer boxes: locations exactly on faces (reg == 0) and equal volumes
        ext[: G // 2] = rng.integers(2, 12, (G // 2, 3)) * 4.0
        ctrs[: G // 2] = rng.integers(0, 16, (G // 2, 3)) * 4.0 + 2.0
        ext[1] = ext[0]
    gt = (np.concatenate([ctrs, ext, (rng.random((G, 1)) - 0.5) * np.pi], 1) if rotated else np.concatenate([ctrs - ext / 2, ctrs + ext / 2], 1)).astype(np.float32)
    if rotated and G > 4:
        gt[: G // 2, 6] = 0.0
    want_l, want_r = O.targets(locs, strides, gt.reshape(G, 7 if rotated else 6), radius, norm)
    c = dict(grids=grids, n_per=n_per, radius=radius, rotated=rotated, gt=[gt.reshape(G, 7 if rotated else 6)])
    loc_all = np.ascontiguousarray(np.concatenate(locs))
    begin = np.concatenate([[0], np.cumsum(n_per)]).astype(np.int32)
    rs = np.array([np.float32(s * radius) if radius > 0 else 0.0 for s in strides], np.float32)
    soi = np.array(O.SIZES_OF_INTEREST, np.float32)
    lo, hi = np.ascontiguousarray(soi[:, 0]), np.ascontiguousarray(soi[:, 1])
    nd = np.array(strides, np.float32)
    D = 8 if rotated else 6
    labels = np.empty(begin[-1], np.float32); rt = np.empty((begin[-1], D), np.float32)
    g2 = np.ascontiguousarray(c["gt"][0])
    shim.shim_fcos_targets(_fp(loc_all), _fp(begin), n_levels, _fp(rs), _fp(lo), _fp(hi), int(norm), _fp(nd), _fp(g2), G, 7 if rotated else 6, _fp(labels), _fp(rt))
    np.testing.assert_array_equal(labels, want_l)
    if rotated:
        np.testing.assert_allclose(rt, want_r, rtol=1e-5, atol=1e-5)
    else:
        np.testing.assert_array_equal(rt, want_r)
    if rotated or G == 0 or want_l.sum() == 0:
        return
    # losses of the axis-aligned head on these targets
    loss_type = ["iou", "linear_iou", "giou", "smooth_l1"][seed % 4]
    N = 1
    tg = torch.Generator().manual_seed(seed)
    cls = [torch.randn(N, 1, *gr, generator=tg) for gr in grids]
    reg = [torch.rand(N, 6, *gr, generator=tg) * 3 + 0.05 for gr in grids]
    ctr = [torch.randn(N, 1, *gr, generator=tg) for gr in grids]
    for lst in (cls, reg, ctr):
        for t in lst:
            t.requires_grad_(True)
    l_cls, l_reg, l_ctr, _ = O.loss(cls, reg, ctr, [want_l], [want_r], None, loss_type, False, False)
    (l_cls + 2.0 * l_reg + 3.0 * l_ctr).backward()
    arrs = [[np.ascontiguousarray(t.detach().numpy()) for t in lst] for lst in (cls, reg, ctr)]
    grads = [[np.full_like(a, np.nan) for a in lst] for lst in arrs]
    PP = ctypes.c_void_p * n_levels
    pp = lambda lst: PP(*[a.ctypes.data for a in lst])
    ct = np.empty((1, begin[-1]), np.float32); sums = np.zeros(6, np.float64)
    lab2, rt2 = np.ascontiguousarray(want_l[None]), np.ascontiguousarray(want_r[None])
    shim.shim_fcos_loss(n_levels, _fp(np.array(n_per, np.int32)), 1, 0, LOSS_TYPE[loss_type], 0, pp(arrs[0]), pp(arrs[1]), pp(arrs[2]), pp(grads[0]), pp(grads[1]),
                        pp(grads[2]), _fp(lab2), _fp(rt2), ctypes.c_void_p(0), _fp(ct), _fp(sums))
    focal, n_pos, sum_ct, reg_raw, bce, _ = sums
    np.testing.assert_allclose([focal / max(n_pos, 1), reg_raw / sum_ct, bce / max(n_pos, 1)], [l_cls.item(), l_reg.item(), l_ctr.item()], rtol=2e-5)
    for l in range(n_levels):
        np.testing.assert_allclose(grads[0][l] / max(n_pos, 1), cls[l].grad.numpy(), rtol=3e-4, atol=1e-7)
        np.testing.assert_allclose(2.0 * grads[1][l] / sum_ct, reg[l].grad.numpy(), rtol=3e-4, atol=1e-7)
        np.testing.assert_allclose(3.0 * grads[2][l] / max(n_pos, 1), ctr[l].grad.numpy(), rtol=3e-4, atol=1e-7)


def _two_rank_loss_worker(rank, world, port, shim_file, golden_file, names, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    patch_host_ops(ctypes.CDLL(shim_file), setattr)
    g = np.load(golden_file)
    c = load_case(g, names[rank])
    mod = _module(c, world_size=world)
    cls, reg, ctr = ([torch.tensor(a, requires_grad=True) for a in c[k]] for k in ("cls", "reg", "ctr"))
    locs = mod.compute_locations(cls)
    masks = mod.compute_padding_masks(locs, [tuple(int(v) for v in s) for s in g[f"{names[rank]}/sizes"]]) if c["batch"] > 1 else None
    _, _, losses = mod._forward_train(locs, cls, reg, ctr, [torch.tensor(t) for t in c["gt"]], masks)
    (losses["loss_cls"] + 2.0 * losses["loss_reg"] + 3.0 * losses["loss_centerness"]).backward()
    q.put((rank, [losses[k].item() for k in ("loss_cls", "loss_reg", "loss_centerness")], [float(t.grad.abs().sum()) for t in (cls[0], reg[0], ctr[0])]))
    dist.destroy_process_group()


def test_two_rank_loss_normalisers_end_to_end_gloo(shim_path, golden, golden_dir):
    """8(e) C3 end to end at world size 2 (gloo, host build of the kernels): rank 0 = case aabb_iou, rank 1 = case aabb_empty.  Each rank's losses and
    gradients must equal its single-rank golden values rescaled by (own normaliser) / (average normaliser over the ranks), loss.py:541-576."""
    import socket
    import torch.multiprocessing as mp
    names = ("aabb_iou", "aabb_empty")
    single = []
    for nm in names:                                   # the single-rank raw sums behind the goldens
        c = load_case(golden, nm)
        N = c["batch"]
        lab = np.stack(per_scene(c["labels"], c["n_per"], N))
        m = np.concatenate([mm.reshape(N, -1) for mm in c["mask"]], 1)
        rt = np.stack(per_scene(c["reg_targets"], c["n_per"], N))
        pos = (lab > 0) & m
        ct = O.centerness_targets(torch.tensor(rt[pos])).numpy()
        single.append((float(pos.sum()), float(ct.sum(dtype=np.float64)), c["losses"], [np.abs(c[k][0]).sum() for k in ("dcls", "dreg", "dctr")]))
    avg_pos = max((single[0][0] + single[1][0]) / 2, 1.0)
    avg_ct = (single[0][1] + single[1][1]) / 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    gfile = os.path.join(golden_dir, "fcos_loss.npz")
    ps = [ctx.Process(target=_two_rank_loss_worker, args=(r, 2, port, shim_path, gfile, names, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for rank, losses, gsum in out:
        n_pos, s_ct, want, gwant = single[rank]
        f_pos, f_ct = max(n_pos, 1.0) / avg_pos, s_ct / avg_ct
        np.testing.assert_allclose(losses, [want[0] * f_pos, want[1] * f_ct, want[2] * f_pos], rtol=2e-5)
        np.testing.assert_allclose(gsum, [gwant[0] * f_pos, gwant[1] * f_ct, gwant[2] * f_pos], rtol=2e-4)
    assert abs(single[0][0] - single[1][0]) > 10           # the two ranks really have different normalisers
