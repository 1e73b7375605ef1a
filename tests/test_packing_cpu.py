"""CPU: host-side weight / input packing against torch's own conv3d (fp32)."""
import torch
import torch.nn.functional as F

from nerf_rpn_b200 import packing
from tests.emulate import emulate_conv, emulate_pack_stem


def test_stem_space_to_depth_equals_conv7_stride2():
    torch.manual_seed(0)
    for dims in [(12, 10, 14), (11, 9, 13)]:              # even and odd extents
        x = torch.randn(2, 4, *dims)
        w = torch.randn(64, 4, 7, 7, 7) * 0.05
        ref = F.conv3d(x, w, stride=2, padding=3).permute(0, 2, 3, 4, 1)
        wp, taps = packing.pack_stem_weight(w)
        packed = emulate_pack_stem(x)
        out_dims = tuple((d + 1) // 2 for d in dims)
        got = emulate_conv(packed, wp.float(), taps, torch.zeros(64), out_dims)
        # bf16 rounding of the packed weights only
        ref_bf = F.conv3d(x, w.to(torch.bfloat16).float(), stride=2, padding=3).permute(0, 2, 3, 4, 1)
        assert got.shape == ref.shape
        torch.testing.assert_close(got, ref_bf, rtol=1e-4, atol=1e-4)


def test_generic_conv_packing_3x3_and_1x1_stride2():
    torch.manual_seed(1)
    x = torch.randn(1, 64, 7, 6, 5)
    w3 = torch.randn(72, 64, 3, 3, 3) * 0.05
    wp, taps = packing.pack_conv_weight(w3)
    assert wp.shape == (27, 128, 64) and len(taps) == 27
    got = emulate_conv(x.permute(0, 2, 3, 4, 1), wp.float(), taps, torch.zeros(128), (7, 6, 5))[..., :72]
    ref = F.conv3d(x, w3.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 4, 1)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)
    w1 = torch.randn(64, 64, 1, 1, 1) * 0.1
    wp, taps = packing.pack_conv_weight(w1)
    got = emulate_conv(x.permute(0, 2, 3, 4, 1), wp.float(), taps, torch.zeros(64), (4, 3, 3), stride=2)
    ref = F.conv3d(x, w1.to(torch.bfloat16).float(), stride=2).permute(0, 2, 3, 4, 1)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


def test_bn_folding():
    torch.manual_seed(2)
    bn = torch.nn.BatchNorm3d(8).eval()
    with torch.no_grad():
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2); bn.weight.normal_(); bn.bias.normal_()
    x = torch.randn(2, 8, 3, 3, 3)
    s, b = packing.fold_bn(bn)
    torch.testing.assert_close(x * s.view(1, -1, 1, 1, 1) + b.view(1, -1, 1, 1, 1), bn(x), rtol=1e-5, atol=1e-5)


def test_stem_stride1_packing_equals_conv7():
    from tests.emulate import emulate_pack_stem_s1
    torch.manual_seed(3)
    x = torch.randn(1, 4, 9, 8, 10)
    w = torch.randn(64, 4, 7, 7, 7) * 0.05
    wp, taps = packing.pack_stem_s1_weight(w)
    assert wp.shape == (28, 64, 64)
    got = emulate_conv(emulate_pack_stem_s1(x), wp.float(), taps, torch.zeros(64), (9, 8, 10))
    ref = F.conv3d(x, w.to(torch.bfloat16).float(), padding=3).permute(0, 2, 3, 4, 1)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


def test_weight_split_hi_lo_reconstructs_fp32():
    """packing.split_hi_lo (nrpn_conv_desc.wsplit): hi + lo == w to ~2^-22 relative for fp16, layout (taps, 2, CoutPad, Cin)."""
    import torch
    from nerf_rpn_b200 import packing
    torch.manual_seed(3)
    w = torch.randn(72, 128, 3, 3, 3) * 0.05
    wp, taps = packing.pack_conv_weight(w, dtype=torch.float16, split=True)
    single, taps1 = packing.pack_conv_weight(w, dtype=torch.float16)
    assert taps == taps1 and wp.shape == (27, 2, 128, 128) and single.shape == (27, 128, 128)
    assert torch.equal(wp[:, 0], single)                                   # plane 0 = the ordinary rounded weights
    exact, _ = packing.pack_conv_weight(w, dtype=torch.float32)
    rec = wp[:, 0].float() + wp[:, 1].float()
    # lo lands in fp16's subnormal range for weights this small: absolute error <= half a subnormal step (2^-25) + fp32 rounding
    assert (rec - exact).abs().max().item() <= 2 ** -25 + 1e-9
    assert (single.float() - exact).abs().max() > 100 * (rec - exact).abs().max()
