"""Pure-torch statement of what nrpn_conv3d_fprop / nrpn_pack_stem_input compute (TEST helper, CPU or GPU).
Used to check the host-side packing logic without a GPU and as the fp32 reference for the CUDA kernels."""
import torch
import torch.nn.functional as F


def emulate_pack_stem(grid: torch.Tensor) -> torch.Tensor:
    """(N,4,X,Y,Z) fp32 -> (N,X2,Y2,Z2+1,64) fp32 with the layout of csrc/pointwise.cu:pack_stem_kernel."""
    n, c, X, Y, Z = grid.shape
    X2, Y2, Z2 = (X + 1) // 2, (Y + 1) // 2, (Z + 1) // 2
    g = F.pad(grid, (0, 2 * Z2 - Z, 0, 2 * Y2 - Y, 0, 2 * X2 - X))
    # s2d block: channel = ((rx*2+ry)*2+rz)*4 + c
    b = g.view(n, c, X2, 2, Y2, 2, Z2, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(n, X2, Y2, Z2, 32)
    zero = torch.zeros_like(b[:, :, :, :1])
    lo = torch.cat([zero, b], 3)          # block k-1 at row k
    hi = torch.cat([b, zero], 3)          # block k   at row k
    return torch.cat([lo, hi], -1)


def emulate_conv(x, w_packed, taps, shift, out_dims, stride=1, relu=False, res=None):
    """x (N,X,Y,Z,Cin) float, w_packed (taps,CoutPad,Cin) float -> (N,Xo,Yo,Zo,CoutPad) float."""
    n, X, Y, Z, cin = x.shape
    xo, yo, zo = out_dims
    out = torch.zeros((n, xo, yo, zo, w_packed.shape[1]), dtype=torch.float32, device=x.device)
    P = 8
    xp = F.pad(x, (0, 0, P, P + xo * stride, P, P + yo * stride, P, P + zo * stride))
    for t, (dx, dy, dz) in enumerate(taps):
        sl = xp[:, P + dx: P + dx + xo * stride: stride, P + dy: P + dy + yo * stride: stride,
                P + dz: P + dz + zo * stride: stride]
        out += sl.float() @ w_packed[t].float().t()
    out = out + shift.float()
    if res is not None:
        r = res.float()
        if r.shape[1:4] != out.shape[1:4]:
            r = F.interpolate(r.permute(0, 4, 1, 2, 3), size=(xo, yo, zo), mode="nearest").permute(0, 2, 3, 4, 1)
        out[..., : r.shape[-1]] += r
    if relu:
        out = out.clamp_min(0)
    return out


def emulate_pack_stem_s1(grid: torch.Tensor) -> torch.Tensor:
    """(N,4,X,Y,Z) -> (N,X,Y+1,Z,64) with the layout of csrc/pointwise.cu:pack_stem_s1_kernel."""
    n, c, X, Y, Z = grid.shape
    g = F.pad(grid, (3, 3, 1, 1))                       # z by 3, y by 1 on both sides
    out = torch.zeros((n, X, Y + 1, Z, 64), dtype=grid.dtype, device=grid.device)
    for yy in range(2):
        for zz in range(7):
            # row yp holds input y = yp - 1 + yy  -> padded index yp + yy ; z neighbour z + zz - 3 -> padded z + zz
            sl = g[:, :, :, yy: yy + Y + 1, zz: zz + Z]                 # (n, c, X, Y+1, Z)
            ch = (yy * 7 + zz) * 4
            out[..., ch:ch + 4] = sl.permute(0, 2, 3, 4, 1)
    return out
