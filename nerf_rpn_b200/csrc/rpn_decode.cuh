// Per-candidate box decoding for the RPN (device functions; also host-compiled by tests/host_shim).
//   AABB : coder/AABB_coder.py:86-137 (decode_single)
//   OBB  : coder/midpoint_offset_coder.py:160-223 (delta_sp2bbox) + coder/misc.py:5-43 (rectpoly2obb,
//          regular_obb, regular_theta with pi = 3.141592)
//   score: torch.sigmoid (rpn.py:342)
// Arithmetic contract as in box_iou.cuh: separately rounded fp32 ops in the reference's order; exp, atan2,
// sin and cos evaluated in fp64 and rounded once; sums left to right.
#pragma once
#ifndef NRPN_SKIP_COMMON
#include "common.cuh"
#endif

namespace nrpn {

__device__ __forceinline__ float f_exp(float v) { return (float)exp((double)v); }

__device__ __forceinline__ float sigmoid_ref(float x) {
    const float t = f_exp(-x);
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, t));
}

// anchor (x1,y1,z1,x2,y2,z2) + deltas (dx,dy,dz,dw,dh,dd) -> box (x1,y1,z1,x2,y2,z2)
__device__ __forceinline__ void decode_aabb(const float* __restrict__ an, const float* __restrict__ d, float* __restrict__ out) {
    const float clip = (float)7.600902459542082;   // math.log(2000.0), AABB_coder.py:66
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float size = __fsub_rn(an[3 + k], an[k]);
        const float ctr = __fadd_rn(an[k], __fmul_rn(0.5f, size));
        float dw = d[3 + k];
        if (dw > clip) dw = clip;                   // torch.clamp(max=) keeps NaN
        const float pc = __fadd_rn(__fmul_rn(d[k], size), ctr);
        const float ps = __fmul_rn(f_exp(dw), size);
        const float half = __fmul_rn(0.5f, ps);
        out[k] = __fsub_rn(pc, half);
        out[3 + k] = __fadd_rn(pc, half);
    }
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) {   // torch.clamp(min, max); NaN passes through
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

// anchor + 8 midpoint-offset deltas (dx,dy,dz,dw,dh,dd,da,db) -> (x,y,z,w,h,d,theta)
__device__ __forceinline__ void decode_obb(const float* __restrict__ an, const float* __restrict__ d, float* __restrict__ out) {
    const float max_ratio = (float)4.135166556742356;   // abs(log(16/1000)), midpoint_offset_coder.py:178
    const float dw = clampf(d[3], -max_ratio, max_ratio);
    const float dh = clampf(d[4], -max_ratio, max_ratio);
    const float dd = clampf(d[5], -max_ratio, max_ratio);
    const float px = __fmul_rn(__fadd_rn(an[0], an[3]), 0.5f);
    const float py = __fmul_rn(__fadd_rn(an[1], an[4]), 0.5f);
    const float pz = __fmul_rn(__fadd_rn(an[2], an[5]), 0.5f);
    const float pw = __fsub_rn(an[3], an[0]);
    const float ph = __fsub_rn(an[4], an[1]);
    const float pd = __fsub_rn(an[5], an[2]);
    const float gw = __fmul_rn(pw, f_exp(dw));
    const float gh = __fmul_rn(ph, f_exp(dh));
    const float gd = __fmul_rn(pd, f_exp(dd));
    const float gx = __fadd_rn(px, __fmul_rn(pw, d[0]));
    const float gy = __fadd_rn(py, __fmul_rn(ph, d[1]));
    const float gz = __fadd_rn(pz, __fmul_rn(pd, d[2]));
    const float hw = __fmul_rn(gw, 0.5f), hh = __fmul_rn(gh, 0.5f);
    const float x1 = __fsub_rn(gx, hw), y1 = __fsub_rn(gy, hh), x2 = __fadd_rn(gx, hw), y2 = __fadd_rn(gy, hh);
    const float da = clampf(d[6], -0.5f, 0.5f), db = clampf(d[7], -0.5f, 0.5f);
    const float ga = __fadd_rn(gx, __fmul_rn(da, gw)), ga_ = __fsub_rn(gx, __fmul_rn(da, gw));
    const float gb = __fadd_rn(gy, __fmul_rn(db, gh)), gb_ = __fsub_rn(gy, __fmul_rn(db, gh));
    float qx[4] = {ga, x2, ga_, x1};
    float qy[4] = {y1, gb, y2, gb_};
    // rectangularise: scale every centred vertex to the longest diagonal
    float cx[4], cy[4], dl[4];
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        cx[i] = __fsub_rn(qx[i], gx); cy[i] = __fsub_rn(qy[i], gy);
        dl[i] = sqrtf(__fadd_rn(__fmul_rn(cx[i], cx[i]), __fmul_rn(cy[i], cy[i])));
        if (i == 0 || dl[i] > mx || dl[i] != dl[i]) mx = dl[i];      // torch.max propagates NaN
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float sc = __fdiv_rn(mx, dl[i]);
        qx[i] = __fadd_rn(__fmul_rn(cx[i], sc), gx);
        qy[i] = __fadd_rn(__fmul_rn(cy[i], sc), gy);
    }
    // rectpoly2obb (misc.py:23-43)
    const float eps = 1e-7f;
    const float ty = -__fsub_rn(qy[1], qy[0]);
    const float tx = __fadd_rn(__fsub_rn(qx[1], qx[0]), eps);
    const float theta = (float)atan2((double)ty, (double)tx);
    const float Cos = (float)cos((double)theta), Sin = (float)sin((double)theta);
    const float nSin = -Sin;
    const float xm = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(qx[0], qx[1]), qx[2]), qx[3]), 4.0f);
    const float ym = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(qy[0], qy[1]), qy[2]), qy[3]), 4.0f);
    float xmin = 0.f, xmax = 0.f, ymin = 0.f, ymax = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float ux = __fsub_rn(qx[i], xm), uy = __fsub_rn(qy[i], ym);
        const float rx = __fadd_rn(__fmul_rn(ux, Cos), __fmul_rn(uy, nSin));   // row . Matrix^T[:,0]
        const float ry = __fadd_rn(__fmul_rn(ux, Sin), __fmul_rn(uy, Cos));
        if (i == 0) { xmin = xmax = rx; ymin = ymax = ry; }
        else {
            if (rx < xmin || rx != rx) xmin = rx;
            if (rx > xmax || rx != rx) xmax = rx;
            if (ry < ymin || ry != ry) ymin = ry;
            if (ry > ymax || ry != ry) ymax = ry;
        }
    }
    const float w = __fsub_rn(xmax, xmin), h = __fsub_rn(ymax, ymin);
    // regular_obb / regular_theta (misc.py:5-20), pi = 3.141592
    const float pi = (float)3.141592, half_pi = (float)(3.141592 / 2), start = (float)(-3.141592 / 2);
    const bool wh = w > h;
    const float wr = wh ? w : h, hr = wh ? h : w;
    float th = wh ? theta : __fadd_rn(theta, half_pi);
    th = __fsub_rn(th, start);
    float md = fmodf(th, pi);                        // torch.remainder: result takes the divisor's sign
    if (md != 0.0f && (md < 0.0f)) md = __fadd_rn(md, pi);
    th = __fadd_rn(md, start);
    out[0] = xm; out[1] = ym; out[2] = gz; out[3] = wr; out[4] = hr; out[5] = gd; out[6] = th;
}

}  // namespace nrpn
