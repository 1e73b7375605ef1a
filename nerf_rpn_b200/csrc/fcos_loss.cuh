// Per-element arithmetic of the FCOS training loss: everything FCOSLossComputation
// (nerf_rpn/model/fcos/loss.py:185-591) computes per location or per ground-truth box, as __host__ __device__ functions so that
// tests/host_shim can run the same code on a GPU-less box against the oracle.  The kernels around them are in fcos_loss.cu.
//   fcos_gt_prepare        : compute_targets_for_locations[_obb] (:318-441) + encode_fcos_obb (fcos/utils.py:64-108) + box2corners_th
//                            (oriented_iou_loss.py:6-36) -- the location-independent part: the box's AABB, volume, midpoint offsets
//   fcos_target_update     : the per-(location, GT) part: regression distances, get_sample_region (:213-260), the size-of-interest test,
//                            first-minimum volume (:357-361, 426-430)
//   fcos_focal             : torchvision.ops.sigmoid_focal_loss (alpha 0.25, gamma 2), value and d/dlogit
//   fcos_centerness_target : compute_centerness_targets (:443-450)
//   fcos_bce_logits        : nn.BCEWithLogitsLoss, value and d/dlogit
//   fcos_aabb_iou_loss     : IOULoss.forward (:78-131) per box, value and d/dpred
//   fcos_smooth_l1         : nn.SmoothL1Loss (beta 1)
#pragma once
#include <math.h>
#include <stdint.h>

namespace nrpn {

constexpr float kFcosInf = 100000000.0f;          // INF of fcos/loss.py:23 (exact in fp32)

struct FcosGt {
    float lo[3], hi[3];     // the AABB itself, or the AABB of the OBB (footprint corners' extrema, z -+ d/2)
    float alpha, beta;      // midpoint offsets of encode_fcos_obb (0 for an AABB)
    float volume;           // (hi - lo) product: the tie breaker of overlapping boxes
};

// gt: 6 floats (x1,y1,z1,x2,y2,z2) or 7 (x,y,z,w,h,d,theta).  sin / cos through fp64, rounded once (like every other box kernel here).
__host__ __device__ inline void fcos_gt_prepare(const float* gt, int dim, FcosGt& g) {
    if (dim == 6) {
        for (int k = 0; k < 3; ++k) { g.lo[k] = gt[k]; g.hi[k] = gt[3 + k]; }
        g.alpha = 0.f; g.beta = 0.f;
    } else {
        const float x = gt[0], y = gt[1], z = gt[2], w = gt[3], h = gt[4], d = gt[5], th = gt[6];
        const float co = (float)cos((double)th), si = (float)sin((double)th);
        const float sx[4] = {0.5f, -0.5f, -0.5f, 0.5f}, sy[4] = {0.5f, 0.5f, -0.5f, -0.5f};
        float xs[4], ys[4];
        float xmax = -INFINITY, ymax = -INFINITY, xmin = INFINITY, ymin = INFINITY;
        for (int k = 0; k < 4; ++k) {
            const float x4 = sx[k] * w, y4 = sy[k] * h;
            xs[k] = (x4 * co + y4 * (-si)) + x;                      // corners @ [[cos, sin], [-sin, cos]], then += centre
            ys[k] = (x4 * si + y4 * co) + y;
            xmax = fmaxf(xmax, xs[k]); xmin = fminf(xmin, xs[k]);
            ymax = fmaxf(ymax, ys[k]); ymin = fminf(ymin, ys[k]);
        }
        float vx = -INFINITY, vy = INFINITY;                        // the corner on the top edge / on the right edge of the AABB
        for (int k = 0; k < 4; ++k) {
            const float xt = (ymax - ys[k] > 0.1f) ? -1e6f : xs[k];
            const float yt = (xmax - xs[k] > 0.1f) ? 1e6f : ys[k];
            vx = fmaxf(vx, xt); vy = fminf(vy, yt);
        }
        // torch.isclose(a, b): |a - b| <= 1e-8 + 1e-5 |b|  (theta too small to be stable: use the AABB's corner)
        const bool cx = vx == xmax || fabsf(vx - xmax) <= 1e-8f + 1e-5f * fabsf(xmax);
        const bool cy = vy == ymin || fabsf(vy - ymin) <= 1e-8f + 1e-5f * fabsf(ymin);
        if (cx && cy) { vx = xmax; vy = ymin; }
        g.alpha = (vx - x) / (xmax - xmin);
        g.beta = (vy - y) / (ymax - ymin);
        g.lo[0] = xmin; g.lo[1] = ymin; g.lo[2] = z - d / 2.f;
        g.hi[0] = xmax; g.hi[1] = ymax; g.hi[2] = z + d / 2.f;
    }
    g.volume = (g.hi[0] - g.lo[0]) * (g.hi[1] - g.lo[1]) * (g.hi[2] - g.lo[2]);
}

struct FcosBest {
    float area;             // running minimum (first minimum wins, like torch.min(dim=1))
    float reg[8];           // the regression target of that box: l, t, f, r, b, ba (distances to lo / hi), alpha, beta
};

__host__ __device__ inline void fcos_best_init(FcosBest& b) {
    b.area = INFINITY;
    for (int k = 0; k < 8; ++k) b.reg[k] = 0.f;
}

// One (location, ground-truth) pair.  radius_stride = fpn_stride * center_sampling_radius of the location's level (<= 0: no centre
// sampling, every location strictly inside the box counts); [size_lo, size_hi] = the level's object_sizes_of_interest row.
__host__ __device__ inline void fcos_target_update(const FcosGt& g, const float* p, float radius_stride, float size_lo, float size_hi, FcosBest& best) {
    float reg[6];
    for (int k = 0; k < 3; ++k) { reg[k] = p[k] - g.lo[k]; reg[3 + k] = g.hi[k] - p[k]; }
    float mn, mx = reg[0];
    for (int k = 1; k < 6; ++k) mx = fmaxf(mx, reg[k]);
    if (radius_stride > 0.f) {
        mn = INFINITY;
        for (int k = 0; k < 3; ++k) {
            const float c = (g.lo[k] + g.hi[k]) / 2.f;
            const float cmin = c - radius_stride, cmax = c + radius_stride;
            const float clo = cmin > g.lo[k] ? cmin : g.lo[k];       // limit the sample region to the box
            const float chi = cmax > g.hi[k] ? g.hi[k] : cmax;
            mn = fminf(mn, fminf(p[k] - clo, chi - p[k]));
        }
    } else {
        mn = reg[0];
        for (int k = 1; k < 6; ++k) mn = fminf(mn, reg[k]);
    }
    const bool inside = mn > 0.f;
    const bool cared = mx >= size_lo && mx <= size_hi;
    const float area = (inside && cared) ? g.volume : kFcosInf;
    if (area < best.area) {
        best.area = area;
        for (int k = 0; k < 6; ++k) best.reg[k] = reg[k];
        best.reg[6] = g.alpha; best.reg[7] = g.beta;
    }
}

__host__ __device__ inline float fcos_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// softplus(-|x|) part of the numerically stable BCE-with-logits: max(x, 0) - x t + log1p(exp(-|x|)).
__host__ __device__ inline float fcos_bce_logits(float x, float t, float* dx) {
    const float p = fcos_sigmoid(x);
    *dx = p - t;
    return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}

// sigmoid_focal_loss(inputs, targets, alpha = 0.25, gamma = 2) for a binary target.
__host__ __device__ inline float fcos_focal(float x, bool positive, float* dx) {
    const float p = fcos_sigmoid(x);
    float ce_d;
    const float ce = fcos_bce_logits(x, positive ? 1.f : 0.f, &ce_d);
    if (positive) {
        const float q = 1.f - p;                                       // 1 - p_t
        *dx = 0.25f * q * q * (ce_d - 2.f * p * ce);                    // d[ce q^2] = q^2 ce' + ce 2 q (-p q)
        return 0.25f * (ce * (q * q));
    }
    *dx = 0.75f * p * p * (ce_d + 2.f * (1.f - p) * ce);                // d[ce p^2] = p^2 ce' + ce 2 p (p (1 - p))
    return 0.75f * (ce * (p * p));
}

__host__ __device__ inline float fcos_centerness_target(const float* rt) {
    const float a = fminf(rt[0], rt[3]) / fmaxf(rt[0], rt[3]);
    const float b = fminf(rt[1], rt[4]) / fmaxf(rt[1], rt[4]);
    const float c = fminf(rt[2], rt[5]) / fmaxf(rt[2], rt[5]);
    return sqrtf((a * b) * c);
}

__host__ __device__ inline float fcos_smooth_l1(float d, float* dd) {
    const float a = fabsf(d);
    if (a < 1.f) { *dd = d; return 0.5f * d * d; }
    *dd = d > 0.f ? 1.f : -1.f;
    return a - 0.5f;
}

// IOULoss.forward for one box: p / t = (left, top, front, right, bottom, back).  type 1 = -log(iou), 2 = 1 - iou, 3 = 1 - giou.
// torch.min / torch.max send half of the gradient to each side on a tie.
__host__ __device__ inline float fcos_aabb_iou_loss(const float* p, const float* t, int type, float* dp) {
    float sp[3], inter[3], outer[3], dmin[6];
    for (int k = 0; k < 3; ++k) {
        sp[k] = p[k] + p[3 + k];
        inter[k] = fminf(p[k], t[k]) + fminf(p[3 + k], t[3 + k]);
        outer[k] = fmaxf(p[k], t[k]) + fmaxf(p[3 + k], t[3 + k]);
        dmin[k] = p[k] < t[k] ? 1.f : (p[k] == t[k] ? 0.5f : 0.f);
        dmin[3 + k] = p[3 + k] < t[3 + k] ? 1.f : (p[3 + k] == t[3 + k] ? 0.5f : 0.f);
    }
    const float tv = (t[0] + t[3]) * (t[1] + t[4]) * (t[2] + t[5]);
    const float pv = sp[0] * sp[1] * sp[2];
    const float ac = outer[0] * outer[1] * outer[2] + 1e-7f;
    const float vi = inter[0] * inter[1] * inter[2];
    const float vu = tv + pv - vi;
    const float iou = (vi + 1.0f) / (vu + 1.0f);
    for (int j = 0; j < 6; ++j) {
        const int k = j % 3, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
        const float dpv = sp[k1] * sp[k2];
        const float dvi = dmin[j] * inter[k1] * inter[k2];
        const float dvu = dpv - dvi;
        const float diou = (dvi * (vu + 1.0f) - (vi + 1.0f) * dvu) / ((vu + 1.0f) * (vu + 1.0f));
        if (type == 1) dp[j] = -diou / iou;
        else if (type == 2) dp[j] = -diou;
        else {
            const float dac = (1.f - dmin[j]) * outer[k1] * outer[k2];
            dp[j] = -(diou + (dvu * ac - vu * dac) / (ac * ac));       // giou = iou - 1 + vu / ac
        }
    }
    if (type == 1) return -logf(iou);
    if (type == 2) return 1.f - iou;
    return 1.f - (iou - (ac - vu) / ac);
}

// ---------------------------------------------------------------------------------------------- one (scene, location) of nrpn_fcos_loss
constexpr int kFcosMaxLevels = 4;          // == NRPN_RPN_MAX_LEVELS
constexpr int kFlSums = 6;

struct FcosLossDev {
    int n_levels, n_img, D, total;            // total = locations per scene over all levels
    int begin[kFcosMaxLevels + 1];
    const float* cls[kFcosMaxLevels]; const float* reg[kFcosMaxLevels]; const float* ctr[kFcosMaxLevels];
    float* dcls[kFcosMaxLevels]; float* dreg[kFcosMaxLevels]; float* dctr[kFcosMaxLevels];
    const float* labels; const float* rt; const uint8_t* mask; float* ct_out;
    int loss_type, use_obb, add_l1, want_grad;
};

// acc: 0 focal, 1 positives, 2 centerness targets, 3 weighted regression loss, 4 centerness BCE, 5 weighted alpha / beta smooth-L1.
__host__ __device__ inline void fcos_loss_element(const FcosLossDev& P, long e, double* acc) {
    const int n = (int)(e / P.total), q = (int)(e % P.total);
    int lvl = 0;
    while (lvl + 1 < P.n_levels && q >= P.begin[lvl + 1]) ++lvl;
    const int pl = P.begin[lvl + 1] - P.begin[lvl], pi = q - P.begin[lvl];
    const size_t s1 = (size_t)n * pl + pi;                                // (N, 1, P_l)
    const size_t sd = (size_t)n * P.D * pl + pi;                          // (N, D, P_l): channel c at + c * pl
    const bool kept = P.mask == nullptr || P.mask[e] != 0;
    const bool pos = kept && P.labels[e] > 0.f;
    float dcls = 0.f, dctr = 0.f, ct = 0.f;
    float dreg[8];
    for (int k = 0; k < 8; ++k) dreg[k] = 0.f;
    if (kept) acc[0] += (double)fcos_focal(P.cls[lvl][s1], pos, &dcls);
    if (pos) {
        float t[8], pr[8];
        for (int k = 0; k < 8; ++k) {
            t[k] = k < P.D ? P.rt[(size_t)e * P.D + k] : 0.f;
            pr[k] = k < P.D ? P.reg[lvl][sd + (size_t)k * pl] : 0.f;
        }
        ct = fcos_centerness_target(t);
        acc[1] += 1.0;
        acc[2] += (double)ct;
        acc[4] += (double)fcos_bce_logits(P.ctr[lvl][s1], ct, &dctr);
        if (P.loss_type == 0) {
            float s = 0.f;
            for (int k = 0; k < 8; ++k)
                if (k < P.D) { float d; s += fcos_smooth_l1(pr[k] - t[k], &d) * ct; dreg[k] = d * ct; }
            acc[3] += (double)s;
        } else if (!P.use_obb) {
            float d6[6];
            acc[3] += (double)(fcos_aabb_iou_loss(pr, t, P.loss_type, d6) * ct);
            for (int k = 0; k < 6; ++k) dreg[k] = d6[k] * ct;
        } else if (P.add_l1) {
            float s = 0.f;
            for (int k = 6; k < 8; ++k) { float d; s += fcos_smooth_l1(pr[k] - t[k], &d) * ct; dreg[k] = d * ct; }
            acc[5] += (double)s;
        }
    }
    if (P.ct_out) P.ct_out[e] = ct;
    if (P.want_grad) {
        P.dcls[lvl][s1] = dcls;
        P.dctr[lvl][s1] = dctr;
        for (int k = 0; k < 8; ++k)
            if (k < P.D) P.dreg[lvl][sd + (size_t)k * pl] = dreg[k];
    }
}

}  // namespace nrpn
