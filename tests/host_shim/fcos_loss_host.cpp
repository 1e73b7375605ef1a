// Host build of nerf_rpn_b200/csrc/fcos_loss.cuh (TEST ONLY): the per-element device functions of the FCOS loss kernels compiled as
// plain C++ to check their logic against oracle/fcos_loss_oracle.py and the reference's golden vectors on GPU-less boxes.
// The loops below stand in for the thread grid of fcos_loss.cu; the product never uses this file.
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <cstring>
#define __device__
#define __host__
#include "../../nerf_rpn_b200/csrc/fcos_loss.cuh"

using namespace nrpn;

extern "C" void shim_fcos_targets(const float* loc, const int* begin, int n_levels, const float* radius_stride, const float* size_lo,
                                  const float* size_hi, int norm, const float* norm_div, const float* gt, int n_gt, int gt_dim,
                                  float* labels, float* reg_targets) {
    const int D = gt_dim == 7 ? 8 : 6;
    FcosGt* g = new FcosGt[n_gt > 0 ? n_gt : 1];
    for (int i = 0; i < n_gt; ++i) fcos_gt_prepare(gt + (size_t)i * gt_dim, gt_dim, g[i]);
    for (int i = 0; i < begin[n_levels]; ++i) {
        int lvl = 0;
        while (lvl + 1 < n_levels && i >= begin[lvl + 1]) ++lvl;
        FcosBest best;
        fcos_best_init(best);
        for (int k = 0; k < n_gt; ++k) fcos_target_update(g[k], loc + (size_t)i * 3, radius_stride[lvl], size_lo[lvl], size_hi[lvl], best);
        labels[i] = (n_gt > 0 && best.area != kFcosInf) ? 1.f : 0.f;
        for (int k = 0; k < D; ++k) {
            float v = best.reg[k];
            if (k < 6 && norm) v = v / norm_div[lvl];
            reg_targets[(size_t)i * D + k] = n_gt > 0 ? v : 0.f;
        }
    }
    delete[] g;
}

extern "C" void shim_fcos_loss(int n_levels, const int* n_points, int n_img, int use_obb, int loss_type, int add_l1, const float* const* cls,
                               const float* const* reg, const float* const* ctr, float* const* dcls, float* const* dreg, float* const* dctr,
                               const float* labels, const float* reg_targets, const uint8_t* mask, float* ct_out, double* sums) {
    FcosLossDev P;
    std::memset(&P, 0, sizeof(P));
    P.n_levels = n_levels; P.n_img = n_img; P.D = use_obb ? 8 : 6;
    P.loss_type = loss_type; P.use_obb = use_obb; P.add_l1 = add_l1; P.want_grad = dcls != nullptr;
    int begin = 0;
    for (int l = 0; l < n_levels; ++l) {
        P.cls[l] = cls[l]; P.reg[l] = reg[l]; P.ctr[l] = ctr[l];
        if (dcls) { P.dcls[l] = dcls[l]; P.dreg[l] = dreg[l]; P.dctr[l] = dctr[l]; }
        P.begin[l] = begin; begin += n_points[l];
    }
    for (int l = n_levels; l <= kFcosMaxLevels; ++l) P.begin[l] = begin;
    P.total = begin; P.labels = labels; P.rt = reg_targets; P.mask = mask; P.ct_out = ct_out;
    double acc[kFlSums] = {0, 0, 0, 0, 0, 0};
    for (long e = 0; e < (long)n_img * P.total; ++e) fcos_loss_element(P, e, acc);
    for (int k = 0; k < kFlSums; ++k) sums[k] = acc[k];
}
