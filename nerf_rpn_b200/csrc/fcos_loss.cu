// FCOS training loss on the device: the fused replacement of FCOSLossComputation
// (nerf_rpn/model/fcos/loss.py:185-591).  Two kernels, both HBM / latency bound on a few hundred kB:
//   fcos_targets_kernel : one thread per location, ground truth streamed through shared memory in chunks (prepared per CTA: AABB of
//                         the box, midpoint offsets, volume); the reference's (locations, G, 8) regression tensor, the (locations, G)
//                         masks and the volume matrix are never materialised.  16 B read per location, 4 + 24|32 B written.
//   fcos_loss_kernel    : one thread per (scene, location): focal term of every location, and on the positives the centerness target,
//                         its BCE, the centerness-weighted regression loss, all with their gradients written into the NCDHW layout the
//                         head produced; six fp64 partial sums per CTA, added in a fixed order by fcos_loss_final_kernel (deterministic).
// The per-element arithmetic is in fcos_loss.cuh (shared with tests/host_shim).
#include <string.h>
#include "common.cuh"
#include "fcos_loss.cuh"

namespace nrpn {

constexpr int kFlThreads = 256;
constexpr int kFlGtChunk = 256;
constexpr int kFlMaxBlocks = 592;          // 4 per SM
static_assert(kFcosMaxLevels == NRPN_RPN_MAX_LEVELS, "level count");

struct FcosTgtDev {
    int n_levels, total;
    int begin[NRPN_RPN_MAX_LEVELS + 1];
    float radius_stride[NRPN_RPN_MAX_LEVELS], size_lo[NRPN_RPN_MAX_LEVELS], size_hi[NRPN_RPN_MAX_LEVELS];
    int norm[NRPN_RPN_MAX_LEVELS];
    float norm_div[NRPN_RPN_MAX_LEVELS];
};

__global__ void __launch_bounds__(kFlThreads) fcos_targets_kernel(const FcosTgtDev P, const float* __restrict__ loc, const float* __restrict__ gt, int n_gt,
                                                                  int gt_dim, float* __restrict__ labels, float* __restrict__ reg_targets) {
    __shared__ FcosGt sg[kFlGtChunk];
    const int i = blockIdx.x * kFlThreads + threadIdx.x;
    const bool live = i < P.total;
    const int D = gt_dim == 7 ? 8 : 6;
    int lvl = 0;
    float p[3] = {0.f, 0.f, 0.f};
    if (live) {
        while (lvl + 1 < P.n_levels && i >= P.begin[lvl + 1]) ++lvl;
        p[0] = loc[(size_t)i * 3 + 0]; p[1] = loc[(size_t)i * 3 + 1]; p[2] = loc[(size_t)i * 3 + 2];
    }
    FcosBest best;
    fcos_best_init(best);
    for (int g0 = 0; g0 < n_gt; g0 += kFlGtChunk) {
        const int cnt = min(kFlGtChunk, n_gt - g0);
        __syncthreads();
        if (threadIdx.x < cnt) fcos_gt_prepare(gt + (size_t)(g0 + threadIdx.x) * gt_dim, gt_dim, sg[threadIdx.x]);
        __syncthreads();
        if (live)
            for (int g = 0; g < cnt; ++g) fcos_target_update(sg[g], p, P.radius_stride[lvl], P.size_lo[lvl], P.size_hi[lvl], best);
    }
    if (!live) return;
    labels[i] = (n_gt > 0 && best.area != kFcosInf) ? 1.f : 0.f;
    for (int k = 0; k < D; ++k) {
        float v = best.reg[k];
        if (k < 6 && P.norm[lvl]) v = v / P.norm_div[lvl];
        reg_targets[(size_t)i * D + k] = n_gt > 0 ? v : 0.f;
    }
}

__global__ void __launch_bounds__(kFlThreads) fcos_loss_kernel(const FcosLossDev P, double* __restrict__ partial) {
    double acc[kFlSums];
#pragma unroll
    for (int k = 0; k < kFlSums; ++k) acc[k] = 0.0;
    const long elements = (long)P.n_img * P.total;
    for (long e = (long)blockIdx.x * kFlThreads + threadIdx.x; e < elements; e += (long)gridDim.x * kFlThreads) fcos_loss_element(P, e, acc);
    __shared__ double red[kFlSums][kFlThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kFlSums; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < kFlSums) {
        double v = 0.0;
        for (int w = 0; w < kFlThreads / 32; ++w) v += red[threadIdx.x][w];
        partial[(size_t)blockIdx.x * kFlSums + threadIdx.x] = v;
    }
}

// Warp k adds the CTA partials of sum k: lane l takes CTAs l, l + 32, ... in order, then a fixed shuffle tree -- the same order on every run
// (a single thread walking the 592 partials took 26 us, half of the whole loss: profiles/r02_fcos_loss_ncu.md).
__global__ void __launch_bounds__(256) fcos_loss_final_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ sums) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = 0.0;
    if (k < kFlSums)
        for (int b = lane; b < blocks; b += 32) v += partial[(size_t)b * kFlSums + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sums[k] = v;                 // k = 0..7; 6 and 7 are zero
}

}  // namespace nrpn

using namespace nrpn;

extern "C" {
#pragma GCC visibility push(default)

int nrpn_fcos_targets(const nrpn_fcos_target_desc* d, const float* locations, const float* gt, int n_gt, int gt_dim, float* labels,
                      float* reg_targets, nrpn_stream_t stream) {
    if (!d || !locations || !labels || !reg_targets || d->n_levels < 1 || d->n_levels > NRPN_RPN_MAX_LEVELS) return NRPN_ERR_INVALID;
    if (n_gt < 0 || (n_gt > 0 && !gt) || (gt_dim != 6 && gt_dim != 7)) return NRPN_ERR_INVALID;
    FcosTgtDev P;
    memset(&P, 0, sizeof(P));
    P.n_levels = d->n_levels;
    long begin = 0;
    for (int l = 0; l < d->n_levels; ++l) {
        if (d->n_points[l] < 0 || d->stride[l] < 1) return NRPN_ERR_INVALID;
        P.begin[l] = (int)begin;
        begin += d->n_points[l];
        P.radius_stride[l] = d->center_sampling_radius > 0.f ? (float)((double)d->stride[l] * (double)d->center_sampling_radius) : 0.f;
        P.size_lo[l] = d->size_lo[l]; P.size_hi[l] = d->size_hi[l];
        P.norm[l] = d->norm_reg_targets ? 1 : 0;
        P.norm_div[l] = (float)d->stride[l];
    }
    if (begin > 0x7fffffffL) return NRPN_ERR_UNSUPPORTED;
    for (int l = d->n_levels; l <= NRPN_RPN_MAX_LEVELS; ++l) P.begin[l] = (int)begin;
    P.total = (int)begin;
    if (P.total == 0) return NRPN_OK;
    fcos_targets_kernel<<<ceil_div(P.total, kFlThreads), kFlThreads, 0, (cudaStream_t)stream>>>(P, locations, gt, n_gt, gt_dim, labels, reg_targets);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

size_t nrpn_fcos_loss_workspace_bytes(void) { return (size_t)kFlMaxBlocks * kFlSums * sizeof(double) + 256; }

int nrpn_fcos_loss(const nrpn_fcos_loss_desc* d, const float* labels, const float* reg_targets, const uint8_t* mask, float* centerness_targets,
                   double* sums, void* workspace, size_t workspace_bytes, nrpn_stream_t stream) {
    if (!d || !labels || !reg_targets || !sums || !workspace || d->n_levels < 1 || d->n_levels > NRPN_RPN_MAX_LEVELS || d->n_images < 1) return NRPN_ERR_INVALID;
    if (d->loss_type < 0 || d->loss_type > 3) return NRPN_ERR_INVALID;
    if (workspace_bytes < nrpn_fcos_loss_workspace_bytes()) return NRPN_ERR_WORKSPACE;
    FcosLossDev P;
    memset(&P, 0, sizeof(P));
    P.n_levels = d->n_levels; P.n_img = d->n_images; P.D = d->use_obb ? 8 : 6;
    P.loss_type = d->loss_type; P.use_obb = d->use_obb ? 1 : 0; P.add_l1 = d->additional_l1 ? 1 : 0;
    long begin = 0;
    int with_grad = 0, without_grad = 0;
    for (int l = 0; l < d->n_levels; ++l) {
        const nrpn_fcos_loss_level& L = d->level[l];
        if (!L.cls || !L.reg || !L.ctr || L.n_points < 0) return NRPN_ERR_INVALID;
        if (L.dcls && L.dreg && L.dctr) ++with_grad;
        else if (!L.dcls && !L.dreg && !L.dctr) ++without_grad;
        else return NRPN_ERR_INVALID;
        P.cls[l] = L.cls; P.reg[l] = L.reg; P.ctr[l] = L.ctr; P.dcls[l] = L.dcls; P.dreg[l] = L.dreg; P.dctr[l] = L.dctr;
        P.begin[l] = (int)begin;
        begin += L.n_points;
    }
    if (with_grad && without_grad) return NRPN_ERR_INVALID;
    if (begin * (long)d->n_images > 0x7fffffffL) return NRPN_ERR_UNSUPPORTED;
    for (int l = d->n_levels; l <= NRPN_RPN_MAX_LEVELS; ++l) P.begin[l] = (int)begin;
    P.total = (int)begin; P.want_grad = with_grad ? 1 : 0;
    P.labels = labels; P.rt = reg_targets; P.mask = mask; P.ct_out = centerness_targets;
    double* partial = reinterpret_cast<double*>(align_up((size_t)workspace, 256));
    cudaStream_t st = (cudaStream_t)stream;
    const long elements = (long)P.n_img * P.total;
    int blocks = (int)ceil_div(elements, (long)kFlThreads);
    blocks = blocks < 1 ? 1 : (blocks > kFlMaxBlocks ? kFlMaxBlocks : blocks);
    fcos_loss_kernel<<<blocks, kFlThreads, 0, st>>>(P, partial);
    NRPN_LAUNCH_CHECK();
    fcos_loss_final_kernel<<<1, 256, 0, st>>>(partial, blocks, sums);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

#pragma GCC visibility pop
}  // extern "C"
