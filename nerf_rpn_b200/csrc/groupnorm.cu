// GroupNorm(32, 256) + ReLU on channels-last bf16 activations, in place (FCOS towers: Conv3d -> GroupNorm -> ReLU,
// nerf_rpn/model/fcos/fcos.py:43-69).  Statistics are per (sample, pyramid level, group) over 8 channels x all voxels.
// Two kernels per call, HBM bound: (1) per-CTA partial sums (fp32 per thread, fp64 across threads) written to a scratch
// array, (2) every CTA first reduces the partials of its (sample, level) in a FIXED order (bit-reproducible), then applies
// y = relu((x - mean) * rstd * gamma + beta).  One thread owns one 16-byte chunk = the 8 channels of one group of one
// voxel, so a warp reads 512 contiguous bytes.  Up to 4 pyramid levels per launch.
#include "common.cuh"

namespace nrpn {

constexpr int kGnBlocks = 64;        // CTAs per (level, sample)
constexpr int kGnGroups = 32;

struct GnDev {
    int n_levels, n, relu;
    float eps;
    __nv_bfloat16* x[NRPN_CONV_MAX_LEVELS];
    int voxels[NRPN_CONV_MAX_LEVELS];
    const float* gamma; const float* beta;
    double* partial;                 // [level][sample][block][group][2]
};

__global__ void __launch_bounds__(256) gn_stats_kernel(GnDev P) {
    const int blk = blockIdx.x % kGnBlocks;
    const int ls = blockIdx.x / kGnBlocks;
    const int smp = ls % P.n, l = ls / P.n;
    const int V = P.voxels[l];
    const __nv_bfloat16* x = P.x[l] + (size_t)smp * V * 256;
    const int g = threadIdx.x & 31, r = threadIdx.x >> 5;       // 32 groups x 8 rows per CTA step
    float s = 0.f, q = 0.f;
    for (int v = blk * 8 + r; v < V; v += kGnBlocks * 8) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + (size_t)v * 256 + g * 8));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); s += f.x + f.y; q += f.x * f.x + f.y * f.y; }
    }
    __shared__ double ss[8][32], sq[8][32];
    ss[r][g] = (double)s; sq[r][g] = (double)q;
    __syncthreads();
    if (r == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += ss[i][g]; b += sq[i][g]; }
        double* o = P.partial + ((((size_t)l * P.n + smp) * kGnBlocks + blk) * kGnGroups + g) * 2;
        o[0] = a; o[1] = b;
    }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(GnDev P) {
    const int blk = blockIdx.x % kGnBlocks;
    const int ls = blockIdx.x / kGnBlocks;
    const int smp = ls % P.n, l = ls / P.n;
    const int V = P.voxels[l];
    __nv_bfloat16* x = P.x[l] + (size_t)smp * V * 256;
    const int g = threadIdx.x & 31, r = threadIdx.x >> 5;
    __shared__ float mean_s[32], rstd_s[32];
    if (threadIdx.x < 32) {
        const double* p = P.partial + (((size_t)l * P.n + smp) * kGnBlocks) * kGnGroups * 2;
        double a = 0.0, b = 0.0;
        for (int i = 0; i < kGnBlocks; ++i) { a += p[((size_t)i * kGnGroups + threadIdx.x) * 2]; b += p[((size_t)i * kGnGroups + threadIdx.x) * 2 + 1]; }
        const double cnt = (double)V * 8.0;
        const double m = a / cnt;
        double var = b / cnt - m * m;                      // biased variance, as torch.nn.GroupNorm
        if (var < 0.0) var = 0.0;
        mean_s[threadIdx.x] = (float)m;
        rstd_s[threadIdx.x] = (float)(1.0 / sqrt(var + (double)P.eps));
    }
    __syncthreads();
    const float mean = mean_s[g], rstd = rstd_s[g];
    float ga[8], be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ga[i] = __ldg(P.gamma + g * 8 + i) * rstd; be[i] = __ldg(P.beta + g * 8 + i) - mean * ga[i]; }
    for (int v = blk * 8 + r; v < V; v += kGnBlocks * 8) {
        uint4* ptr = reinterpret_cast<uint4*>(x + (size_t)v * 256 + g * 8);
        uint4 raw = *ptr;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            f.x = f.x * ga[2 * i] + be[2 * i]; f.y = f.y * ga[2 * i + 1] + be[2 * i + 1];
            if (P.relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
            h[i] = __floats2bfloat162_rn(f.x, f.y);
        }
        *ptr = raw;
    }
}

}  // namespace nrpn

using namespace nrpn;

extern "C" {
#pragma GCC visibility push(default)

size_t nrpn_groupnorm_workspace_bytes(int n_levels, int n) {
    if (n_levels < 1 || n_levels > NRPN_CONV_MAX_LEVELS || n < 1) return 0;
    return (size_t)n_levels * n * kGnBlocks * kGnGroups * 2 * sizeof(double);
}

int nrpn_groupnorm_relu(const nrpn_gn_level* levels, int n_levels, int n, int c, int groups, const float* gamma,
                        const float* beta, float eps, int relu, void* workspace, size_t workspace_bytes, nrpn_stream_t stream) {
    if (!levels || !gamma || !beta || !workspace || n_levels < 1 || n_levels > NRPN_CONV_MAX_LEVELS || n < 1) return NRPN_ERR_INVALID;
    if (c != 256 || groups != 32) return NRPN_ERR_UNSUPPORTED;       // 8 channels per group = one 16-byte chunk per thread
    if (workspace_bytes < nrpn_groupnorm_workspace_bytes(n_levels, n)) return NRPN_ERR_WORKSPACE;
    GnDev P;
    P.n_levels = n_levels; P.n = n; P.relu = relu; P.eps = eps; P.gamma = gamma; P.beta = beta;
    P.partial = reinterpret_cast<double*>(workspace);
    for (int l = 0; l < n_levels; ++l) {
        if (!levels[l].x || levels[l].voxels < 1) return NRPN_ERR_INVALID;
        P.x[l] = reinterpret_cast<__nv_bfloat16*>(levels[l].x); P.voxels[l] = levels[l].voxels;
    }
    const int grid = n_levels * n * kGnBlocks;
    gn_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    NRPN_LAUNCH_CHECK();
    gn_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

#pragma GCC visibility pop
}  // extern "C"
