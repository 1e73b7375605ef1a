// Greedy proposal <-> ground-truth matching of the recall metric (eval.py:14-81, evaluate_box_proposals_recall), on device.
//
// The reference builds overlaps = box_iou_3d(proposals, gt) (P x G) and loops min(P, G) times on the host: take the
// per-GT maximum over proposals, pick the best-covered GT, record its IoU, overwrite that proposal's row and that GT's
// column with -1.  One CTA does the same loop here on the IoU matrix produced by nrpn_iou3d_matrix; nothing returns to
// the host until the recall is read.  Ties resolve like torch.max on CPU: first (lowest) index wins, GT first, then
// proposal.
#include "common.cuh"

namespace nrpn {

constexpr int kMatchThreads = 256;
constexpr int kMatchMaxGt = 4096;            // 12 bits of the key
constexpr int kMatchMaxProposals = 262144;   // one bit each in (dynamic) shared memory: 32 KB; the key has 20 bits for them

// key: larger is better. value (fp32, -1 for used) ordered, then lower gt, then lower proposal.
__device__ __forceinline__ unsigned long long match_key(float v, int gt, int prop) {
    return ((unsigned long long)float_to_ordered(v) << 32) | ((unsigned long long)(0xFFFu - (unsigned)gt) << 20) |
           (unsigned long long)(0xFFFFFu - (unsigned)prop);
}

__global__ void __launch_bounds__(kMatchThreads) recall_match_kernel(const float* __restrict__ overlaps, int P, int G,
                                                                     float* __restrict__ gt_overlaps) {
    __shared__ unsigned long long red[kMatchThreads / 32];
    extern __shared__ unsigned row_used[];                  // ceil(P / 32) words (host check: P <= kMatchMaxProposals)
    __shared__ unsigned col_used[kMatchMaxGt / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < (P + 31) / 32; i += kMatchThreads) row_used[i] = 0u;
    for (int i = tid; i < (G + 31) / 32; i += kMatchThreads) col_used[i] = 0u;
    __syncthreads();
    const int iters = P < G ? P : G;
    const size_t total = (size_t)P * G;
    for (int j = 0; j < iters; ++j) {
        unsigned long long best = 0ull;
        for (size_t e = tid; e < total; e += kMatchThreads) {
            const int p = (int)(e / G), g = (int)(e - (size_t)p * G);
            const bool used = ((row_used[p >> 5] >> (p & 31)) & 1u) || ((col_used[g >> 5] >> (g & 31)) & 1u);
            const float v = used ? -1.0f : overlaps[e];
            const unsigned long long k = match_key(v, g, p);
            best = k > best ? k : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) red[wid] = best;
        __syncthreads();
        if (tid == 0) {
            unsigned long long b = red[0];
            for (int w = 1; w < kMatchThreads / 32; ++w) b = red[w] > b ? red[w] : b;
            const int g = (int)(0xFFFu - (unsigned)((b >> 20) & 0xFFFu));
            const int p = (int)(0xFFFFFu - (unsigned)(b & 0xFFFFFu));
            gt_overlaps[j] = overlaps[(size_t)p * G + g];
            row_used[p >> 5] |= 1u << (p & 31);
            col_used[g >> 5] |= 1u << (g & 31);
        }
        __syncthreads();
    }
}

// Per-row maximum and its first index (torch.max(dim=1) on CPU), one warp per row: the detection -> best ground truth step of
// the VOC-style AP (eval.py:355-358), for all detections of a scene at once instead of one box_iou_3d call per detection.
__global__ void __launch_bounds__(256) rowmax_kernel(const float* __restrict__ m, int rows, int cols, float* __restrict__ maxv,
                                                     int* __restrict__ argmax) {
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    unsigned long long best = 0ull;
    for (int c = lane; c < cols; c += 32)
        best = max(best, ((unsigned long long)float_to_ordered(m[(size_t)row * cols + c]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
    if (lane == 0) {
        const int c = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
        argmax[row] = c;
        maxv[row] = m[(size_t)row * cols + c];
    }
}

}  // namespace nrpn

using namespace nrpn;

extern "C" {
#pragma GCC visibility push(default)

int nrpn_recall_match(const float* overlaps, int n_proposals, int n_gt, float* gt_overlaps, nrpn_stream_t stream) {
    if (n_proposals < 0 || n_gt < 0) return NRPN_ERR_INVALID;
    if (n_proposals == 0 || n_gt == 0) return NRPN_OK;
    if (!overlaps || !gt_overlaps) return NRPN_ERR_INVALID;
    if (n_gt > kMatchMaxGt || n_proposals > kMatchMaxProposals) return NRPN_ERR_UNSUPPORTED;
    recall_match_kernel<<<1, kMatchThreads, (size_t)ceil_div(n_proposals, 32) * sizeof(unsigned), (cudaStream_t)stream>>>(overlaps, n_proposals, n_gt, gt_overlaps);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

int nrpn_rowmax_f32(const float* m, int rows, int cols, float* maxv, int32_t* argmax, nrpn_stream_t stream) {
    if (rows < 0 || cols < 1) return NRPN_ERR_INVALID;
    if (rows == 0) return NRPN_OK;
    if (!m || !maxv || !argmax) return NRPN_ERR_INVALID;
    rowmax_kernel<<<(unsigned)ceil_div((long)rows * 32, 256L), 256, 0, (cudaStream_t)stream>>>(m, rows, cols, maxv, argmax);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

#pragma GCC visibility pop
}  // extern "C"
