// Internal (non-ABI) entry points shared between nms.cu and rpn_post.cu.
#pragma once
#include "common.cuh"

namespace nrpn {

// In-place ascending sort of n_pad (power of two) 64-bit keys.
int bitonic_sort_u64(unsigned long long* keys, int n_pad, cudaStream_t st);

size_t nms_workspace_bytes(int n);

// Per-group greedy NMS; entries whose group equals ignore_group are neither kept nor suppress anything.
// max_group: upper bound on the number of boxes sharing one group id (0 = unknown); it only selects between the full
// bit-matrix and the chunked evaluation order, never the result.
int nms_run(const float* boxes, int box_dim, const float* scores, const int32_t* group, int n, float thr, int ignore_group,
            int64_t* keep, int32_t* n_keep, void* ws, size_t ws_bytes, cudaStream_t st, int max_group = 0);

}  // namespace nrpn
