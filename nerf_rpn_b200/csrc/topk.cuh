// Per-level top-k by MSB-first radix select on unique 56-bit keys (12-bit digits, 5 passes), shared by the anchor RPN
// (rpn_post.cu: key = ordered logit) and the FCOS post-processing (fcos_post.cu: key = ordered cls*centerness score).
// A source type `Src` (passed by value) provides:
//   int levels() ; int count(l) ; int k(l) ; int cand_off(l)
//   bool key(l, i, unsigned long long& key)      -> false when element i is not a candidate
// Keys must be unique per level (the index is folded into the low 24 bits), larger = better.
#pragma once
#include "common.cuh"

namespace nrpn {

constexpr int kDigitBits = 12;
constexpr int kBins = 1 << kDigitBits;
constexpr int kPasses = 5;                 // 5 x 12 = 60 >= 56 key bits
constexpr int kIdxBits = 24;
constexpr unsigned kIdxMask = (1u << kIdxBits) - 1u;
constexpr int kTopkMaxLevels = 4;

__device__ __forceinline__ unsigned long long make_key56(unsigned ordered_value, int idx) {
    return ((unsigned long long)ordered_value << kIdxBits) | (unsigned long long)(kIdxMask - (unsigned)idx);
}
__device__ __forceinline__ int key56_index(unsigned long long key) { return (int)(kIdxMask - (unsigned)(key & kIdxMask)); }

// state per level: prefix (selected high digits), remaining k, number selected in total
struct SelState { unsigned long long prefix; int remaining; int selected; };

template <class Src>
__global__ void topk_hist_kernel(Src src, int pass, const SelState* __restrict__ st, unsigned* __restrict__ hist) {
    __shared__ unsigned sh[kBins];
    const int l = blockIdx.y;
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const int shift = (kPasses - 1 - pass) * kDigitBits;
    const unsigned long long prefix = st[l].prefix;
    const int count = src.count(l);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        unsigned long long key;
        if (!src.key(l, i, key)) continue;
        if (pass == 0 || (key >> (shift + kDigitBits)) == prefix)
            atomicAdd(&sh[(unsigned)(key >> shift) & (kBins - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) if (sh[i]) atomicAdd(&hist[l * kBins + i], sh[i]);
}

// one CTA (1024 threads) per level: find the digit that contains the k-th largest key, clear the histogram.
// In pass 0 the number of candidates is known (histogram total): k is clipped to it (FCOS: data-dependent candidate count).
static __global__ void __launch_bounds__(1024) topk_select_kernel(int pass, SelState* __restrict__ st, unsigned* __restrict__ hist) {
    __shared__ unsigned part[1024];
    __shared__ int sel_bin, sel_above;
    const int l = blockIdx.x;
    unsigned* h = hist + l * kBins;
    const int t = threadIdx.x;
    unsigned c[4]; unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[j] = h[kBins - 1 - (4 * t + j)]; s += c[j]; }     // thread t owns 4 bins, counted from the top
    part[t] = s;
    if (t == 0) { sel_bin = 0; sel_above = 0; }
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {                                        // inclusive scan (Hillis-Steele)
        unsigned v = (t >= off) ? part[t - off] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned rem = (unsigned)st[l].remaining;
    const unsigned total = part[1023];
    if (pass == 0 && rem > total) rem = total;
    const unsigned before = part[t] - s;
    if (rem > 0 && before < rem && part[t] >= rem) {
        unsigned acc = before;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (acc < rem && acc + c[j] >= rem) { sel_bin = kBins - 1 - (4 * t + j); sel_above = (int)acc; }
            acc += c[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) h[kBins - 1 - (4 * t + j)] = 0u;
    if (t == 0) {
        if (pass == 0) st[l].selected = (int)rem;
        if (rem == 0) {                     // nothing to select: a prefix no key can reach
            st[l].prefix = ~0ull; st[l].remaining = 0;
        } else {
            st[l].prefix = (st[l].prefix << kDigitBits) | (unsigned long long)sel_bin;
            st[l].remaining = (int)rem - sel_above;
        }
    }
}

// after the last pass st[l].prefix is the key of the k-th largest element: gather everything >= it.
// cand entry = (level << 56) | (~key & mask56): an ascending sort yields level-major, best-first order.
template <class Src>
__global__ void topk_collect_kernel(Src src, const SelState* __restrict__ st, unsigned* __restrict__ counters,
                                    unsigned long long* __restrict__ cand) {
    const int l = blockIdx.y;
    const unsigned long long thr = st[l].prefix;
    const int count = src.count(l), k = src.k(l), off = src.cand_off(l);
    if (st[l].selected == 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        unsigned long long key;
        if (!src.key(l, i, key)) continue;
        if (key >= thr) {
            const unsigned slot = atomicAdd(&counters[l], 1u);
            if (slot < (unsigned)k) cand[off + slot] = ((unsigned long long)l << 56) | (~key & 0x00FFFFFFFFFFFFFFull);
        }
    }
}

template <class Src>
__global__ void topk_init_kernel(Src src, SelState* __restrict__ st) {
    const int l = threadIdx.x;
    if (l < kTopkMaxLevels) { st[l].prefix = 0ull; st[l].remaining = l < src.levels() ? src.k(l) : 0; st[l].selected = 0; }
}

__global__ void fill_u64_kernel(unsigned long long* p, int n, unsigned long long v);

}  // namespace nrpn
