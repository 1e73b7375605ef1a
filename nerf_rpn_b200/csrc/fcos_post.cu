// FCOS (anchor-free) inference post-processing for one scene, fully on device:
//   head output transform (Scale, ReLU, x stride)          fcos/fcos.py:116-126
//   per level: sigmoid, candidates, x centerness, top-k      fcos/inference.py:48-104  (radix select, topk.cuh)
//   decode (AABB or midpoint-offset OBB), clip, min-size     fcos/inference.py:106-124, fcos/utils.py:12-61
//   one NMS across all levels, cap by the k-th best score     fcos/inference.py:164-195
// Same conventions as rpn_post.cu / oracle/fcos_post.py (ties -> lower index first, fp32 ops separately rounded).
#include "rpn_decode.cuh"
#include "nms_internal.cuh"
#include "topk.cuh"

namespace nrpn {

constexpr int kFcosIgnore = 255;

struct FcosLevelDev {
    const float* cls; const float* reg; int ldc, ldr; int gx, gy, gz; int stride; float scale;
    int count, k, cand_off;
};
struct FcosDev {
    int n_levels, obb, code, padded;
    FcosLevelDev lv[NRPN_RPN_MAX_LEVELS];
    float valid[3]; float pre_thresh, min_size;
    int total_cand;
};

__device__ __forceinline__ void fcos_loc(const FcosLevelDev& L, int i, float& x, float& y, float& z) {
    const int iz = i % L.gz; const int t = i / L.gz; const int iy = t % L.gy; const int ix = t / L.gy;
    const float half = (float)(L.stride / 2);
    x = __fadd_rn((float)(ix * L.stride), half); y = __fadd_rn((float)(iy * L.stride), half); z = __fadd_rn((float)(iz * L.stride), half);
}

struct FcosSrc {
    FcosDev P;
    __device__ __forceinline__ int levels() const { return P.n_levels; }
    __device__ __forceinline__ int count(int l) const { return P.lv[l].count; }
    __device__ __forceinline__ int k(int l) const { return P.lv[l].k; }
    __device__ __forceinline__ int cand_off(int l) const { return P.lv[l].cand_off; }
    // candidate iff sigmoid(cls) > pre_nms_thresh (padded locations get -1e5); key = cls*centerness score
    __device__ __forceinline__ bool key(int l, int i, unsigned long long& key) const {
        const FcosLevelDev& L = P.lv[l];
        float c = sigmoid_ref(L.cls[(size_t)i * L.ldc]);
        if (P.padded) {
            float x, y, z; fcos_loc(L, i, x, y, z);
            if (!(x < P.valid[0] && y < P.valid[1] && z < P.valid[2])) c = -1e5f;
        }
        if (!(c > P.pre_thresh)) return false;
        const float ctr = sigmoid_ref(L.reg[(size_t)i * L.ldr + P.code]);
        key = make_key56(float_to_ordered(__fmul_rn(c, ctr)), i);
        return true;
    }
};

__device__ __forceinline__ float norm2(float x, float y) { return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))); }

// decode_fcos_obb (fcos/utils.py:12-61): location + (x0,y0,z0,x1,y1,z1 distances, alpha, beta) -> (cx,cy,cz,w,l,h,theta)
__device__ __forceinline__ void decode_fcos_obb(float lx, float ly, float lz, const float* __restrict__ r, float* __restrict__ out) {
    const float x0 = __fsub_rn(lx, r[0]), y0 = __fsub_rn(ly, r[1]), z0 = __fsub_rn(lz, r[2]);
    const float x1 = __fadd_rn(lx, r[3]), y1 = __fadd_rn(ly, r[4]), z1 = __fadd_rn(lz, r[5]);
    float vx = __fadd_rn(__fdiv_rn(__fadd_rn(x1, x0), 2.0f), __fmul_rn(r[6], __fsub_rn(x1, x0)));
    float vy = __fadd_rn(__fdiv_rn(__fadd_rn(y1, y0), 2.0f), __fmul_rn(r[7], __fsub_rn(y1, y0)));
    vx = fminf(fmaxf(vx, x0), x1); vy = fminf(fmaxf(vy, y0), y1);
    const float cx = __fdiv_rn(__fadd_rn(x0, x1), 2.0f), cy = __fdiv_rn(__fadd_rn(y0, y1), 2.0f), cz = __fdiv_rn(__fadd_rn(z0, z1), 2.0f);
    float v0x = __fsub_rn(vx, cx), v0y = __fsub_rn(y1, cy), v1x = __fsub_rn(x1, cx), v1y = __fsub_rn(vy, cy);
    const float d0 = norm2(v0x, v0y), d1 = norm2(v1x, v1y);
    const float dmax = fmaxf(d0, d1);
    const float e0 = __fadd_rn(d0, 1e-7f), e1 = __fadd_rn(d1, 1e-7f);
    v0x = __fadd_rn(__fmul_rn(__fdiv_rn(v0x, e0), dmax), cx); v0y = __fadd_rn(__fmul_rn(__fdiv_rn(v0y, e0), dmax), cy);
    v1x = __fadd_rn(__fmul_rn(__fdiv_rn(v1x, e1), dmax), cx); v1y = __fadd_rn(__fmul_rn(__fdiv_rn(v1y, e1), dmax), cy);
    const float ln = norm2(__fsub_rn(v0x, v1x), __fsub_rn(v0y, v1y));
    float mx = __fsub_rn(__fdiv_rn(__fadd_rn(v0x, v1x), 2.0f), cx);
    const float my = __fsub_rn(__fdiv_rn(__fadd_rn(v0y, v1y), 2.0f), cy);
    const float w = __fmul_rn(norm2(mx, my), 2.0f);
    if (mx == 0.0f && my == 0.0f) mx = 1e-7f;
    const float theta = (float)atan2((double)my, (double)mx);
    out[0] = cx; out[1] = cy; out[2] = cz; out[3] = w; out[4] = ln; out[5] = __fsub_rn(z1, z0); out[6] = theta;
}

// slot p of the sorted candidate list -> box, score, NMS group (0 = live, 255 = dropped)
__global__ void fcos_decode_kernel(FcosDev P, const unsigned long long* __restrict__ cand, float* __restrict__ nbox,
                                   float* __restrict__ fscore, int* __restrict__ flevel, int* __restrict__ group) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.total_cand) return;
    const int dim = P.obb ? 7 : 6;
    const unsigned long long ck = cand[p];
    float* ob = nbox + (size_t)p * dim;
    if (ck == ~0ull) {                                   // fewer candidates than slots
        group[p] = kFcosIgnore; fscore[p] = 0.f; flevel[p] = 0;
        for (int k = 0; k < dim; ++k) ob[k] = 0.f;
        return;
    }
    const int l = (int)(ck >> 56);
    const unsigned long long key = (~ck) & 0x00FFFFFFFFFFFFFFull;
    const int i = key56_index(key);
    const FcosLevelDev& L = P.lv[l];
    float lx, ly, lz; fcos_loc(L, i, lx, ly, lz);
    const float* raw = L.reg + (size_t)i * L.ldr;
    float r[8];
    const float fs = (float)L.stride;
    for (int c = 0; c < P.code; ++c) {
        float v = __fmul_rn(raw[c], L.scale);            // Scale (fcos.py:116)
        if (c < 6) { v = fmaxf(v, 0.0f); v = __fmul_rn(v, fs); }   // ReLU on the six distances, x stride at eval (:118-124)
        r[c] = v;
    }
    // recompute the score exactly as the top-k key did
    float c0 = sigmoid_ref(L.cls[(size_t)i * L.ldc]);
    const float ctr = sigmoid_ref(raw[P.code]);
    const float score = __fsqrt_rn(__fmul_rn(c0, ctr));
    bool keep;
    if (!P.obb) {
        float b[6] = {__fsub_rn(lx, r[0]), __fsub_rn(ly, r[1]), __fsub_rn(lz, r[2]), __fadd_rn(lx, r[3]), __fadd_rn(ly, r[4]), __fadd_rn(lz, r[5])};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float lo = b[k], hi = b[3 + k];
            if (lo < 0.f) lo = 0.f; if (lo > P.valid[k]) lo = P.valid[k];
            if (hi < 0.f) hi = 0.f; if (hi > P.valid[k]) hi = P.valid[k];
            b[k] = lo; b[3 + k] = hi;
        }
        keep = (__fsub_rn(b[3], b[0]) >= P.min_size) && (__fsub_rn(b[4], b[1]) >= P.min_size) && (__fsub_rn(b[5], b[2]) >= P.min_size);
#pragma unroll
        for (int k = 0; k < 6; ++k) ob[k] = b[k];
    } else {
        float b[7];
        decode_fcos_obb(lx, ly, lz, r, b);
        keep = (b[3] >= P.min_size) && (b[4] >= P.min_size) && (b[5] >= P.min_size);
#pragma unroll
        for (int k = 0; k < 7; ++k) ob[k] = b[k];
    }
    fscore[p] = score; flevel[p] = l; group[p] = keep ? 0 : kFcosIgnore;
}

// keep[] is score-descending. If more than post_top_n survive, everything scoring >= the post_top_n-th best stays
// (torch.kthvalue cut, ties included): a prefix of keep[].
__global__ void fcos_emit_kernel(const int64_t* __restrict__ keep, const int32_t* __restrict__ n_keep, int post_top_n, int cap, int dim,
                                 const float* __restrict__ nbox, const float* __restrict__ fscore, const int* __restrict__ flevel,
                                 float* __restrict__ boxes, float* __restrict__ scores, int32_t* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cap) return;
    const int nk = *n_keep;
    float thr = -INFINITY;
    if (post_top_n > 0 && nk > post_top_n) thr = fscore[keep[post_top_n - 1]];
    const bool live = p < nk && fscore[keep[p < nk ? p : 0]] >= thr;
    float* ob = boxes + (size_t)p * (dim + 1);
    if (live) {
        const int src = (int)keep[p];
        ob[0] = (float)flevel[src];
        for (int k = 0; k < dim; ++k) ob[1 + k] = nbox[(size_t)src * dim + k];
        scores[p] = fscore[src];
        const bool next_live = (p + 1 < nk) && fscore[keep[p + 1]] >= thr;
        if (!next_live) *count = p + 1;
    } else {
        for (int k = 0; k <= dim; ++k) ob[k] = 0.f;
        scores[p] = 0.f;
        if (p == 0) *count = 0;
    }
}

static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct FcosWs {
    unsigned* hist; SelState* st; unsigned* counters; unsigned long long* cand;
    float* nbox; float* fscore; int* flevel; int* group; int64_t* keep; int32_t* n_keep; void* nms_ws; size_t nms_bytes; size_t total;
};

static FcosWs fcos_layout(void* base, int M) {
    FcosWs w; size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    char* b = (char*)base;
    const int cpad = next_pow2(M < 2 ? 2 : M);
    w.hist = (unsigned*)(b + take((size_t)kTopkMaxLevels * kBins * 4));
    w.st = (SelState*)(b + take(kTopkMaxLevels * sizeof(SelState)));
    w.counters = (unsigned*)(b + take(kTopkMaxLevels * 4));
    w.cand = (unsigned long long*)(b + take((size_t)cpad * 8));
    w.nbox = (float*)(b + take((size_t)M * 7 * 4));
    w.fscore = (float*)(b + take((size_t)M * 4));
    w.flevel = (int*)(b + take((size_t)M * 4));
    w.group = (int*)(b + take((size_t)M * 4));
    w.keep = (int64_t*)(b + take((size_t)M * 8));
    w.n_keep = (int32_t*)(b + take(256));
    w.nms_bytes = nms_workspace_bytes(M);
    w.nms_ws = (void*)(b + take(w.nms_bytes));
    w.total = off;
    return w;
}

static int fcos_build(const nrpn_fcos_desc* d, FcosDev& P) {
    if (!d || d->n_levels < 1 || d->n_levels > NRPN_RPN_MAX_LEVELS || d->pre_nms_top_n < 1) return NRPN_ERR_INVALID;
    P.n_levels = d->n_levels; P.obb = d->use_obb ? 1 : 0; P.code = d->use_obb ? 8 : 6; P.padded = d->padded ? 1 : 0;
    int off = 0;
    for (int l = 0; l < d->n_levels; ++l) {
        const nrpn_fcos_level& s = d->level[l];
        FcosLevelDev& L = P.lv[l];
        if (s.gx < 1 || s.gy < 1 || s.gz < 1 || s.stride < 1 || s.ld_cls < 1 || s.ld_reg < P.code + 1) return NRPN_ERR_INVALID;
        const long cnt = (long)s.gx * s.gy * s.gz;
        if (cnt >= (1L << kIdxBits)) return NRPN_ERR_UNSUPPORTED;
        L.cls = s.cls; L.reg = s.reg; L.ldc = s.ld_cls; L.ldr = s.ld_reg; L.gx = s.gx; L.gy = s.gy; L.gz = s.gz;
        L.stride = s.stride; L.scale = s.scale; L.count = (int)cnt;
        L.k = d->pre_nms_top_n < L.count ? d->pre_nms_top_n : L.count; L.cand_off = off;
        off += L.k;
    }
    P.total_cand = off;
    for (int k = 0; k < 3; ++k) P.valid[k] = (float)d->grid_size[k];
    P.pre_thresh = d->pre_nms_thresh; P.min_size = d->min_size;
    return NRPN_OK;
}

}  // namespace nrpn

using namespace nrpn;

extern "C" {
#pragma GCC visibility push(default)

int nrpn_fcos_max_proposals(const nrpn_fcos_desc* desc) {
    FcosDev P;
    if (fcos_build(desc, P) != NRPN_OK) return 0;
    return P.total_cand;
}

size_t nrpn_fcos_workspace_bytes(const nrpn_fcos_desc* desc) {
    FcosDev P;
    if (fcos_build(desc, P) != NRPN_OK) return 0;
    return fcos_layout(nullptr, P.total_cand).total + 256;
}

int nrpn_fcos_proposals(const nrpn_fcos_desc* desc, float* boxes, float* scores, int32_t* count, void* workspace,
                        size_t workspace_bytes, nrpn_stream_t stream) {
    FcosDev P;
    int rc = fcos_build(desc, P);
    if (rc) return rc;
    if (!boxes || !scores || !count || !workspace) return NRPN_ERR_INVALID;
    for (int l = 0; l < P.n_levels; ++l) if (!P.lv[l].cls || !P.lv[l].reg) return NRPN_ERR_INVALID;
    const int M = P.total_cand;
    if (M > nrpn_nms_max_boxes()) return NRPN_ERR_UNSUPPORTED;
    void* base = (void*)align_up((size_t)workspace, 256);
    FcosWs w = fcos_layout(base, M);
    if (workspace_bytes < w.total + ((char*)base - (char*)workspace)) return NRPN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = P.n_levels;
    const int cpad = next_pow2(M < 2 ? 2 : M);
    NRPN_CUDA_TRY(cudaMemsetAsync(w.hist, 0, (size_t)kTopkMaxLevels * kBins * 4, st));
    NRPN_CUDA_TRY(cudaMemsetAsync(w.counters, 0, kTopkMaxLevels * 4, st));
    FcosSrc src{P};
    topk_init_kernel<FcosSrc><<<1, 32, 0, st>>>(src, w.st);
    NRPN_LAUNCH_CHECK();
    int max_count = 0;
    for (int l = 0; l < L; ++l) max_count = P.lv[l].count > max_count ? P.lv[l].count : max_count;
    int bx = ceil_div(max_count, 256 * 4);
    if (bx > 2 * num_sms()) bx = 2 * num_sms();
    if (bx < 1) bx = 1;
    for (int pass = 0; pass < kPasses; ++pass) {
        topk_hist_kernel<FcosSrc><<<dim3(bx, L), 256, 0, st>>>(src, pass, w.st, w.hist);
        NRPN_LAUNCH_CHECK();
        topk_select_kernel<<<L, 1024, 0, st>>>(pass, w.st, w.hist);
        NRPN_LAUNCH_CHECK();
    }
    fill_u64_kernel<<<ceil_div(cpad, 256), 256, 0, st>>>(w.cand, cpad, ~0ull);
    NRPN_LAUNCH_CHECK();
    topk_collect_kernel<FcosSrc><<<dim3(bx, L), 256, 0, st>>>(src, w.st, w.counters, w.cand);
    NRPN_LAUNCH_CHECK();
    rc = bitonic_sort_u64(w.cand, cpad, st);
    if (rc) return rc;
    fcos_decode_kernel<<<ceil_div(M, 128), 128, 0, st>>>(P, w.cand, w.nbox, w.fscore, w.flevel, w.group);
    NRPN_LAUNCH_CHECK();
    const int dim = P.obb ? 7 : 6;
    rc = nms_run(w.nbox, dim, w.fscore, w.group, M, desc->nms_thresh, kFcosIgnore, w.keep, w.n_keep, w.nms_ws, w.nms_bytes, st);
    if (rc) return rc;
    fcos_emit_kernel<<<ceil_div(M, 256), 256, 0, st>>>(w.keep, w.n_keep, desc->post_nms_top_n, M, dim, w.nbox, w.fscore, w.flevel,
                                                      boxes, scores, count);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

#pragma GCC visibility pop
}  // extern "C"
