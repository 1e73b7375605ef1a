// RPN post-processing for one scene, fully on device (no host round trips):
//   per-level top-k on raw logits  -> rpn.py:292-301      (radix select on unique 56-bit keys, 12-bit digits)
//   anchors + decode of the selected candidates only       -> anchor.py:98-122, coder/*.py
//     (the reference decodes all ~2.4 M anchors before the top-k, rpn.py:508)
//   sigmoid, clip / centre filter, min-size, score filter   -> rpn.py:342-357, utils.py:268-367
//   per-level greedy NMS, score-ordered merge, truncation    -> rpn.py:360-364, utils.py:233-265
// Candidate order is level-major, then (logit desc, flat anchor index asc); this is the tie convention shared
// with oracle/rpn_post.py (torch.topk leaves ties unspecified).
#include "rpn_decode.cuh"
#include "nms_internal.cuh"
#include "topk.cuh"

namespace nrpn {

constexpr int kIgnoreGroup = 255;

struct LevelDev {
    const float* pred; int ld; int gx, gy, gz; int sx, sy, sz; int lim_x, lim_y, lim_z;
    int count;      // anchors in this level = gx*gy*gz*A
    int k;          // min(pre_nms_top_n, count)
    int cand_off;   // offset of this level's candidates
};
struct RpnDev {
    int n_levels; int A; int rotated; int code;
    LevelDev lv[NRPN_RPN_MAX_LEVELS];
    float cell[NRPN_RPN_MAX_LEVELS][16][6];
    float mesh[3]; float min_size; float score_thresh;
    int total_cand;
};

__device__ __forceinline__ float level_logit(const LevelDev& L, int A, int idx) {
    const int vox = idx / A, a = idx - vox * A;
    const int iz = vox % L.gz; const int t = vox / L.gz; const int iy = t % L.gy; const int ix = t / L.gy;
    float v = L.pred[(size_t)vox * L.ld + a];
    if (ix >= L.lim_x || iy >= L.lim_y || iz >= L.lim_z) v = -INFINITY;     // padded voxels (rpn.py:321-322)
    return v;
}

// top-k source: every anchor is a candidate, key = (ordered raw logit, lower flat index wins ties)
struct RpnSrc {
    RpnDev P;
    __device__ __forceinline__ int levels() const { return P.n_levels; }
    __device__ __forceinline__ int count(int l) const { return P.lv[l].count; }
    __device__ __forceinline__ int k(int l) const { return P.lv[l].k; }
    __device__ __forceinline__ int cand_off(int l) const { return P.lv[l].cand_off; }
    __device__ __forceinline__ bool key(int l, int i, unsigned long long& key) const {
        key = make_key56(float_to_ordered(level_logit(P.lv[l], P.A, i)), i);
        return true;
    }
};

__global__ void fill_u64_kernel(unsigned long long* p, int n, unsigned long long v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// candidate p (sorted): decode, score, clip, flags
__global__ void rpn_decode_kernel(RpnDev P, const unsigned long long* __restrict__ cand, float* __restrict__ cbox,
                                  float* __restrict__ cscore, int* __restrict__ clevel, int* __restrict__ cflag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.total_cand) return;
    const unsigned long long ck = cand[p];
    const int l = (int)(ck >> 56);
    const unsigned long long key = (~ck) & 0x00FFFFFFFFFFFFFFull;
    const int idx = key56_index(key);
    const LevelDev& L = P.lv[l];
    const int A = P.A;
    const int vox = idx / A, a = idx - vox * A;
    const int iz = vox % L.gz; const int t = vox / L.gz; const int iy = t % L.gy; const int ix = t / L.gy;
    const float* row = L.pred + (size_t)vox * L.ld;
    const float logit = level_logit(L, A, idx);
    float an[6];
    const float sx = (float)(ix * L.sx), sy = (float)(iy * L.sy), sz = (float)(iz * L.sz);
    an[0] = __fadd_rn(sx, P.cell[l][a][0]); an[1] = __fadd_rn(sy, P.cell[l][a][1]); an[2] = __fadd_rn(sz, P.cell[l][a][2]);
    an[3] = __fadd_rn(sx, P.cell[l][a][3]); an[4] = __fadd_rn(sy, P.cell[l][a][4]); an[5] = __fadd_rn(sz, P.cell[l][a][5]);
    float d[8];
    for (int c = 0; c < P.code; ++c) d[c] = row[A + a * P.code + c];
    const float score = sigmoid_ref(logit);
    float* ob = cbox + (size_t)p * 8;
    int flag;
    if (!P.rotated) {
        float b[6];
        decode_aabb(an, d, b);
#pragma unroll
        for (int k = 0; k < 3; ++k) {       // clip_boxes_to_mesh, utils.py:344-357 (clamp keeps NaN)
            float lo = b[k], hi = b[3 + k];
            if (lo < 0.f) lo = 0.f; if (lo > P.mesh[k]) lo = P.mesh[k];
            if (hi < 0.f) hi = 0.f; if (hi > P.mesh[k]) hi = P.mesh[k];
            b[k] = lo; b[3 + k] = hi;
        }
        const bool big = (__fsub_rn(b[3], b[0]) >= P.min_size) && (__fsub_rn(b[4], b[1]) >= P.min_size) &&
                         (__fsub_rn(b[5], b[2]) >= P.min_size);
        flag = (big && score >= P.score_thresh) ? 1 : 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) ob[k] = b[k];
    } else {
        float b[7];
        decode_obb(an, d, b);
        const bool inside = (b[0] >= 0.f) && (b[0] <= P.mesh[0]) && (b[1] >= 0.f) && (b[1] <= P.mesh[1]) &&
                            (b[2] >= 0.f) && (b[2] <= P.mesh[2]);
        flag = inside ? 1 : 0;               // compaction + remaining filters happen in rpn_compact_kernel
#pragma unroll
        for (int k = 0; k < 7; ++k) ob[k] = b[k];
    }
    cscore[p] = score; clevel[p] = l; cflag[p] = flag;
}

// AABB: group = valid ? level : ignore.  One thread per candidate.
__global__ void rpn_group_aabb_kernel(int n, const int* __restrict__ clevel, const int* __restrict__ cflag,
                                      const float* __restrict__ cbox, float* __restrict__ fbox, int* __restrict__ group) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    group[p] = cflag[p] ? clevel[p] : kIgnoreGroup;
#pragma unroll
    for (int k = 0; k < 6; ++k) fbox[(size_t)p * 6 + k] = cbox[(size_t)p * 8 + k];
}

// OBB: order-preserving compaction of the boxes whose centre lies inside the mesh. The compacted box q is then
// paired with score[q] / level[q] of the ORIGINAL position q -- the reference's behaviour (utils.py:359-367 drops
// rows from `boxes` only; rpn.py:348-351 keeps indexing scores/levels with the shortened positions).
__global__ void __launch_bounds__(1024) rpn_compact_obb_kernel(int n, float min_size, float score_thresh,
                                                               const int* __restrict__ clevel, const int* __restrict__ cflag,
                                                               const float* __restrict__ cscore, const float* __restrict__ cbox,
                                                               float* __restrict__ fbox, int* __restrict__ group) {
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int per = ceil_div(n, 1024);
    const int lo = t * per, hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += cflag[i];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int v = (t >= off) ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const int total = part[1023];
    int q = part[t] - s;
    for (int i = lo; i < hi; ++i) {
        if (cflag[i]) {
            const float* b = cbox + (size_t)i * 8;
            const bool big = (b[3] >= min_size) && (b[4] >= min_size) && (b[5] >= min_size);
            const bool ok = big && (cscore[q] >= score_thresh);
            group[q] = ok ? clevel[q] : kIgnoreGroup;
#pragma unroll
            for (int k = 0; k < 7; ++k) fbox[(size_t)q * 7 + k] = b[k];
            ++q;
        }
    }
    __syncthreads();
    for (int i = total + t; i < n; i += 1024) {
        group[i] = kIgnoreGroup;
        for (int k = 0; k < 7; ++k) fbox[(size_t)i * 7 + k] = 0.f;
    }
}

__global__ void rpn_emit_kernel(const int64_t* __restrict__ keep, const int32_t* __restrict__ n_keep, int post_top_n,
                                int box_dim, const float* __restrict__ fbox, const float* __restrict__ cscore,
                                const int* __restrict__ clevel, float* __restrict__ boxes, float* __restrict__ scores,
                                float* __restrict__ levels, int32_t* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = min(*n_keep, post_top_n);
    if (p == 0) *count = m;
    if (p >= post_top_n) return;
    if (p < m) {
        const int src = (int)keep[p];
        for (int k = 0; k < box_dim; ++k) boxes[(size_t)p * box_dim + k] = fbox[(size_t)src * box_dim + k];
        scores[p] = cscore[src]; levels[p] = (float)clevel[src];
    } else {
        for (int k = 0; k < box_dim; ++k) boxes[(size_t)p * box_dim + k] = 0.f;
        scores[p] = 0.f; levels[p] = 0.f;
    }
}

static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct RpnWs {
    unsigned* hist; SelState* st; unsigned* counters; unsigned long long* cand;
    float* cbox; float* cscore; int* clevel; int* cflag; float* fbox; int* group; int64_t* keep; int32_t* n_keep;
    void* nms_ws; size_t nms_bytes; size_t total;
};

static RpnWs rpn_layout(void* base, int total_cand) {
    RpnWs w; size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    char* b = (char*)base;
    const int cpad = next_pow2(total_cand < 2 ? 2 : total_cand);
    w.hist = (unsigned*)(b + take((size_t)NRPN_RPN_MAX_LEVELS * kBins * 4));
    w.st = (SelState*)(b + take(NRPN_RPN_MAX_LEVELS * sizeof(SelState)));
    w.counters = (unsigned*)(b + take(NRPN_RPN_MAX_LEVELS * 4));
    w.cand = (unsigned long long*)(b + take((size_t)cpad * 8));
    w.cbox = (float*)(b + take((size_t)total_cand * 8 * 4));
    w.cscore = (float*)(b + take((size_t)total_cand * 4));
    w.clevel = (int*)(b + take((size_t)total_cand * 4));
    w.cflag = (int*)(b + take((size_t)total_cand * 4));
    w.fbox = (float*)(b + take((size_t)total_cand * 7 * 4));
    w.group = (int*)(b + take((size_t)total_cand * 4));
    w.keep = (int64_t*)(b + take((size_t)total_cand * 8));
    w.n_keep = (int32_t*)(b + take(256));
    w.nms_bytes = nms_workspace_bytes(total_cand);
    w.nms_ws = (void*)(b + take(w.nms_bytes));
    w.total = off;
    return w;
}

static int build_dev(const nrpn_rpn_desc* d, RpnDev& P) {
    if (!d || d->n_levels < 1 || d->n_levels > NRPN_RPN_MAX_LEVELS || d->num_anchors < 1 || d->num_anchors > 16)
        return NRPN_ERR_INVALID;
    if (d->pre_nms_top_n < 1 || d->post_nms_top_n < 1) return NRPN_ERR_INVALID;
    P.n_levels = d->n_levels; P.A = d->num_anchors; P.rotated = d->rotated ? 1 : 0; P.code = d->rotated ? 8 : 6;
    int off = 0;
    for (int l = 0; l < d->n_levels; ++l) {
        const nrpn_rpn_level& s = d->level[l];
        LevelDev& L = P.lv[l];
        if (s.gx < 1 || s.gy < 1 || s.gz < 1 || s.sx < 1 || s.sy < 1 || s.sz < 1) return NRPN_ERR_INVALID;
        if (s.ld < P.A * (1 + P.code)) return NRPN_ERR_INVALID;
        const long cnt = (long)s.gx * s.gy * s.gz * P.A;
        if (cnt >= (1L << kIdxBits)) return NRPN_ERR_UNSUPPORTED;
        L.pred = s.pred; L.ld = s.ld; L.gx = s.gx; L.gy = s.gy; L.gz = s.gz; L.sx = s.sx; L.sy = s.sy; L.sz = s.sz;
        L.lim_x = (d->valid[0] + s.sx - 1) / s.sx; L.lim_y = (d->valid[1] + s.sy - 1) / s.sy; L.lim_z = (d->valid[2] + s.sz - 1) / s.sz;
        L.count = (int)cnt; L.k = d->pre_nms_top_n < L.count ? d->pre_nms_top_n : L.count; L.cand_off = off;
        off += L.k;
        for (int a = 0; a < 16; ++a) for (int k = 0; k < 6; ++k) P.cell[l][a][k] = d->cell_anchors[l][a][k];
    }
    P.total_cand = off;
    for (int k = 0; k < 3; ++k) P.mesh[k] = (float)d->mesh[k];
    P.min_size = d->min_size; P.score_thresh = d->score_thresh;
    return NRPN_OK;
}

}  // namespace nrpn

using namespace nrpn;

extern "C" {
#pragma GCC visibility push(default)

size_t nrpn_rpn_workspace_bytes(const nrpn_rpn_desc* desc) {
    RpnDev P;
    if (build_dev(desc, P) != NRPN_OK) return 0;
    return rpn_layout(nullptr, P.total_cand).total + 256;
}

int nrpn_rpn_proposals(const nrpn_rpn_desc* desc, float* boxes, float* scores, float* levels, int32_t* count,
                       void* workspace, size_t workspace_bytes, nrpn_stream_t stream) {
    RpnDev P;
    int rc = build_dev(desc, P);
    if (rc) return rc;
    if (!boxes || !scores || !levels || !count || !workspace) return NRPN_ERR_INVALID;
    for (int l = 0; l < P.n_levels; ++l) if (!P.lv[l].pred) return NRPN_ERR_INVALID;
    if (P.total_cand > nrpn_nms_max_boxes()) return NRPN_ERR_UNSUPPORTED;
    void* base = (void*)align_up((size_t)workspace, 256);
    RpnWs w = rpn_layout(base, P.total_cand);
    if (workspace_bytes < w.total + ((char*)base - (char*)workspace)) return NRPN_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = P.n_levels, M = P.total_cand;
    const int cpad = next_pow2(M < 2 ? 2 : M);

    NRPN_CUDA_TRY(cudaMemsetAsync(w.hist, 0, (size_t)NRPN_RPN_MAX_LEVELS * kBins * 4, st));
    NRPN_CUDA_TRY(cudaMemsetAsync(w.counters, 0, NRPN_RPN_MAX_LEVELS * 4, st));
    RpnSrc src{P};
    topk_init_kernel<RpnSrc><<<1, 32, 0, st>>>(src, w.st);
    NRPN_LAUNCH_CHECK();

    int max_count = 0;
    for (int l = 0; l < L; ++l) max_count = P.lv[l].count > max_count ? P.lv[l].count : max_count;
    int bx = ceil_div(max_count, 256 * 8);
    if (bx > 2 * num_sms()) bx = 2 * num_sms();
    if (bx < 1) bx = 1;
    for (int pass = 0; pass < kPasses; ++pass) {
        topk_hist_kernel<RpnSrc><<<dim3(bx, L), 256, 0, st>>>(src, pass, w.st, w.hist);
        NRPN_LAUNCH_CHECK();
        topk_select_kernel<<<L, 1024, 0, st>>>(pass, w.st, w.hist);
        NRPN_LAUNCH_CHECK();
    }
    fill_u64_kernel<<<ceil_div(cpad, 256), 256, 0, st>>>(w.cand, cpad, ~0ull);
    NRPN_LAUNCH_CHECK();
    topk_collect_kernel<RpnSrc><<<dim3(bx, L), 256, 0, st>>>(src, w.st, w.counters, w.cand);
    NRPN_LAUNCH_CHECK();
    rc = bitonic_sort_u64(w.cand, cpad, st);
    if (rc) return rc;
    rpn_decode_kernel<<<ceil_div(M, 128), 128, 0, st>>>(P, w.cand, w.cbox, w.cscore, w.clevel, w.cflag);
    NRPN_LAUNCH_CHECK();
    const int box_dim = P.rotated ? 7 : 6;
    if (!P.rotated) {
        rpn_group_aabb_kernel<<<ceil_div(M, 256), 256, 0, st>>>(M, w.clevel, w.cflag, w.cbox, w.fbox, w.group);
    } else {
        rpn_compact_obb_kernel<<<1, 1024, 0, st>>>(M, P.min_size, P.score_thresh, w.clevel, w.cflag, w.cscore, w.cbox, w.fbox, w.group);
    }
    NRPN_LAUNCH_CHECK();
    rc = nms_run(w.fbox, box_dim, w.cscore, w.group, M, desc->nms_thresh, kIgnoreGroup, w.keep, w.n_keep, w.nms_ws, w.nms_bytes, st,
                 /*max_group=*/desc->pre_nms_top_n);      // one group per pyramid level, each at most pre_nms_top_n candidates
    if (rc) return rc;
    rpn_emit_kernel<<<ceil_div(desc->post_nms_top_n, 256), 256, 0, st>>>(w.keep, w.n_keep, desc->post_nms_top_n, box_dim, w.fbox,
                                                                       w.cscore, w.clevel, boxes, scores, levels, count);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

#pragma GCC visibility pop
}  // extern "C"
