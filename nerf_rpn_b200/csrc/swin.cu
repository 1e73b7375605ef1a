// Swin-3D support kernels (SwinTransformer_FPN, nerf_rpn/model/feature_extractor.py:382-789). The GEMMs of the transformer
// (qkv, proj, MLP, patch merging, patch embedding) run on the tcgen05 implicit-GEMM kernel as 1x1x1 convolutions over
// channels-last token grids; this file holds what is left around them:
//   patch_embed_pack      fp32 NCDHW grid -> (X/4, Y/4, Z/4, 256) bf16 patch rows    (Conv3d k4 s4 as a GEMM, :731-733)
//   layernorm             per-token LayerNorm over C channels                        (:621,634,736)
//   patch_merge_ln        2x2x2 gather (zero pad on odd extents) + LayerNorm(8C)     (:661-685)
//   window_attention      4^3 shifted-window multi-head attention, head_dim 32        (:382-497)
// Activations are (N, H, W, D, ld) bf16 with C real channels and ld >= C (ld a multiple of 64; channels [C, ld) stay zero).
#include <cstdlib>
#include "common.cuh"
#include "tcgen05.cuh"

namespace nrpn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------- patch embedding
// row (i,j,k) channel ((c*4+px)*4+py)*4+pz = grid[c][4i+px][4j+py][4k+pz]; one thread per 16-byte chunk (8 channels =
// fixed c,px,py half of pz... 8 consecutive channels = pz 0..3 for py and py+1).
__global__ void patch_embed_pack_kernel(const float* __restrict__ grid, int n, int X, int Y, int Z, int H, int W, int D,
                                        __nv_bfloat16* __restrict__ out, int fp16) {
    const size_t total = (size_t)n * H * W * D * 32;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(t & 31);
        size_t v = t >> 5;
        const int k = (int)(v % D); v /= D;
        const int j = (int)(v % W); v /= W;
        const int i = (int)(v % H); const int b = (int)(v / H);
        // chunk s covers channels 8s..8s+7: c = s/8, px = (s/2)%4, py = (s%2)*2 + {0,1}, pz = 0..3
        const int c = s >> 3, px = (s >> 1) & 3, py0 = (s & 1) * 2;
        float val[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int py = py0 + (e >> 2), pz = e & 3;
            val[e] = grid[((((size_t)b * 4 + c) * X + 4 * i + px) * Y + 4 * j + py) * Z + 4 * k + pz];
        }
        uint32_t h[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = pack_act2(val[2 * q], val[2 * q + 1], fp16);
        *reinterpret_cast<uint4*>(out + (t << 3)) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// one warp per token; two passes over the row held in registers (C <= 3072 -> <= 96 values per lane)
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, __nv_bfloat16* __restrict__ out,
                                                        int ld_out, long tokens, int C, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int fp16) {
    const long tok = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tok >= tokens) return;
    const __nv_bfloat16* x = in + tok * ld_in;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += load_act(x + c, fp16);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = load_act(x + c, fp16) - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    __nv_bfloat16* y = out + tok * ld_out;
    for (int c = lane; c < C; c += 32) store_act(y + c, (load_act(x + c, fp16) - mean) * rstd * gamma[c] + beta[c], fp16);
}

// ---------------------------------------------------------------------------------------------- patch merging
// output token (i,j,k) of the (ceil(H/2), ceil(W/2), ceil(D/2)) grid = LN(concat of the 8 input tokens (2i+a, 2j+b, 2k+c)),
// parity order 000,100,010,110,001,101,011,111 over (H,W,D), zero for positions past an odd extent.
__global__ void __launch_bounds__(256) patch_merge_ln_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, int n, int H, int W, int D, int C,
                                                             __nv_bfloat16* __restrict__ out, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int fp16) {
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, Do = (D + 1) / 2;
    const long tokens = (long)n * Ho * Wo * Do;
    const long tok = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tok >= tokens) return;
    long v = tok;
    const int k = (int)(v % Do); v /= Do;
    const int j = (int)(v % Wo); v /= Wo;
    const int i = (int)(v % Ho); const int b = (int)(v / Ho);
    const int C8 = 8 * C;
    auto src = [&](int c8) -> float {
        const int part = c8 / C, c = c8 - part * C;
        const int a = part & 1, bb = (part >> 1) & 1, cc = (part >> 2) & 1;
        const int h = 2 * i + a, w = 2 * j + bb, d = 2 * k + cc;
        if (h >= H || w >= W || d >= D) return 0.f;
        return load_act(in + ((((size_t)b * H + h) * W + w) * D + d) * ld_in + c, fp16);
    };
    float s = 0.f;
    for (int c = lane; c < C8; c += 32) s += src(c);
    const float mean = warp_sum(s) / (float)C8;
    float q = 0.f;
    for (int c = lane; c < C8; c += 32) { const float d = src(c) - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C8 + eps);
    __nv_bfloat16* y = out + tok * C8;
    for (int c = lane; c < C8; c += 32) store_act(y + c, (src(c) - mean) * rstd * gamma[c] + beta[c], fp16);
}

// ---------------------------------------------------------------------------------------------- window attention
// One CTA (64 threads) per (window, head); thread t = query token t of the 4x4x4 window.  K and V of the window are staged
// in shared memory as fp32; scores, softmax and the output row live in registers.  Window w of the zero-padded, cyclically
// shifted grid: padded coordinate p -> source coordinate (p + shift) mod P; sources past the real extent are padding
// tokens whose q/k/v equal the qkv bias (the reference pads AFTER norm1, so padded tokens enter the Linear as zeros).
struct AttnDev {
    const __nv_bfloat16* qkv; int ld_qkv;      // (N,H,W,D, 3C): [q | k | v], each head-major (head*32 + d)
    __nv_bfloat16* out; int ld_out;            // (N,H,W,D, ld): heads concatenated
    const float* qkv_bias;                     // (3C)
    const float* table;                        // (343, heads) relative position bias table
    int n, H, W, D, C, heads;
    int PH, PW, PD, sh, sw, sd;                // padded extents, effective shifts
    int nwh, nww, nwd;
    int fp16;                                  // 16-bit format of qkv / out: 0 bf16, 1 fp16
};

__global__ void __launch_bounds__(64) window_attention_kernel(AttnDev P) {
    __shared__ float Ks[64][33];
    __shared__ float Vs[64][33];
    __shared__ int region[64];
    const int head = blockIdx.y;
    int w = blockIdx.x;
    const int wd = w % P.nwd; w /= P.nwd;
    const int ww = w % P.nww; w /= P.nww;
    const int wh = w % P.nwh; const int b = w / P.nwh;
    const int t = threadIdx.x;
    const int ti = t >> 4, tj = (t >> 2) & 3, tk = t & 3;             // token order inside the window: (i*4 + j)*4 + k
    const int ph = wh * 4 + ti, pw = ww * 4 + tj, pd = wd * 4 + tk;   // position in the padded, shifted grid
    const int sh_ = (ph + P.sh) % P.PH, sw_ = (pw + P.sw) % P.PW, sd_ = (pd + P.sd) % P.PD;   // source position
    const bool real = sh_ < P.H && sw_ < P.W && sd_ < P.D;
    const size_t tok = (((size_t)b * P.H + sh_) * P.W + sw_) * P.D + sd_;
    const __nv_bfloat16* row = P.qkv + tok * P.ld_qkv;
    const int hc = head * 32;
    float q[32];
    const float scale = 0.17677669529663687f;                          // 32^-0.5
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        const float qv = real ? load_act(row + hc + d, P.fp16) : P.qkv_bias[hc + d];
        const float kv = real ? load_act(row + P.C + hc + d, P.fp16) : P.qkv_bias[P.C + hc + d];
        const float vv = real ? load_act(row + 2 * P.C + hc + d, P.fp16) : P.qkv_bias[2 * P.C + hc + d];
        q[d] = qv * scale; Ks[t][d] = kv; Vs[t][d] = vv;
    }
    // shift-mask region of this token (ids as built by the reference: per axis 0 / 1 / 2, an unshifted axis contributes one id)
    auto reg1 = [](int p, int Pext, int s) { return s == 0 ? 2 : (p < Pext - 4 ? 0 : (p < Pext - s ? 1 : 2)); };
    const bool shifted = (P.sh + P.sw + P.sd) > 0;
    region[t] = shifted ? (reg1(ph, P.PH, P.sh) * 9 + reg1(pw, P.PW, P.sw) * 3 + reg1(pd, P.PD, P.sd)) : 0;
    __syncthreads();
    const int myreg = region[t];
    float sc[64];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) a += q[d] * Ks[j][d];
        const int ji = j >> 4, jj = (j >> 2) & 3, jk = j & 3;
        const int idx = (ti - ji + 3) * 49 + (tj - jj + 3) * 7 + (tk - jk + 3);
        a += P.table[idx * P.heads + head];
        if (shifted && region[j] != myreg) a += -100.0f;
        sc[j] = a; mx = fmaxf(mx, a);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.0f / sum;
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        const float p = sc[j] * inv;
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] += p * Vs[j][d];
    }
    if (real) {
        __nv_bfloat16* orow = P.out + tok * P.ld_out + hc;
#pragma unroll
        for (int d = 0; d < 32; d += 2) *reinterpret_cast<uint32_t*>(orow + d) = pack_act2(o[d], o[d + 1], P.fp16);
    }
}

// ---------------------------------------------------------------------------------------------- window attention on tcgen05
// Two 64-token windows of the same head are stacked into ONE M = 128 problem: S = Q K^T is a 128 x 128 tcgen05.mma whose two
// diagonal 64 x 64 blocks are the windows' score matrices (the off-diagonal blocks are never read); softmax runs from TMEM
// (thread r <-> query row r: scale, relative-position bias, -100 region mask, exp); the un-normalised probabilities go back to
// shared memory as bf16 with the other window's block zeroed, and O = P V is a second MMA (M 128, N 32, K 128) against V^T.
// The 1/sum normalisation is applied to the fp32 accumulator.  Operands are written to shared memory by the threads
// themselves (a token's q/k/v are 64 contiguous bytes each) in the 128B-swizzled K-major layout; V is transposed on the way.
constexpr int kAttnThreads = 128;
constexpr int kAttnSmem = 16384 /*Q*/ + 16384 /*K*/ + 2 * 4096 /*V^T*/ + 2 * 16384 /*P*/ + 1024 /*align*/ + 2048 /*table, regions, barriers*/;

__device__ __forceinline__ uint32_t attn_sw(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__global__ void __launch_bounds__(kAttnThreads) window_attention_tc_kernel(AttnDev P, int n_windows) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* Qs = smem;
    uint8_t* Ks = smem + 16384;
    uint8_t* Vt = smem + 32768;                 // [2 tiles][32 rows (d)][64 tokens]
    uint8_t* Ps = smem + 40960;                 // [2 tiles (key halves)][128 rows][64 keys]
    float* tbl = reinterpret_cast<float*>(smem + 73728);          // 343 relative-position biases of this head
    int* region = reinterpret_cast<int*>(smem + 73728 + 1408);    // 128
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 73728 + 1408 + 512);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int head = blockIdx.y;
    const int t = threadIdx.x, warp = t >> 5;
    const int hc = head * 32;
    for (int i = t; i < 343; i += kAttnThreads) tbl[i] = P.table[i * P.heads + head];
    if (t == 0) { ptx::mbar_init(&bars[0], 1); ptx::mbar_init(&bars[1], 1); ptx::fence_barrier_init(); }
    if (warp == 0) { ptx::tmem_alloc(tmem_slot, 128); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const bool shifted = (P.sh + P.sw + P.sd) > 0;
    const int half = t >> 6, tw = t & 63;                          // stacked window (0/1), token inside it
    const int ti = tw >> 4, tj = (tw >> 2) & 3, tk = tw & 3;
    const uint32_t kIdescS = P.fp16 ? ptx::make_idesc_f16(128, 128) : ptx::make_idesc_bf16(128, 128);
    const uint32_t kIdescO = P.fp16 ? ptx::make_idesc_f16(128, 32) : ptx::make_idesc_bf16(128, 32);
    uint32_t phase = 0;
    for (int pair = blockIdx.x; pair * 2 < n_windows; pair += gridDim.x) {
        int w = pair * 2 + half;
        const bool have = w < n_windows;
        const int wd = w % P.nwd; w /= P.nwd;
        const int ww = w % P.nww; w /= P.nww;
        const int wh = w % P.nwh; const int b = w / P.nwh;
        const int ph = wh * 4 + ti, pw = ww * 4 + tj, pd = wd * 4 + tk;   // position in the padded, shifted grid
        const int sh_ = (ph + P.sh) % P.PH, sw_ = (pw + P.sw) % P.PW, sd_ = (pd + P.sd) % P.PD;
        const bool real = have && sh_ < P.H && sw_ < P.W && sd_ < P.D;
        const size_t tok = have ? ((((size_t)b * P.H + sh_) * P.W + sw_) * P.D + sd_) : 0;
        {   // ---- stage q, k (row t of the K-major tiles) and v (column t of V^T)
            uint4 qv[4], kv[4], vv[4];
            if (real) {
                const __nv_bfloat16* row = P.qkv + tok * P.ld_qkv;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    qv[i] = __ldg(reinterpret_cast<const uint4*>(row + hc) + i);
                    kv[i] = __ldg(reinterpret_cast<const uint4*>(row + P.C + hc) + i);
                    vv[i] = __ldg(reinterpret_cast<const uint4*>(row + 2 * P.C + hc) + i);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t a[4], c[4], e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int d = i * 8 + u * 2;
                        a[u] = have ? pack_act2(P.qkv_bias[hc + d], P.qkv_bias[hc + d + 1], P.fp16) : 0u;
                        c[u] = have ? pack_act2(P.qkv_bias[P.C + hc + d], P.qkv_bias[P.C + hc + d + 1], P.fp16) : 0u;
                        e[u] = have ? pack_act2(P.qkv_bias[2 * P.C + hc + d], P.qkv_bias[2 * P.C + hc + d + 1], P.fp16) : 0u;
                    }
                    qv[i] = make_uint4(a[0], a[1], a[2], a[3]); kv[i] = make_uint4(c[0], c[1], c[2], c[3]); vv[i] = make_uint4(e[0], e[1], e[2], e[3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                *reinterpret_cast<uint4*>(Qs + attn_sw(t, i)) = qv[i];
                *reinterpret_cast<uint4*>(Ks + attn_sw(t, i)) = kv[i];
            }
            uint8_t* vt = Vt + half * 4096;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&vv[i]);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int d = i * 8 + u;
                    *reinterpret_cast<__nv_bfloat16*>(vt + attn_sw(d, tw >> 3) + (tw & 7) * 2) = e[u];
                }
            }
            auto reg1 = [](int p, int Pext, int s) { return s == 0 ? 2 : (p < Pext - 4 ? 0 : (p < Pext - s ? 1 : 2)); };
            region[t] = shifted ? (reg1(ph, P.PH, P.sh) * 9 + reg1(pw, P.PW, P.sw) * 3 + reg1(pd, P.PD, P.sd)) : 0;
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t da = ptx::make_desc_sw128(ptx::smem_u32(Qs)), db = ptx::make_desc_sw128(ptx::smem_u32(Ks));
#pragma unroll
                for (int k = 0; k < 2; ++k)                                   // head_dim 32 = two K-steps of 16
                    ptx::umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdescS, k ? 1u : 0u);
                ptx::umma_commit(&bars[0]);
            }
            __syncwarp();
        }
        ptx::mbar_wait(&bars[0], phase);
        ptx::tc_fence_after();
        // ---- softmax of row t over its own window's 64 keys
        float sc[64];
        {
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 64);
            ptx::tmem_ld_32x32(taddr, r0);
            ptx::tmem_ld_32x32(taddr + 32, r1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) { sc[j] = __uint_as_float(r0[j]); sc[32 + j] = __uint_as_float(r1[j]); }
        }
        const int myreg = region[t];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            const int ji = j >> 4, jj = (j >> 2) & 3, jk = j & 3;
            float a = sc[j] * 0.17677669529663687f + tbl[(ti - ji + 3) * 49 + (tj - jj + 3) * 7 + (tk - jk + 3)];
            if (shifted && region[half * 64 + j] != myreg) a += -100.0f;
            sc[j] = a; mx = fmaxf(mx, a);
        }
        float sum = 0.f;
        {
            uint8_t* mine = Ps + half * 16384;
            uint8_t* other = Ps + (1 - half) * 16384;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float e[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { e[u] = __expf(sc[c * 8 + u] - mx); sum += e[u]; }
                uint32_t pk[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) pk[u] = pack_act2(e[2 * u], e[2 * u + 1], P.fp16);
                *reinterpret_cast<uint4*>(mine + attn_sw(t, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(other + attn_sw(t, c)) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncthreads();                                   // every S row has been read: its TMEM columns may be overwritten by O
        if (warp == 0) {
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t da = ptx::make_desc_sw128(ptx::smem_u32(Ps + h * 16384));
                    const uint64_t db = ptx::make_desc_sw128(ptx::smem_u32(Vt + h * 4096));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdescO, (h | k) ? 1u : 0u);
                }
                ptx::umma_commit(&bars[1]);
            }
            __syncwarp();
        }
        ptx::mbar_wait(&bars[1], phase);
        ptx::tc_fence_after();
        {
            uint32_t o[32];
            ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), o);
            ptx::tmem_ld_wait();
            if (real) {
                const float inv = 1.0f / sum;
                __nv_bfloat16* orow = P.out + tok * P.ld_out + hc;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t pk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        pk[u] = pack_act2(__uint_as_float(o[i * 8 + 2 * u]) * inv, __uint_as_float(o[i * 8 + 2 * u + 1]) * inv, P.fp16);
                    *(reinterpret_cast<uint4*>(orow) + i) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
        ptx::tc_fence_before();
        __syncthreads();                                   // O has been read, the operand tiles may be rewritten
        ptx::tc_fence_after();
        phase ^= 1u;
    }
    if (warp == 0) ptx::tmem_dealloc(tmem, 128);
}

static inline unsigned grid_for(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    const size_t cap = (size_t)num_sms() * 16;
    return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace nrpn

using namespace nrpn;

extern "C" {
#pragma GCC visibility push(default)

int nrpn_patch_embed_pack(const float* grid, int n, int x, int y, int z, void* out, int act_fp16, nrpn_stream_t stream) {
    if (!grid || !out || n < 1 || x < 4 || y < 4 || z < 4) return NRPN_ERR_INVALID;
    const int H = x / 4, W = y / 4, D = z / 4;
    const size_t total = (size_t)n * H * W * D * 32;
    patch_embed_pack_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(grid, n, x, y, z, H, W, D, reinterpret_cast<__nv_bfloat16*>(out), act_fp16 ? 1 : 0);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

int nrpn_layernorm(const void* in, int ld_in, void* out, int ld_out, long tokens, int c, const float* gamma, const float* beta,
                   float eps, int act_fp16, nrpn_stream_t stream) {
    if (!in || !out || !gamma || !beta || tokens < 1 || c < 1 || ld_in < c || ld_out < c) return NRPN_ERR_INVALID;
    layernorm_kernel<<<(unsigned)ceil_div(tokens, 8L), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(in), ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out, tokens, c, gamma, beta, eps, act_fp16 ? 1 : 0);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

int nrpn_patch_merge_ln(const void* in, int ld_in, int n, int h, int w, int d, int c, void* out, const float* gamma, const float* beta,
                        float eps, int act_fp16, nrpn_stream_t stream) {
    if (!in || !out || !gamma || !beta || n < 1 || h < 1 || w < 1 || d < 1 || c < 1 || ld_in < c) return NRPN_ERR_INVALID;
    const long tokens = (long)n * ((h + 1) / 2) * ((w + 1) / 2) * ((d + 1) / 2);
    patch_merge_ln_kernel<<<(unsigned)ceil_div(tokens, 8L), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(in), ld_in, n, h, w, d, c, reinterpret_cast<__nv_bfloat16*>(out), gamma, beta, eps, act_fp16 ? 1 : 0);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

int nrpn_window_attention(const void* qkv, int ld_qkv, void* out, int ld_out, const float* qkv_bias, const float* table, int n, int h,
                          int w, int d, int c, int heads, int shift, int act_fp16, nrpn_stream_t stream) {
    if (!qkv || !out || !qkv_bias || !table || n < 1 || h < 1 || w < 1 || d < 1 || heads < 1 || c != heads * 32) return NRPN_ERR_INVALID;
    if (ld_qkv < 3 * c || ld_out < c || (shift != 0 && shift != 2)) return NRPN_ERR_INVALID;
    AttnDev P;
    P.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv); P.ld_qkv = ld_qkv; P.out = reinterpret_cast<__nv_bfloat16*>(out); P.ld_out = ld_out;
    P.fp16 = act_fp16 ? 1 : 0;
    P.qkv_bias = qkv_bias; P.table = table; P.n = n; P.H = h; P.W = w; P.D = d; P.C = c; P.heads = heads;
    P.PH = (h + 3) / 4 * 4; P.PW = (w + 3) / 4 * 4; P.PD = (d + 3) / 4 * 4;
    P.sh = (4 >= P.PH) ? 0 : shift; P.sw = (4 >= P.PW) ? 0 : shift; P.sd = (4 >= P.PD) ? 0 : shift;   // no shift along an axis one window wide
    P.nwh = P.PH / 4; P.nww = P.PW / 4; P.nwd = P.PD / 4;
    const long windows = (long)n * P.nwh * P.nww * P.nwd;
    if (windows > 0x7fffffffL || heads > 65535) return NRPN_ERR_UNSUPPORTED;
    // tcgen05 path: needs 16-byte aligned 64-byte q/k/v segments; NRPN_ATTN_TC=0 selects the CUDA-core kernel (A/B testing)
    const char* e = getenv("NRPN_ATTN_TC");
    const bool tc = !(e && e[0] == '0') && ld_qkv % 8 == 0 && ld_out % 8 == 0 && c % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(qkv) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    if (tc) {
        static bool attr_set = false;
        if (!attr_set) {
            NRPN_CUDA_TRY(cudaFuncSetAttribute(window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
            attr_set = true;
        }
        const long pairs = (windows + 1) / 2;
        long gx = (long)num_sms() * 3 / heads;              // ~3 resident CTAs per SM in total, each looping over window pairs
        if (gx < 1) gx = 1;
        if (gx > pairs) gx = pairs;
        window_attention_tc_kernel<<<dim3((unsigned)gx, heads), kAttnThreads, kAttnSmem, (cudaStream_t)stream>>>(P, (int)windows);
        NRPN_LAUNCH_CHECK();
        return NRPN_OK;
    }
    window_attention_kernel<<<dim3((unsigned)windows, heads), 64, 0, (cudaStream_t)stream>>>(P);
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

#pragma GCC visibility pop
}  // extern "C"
