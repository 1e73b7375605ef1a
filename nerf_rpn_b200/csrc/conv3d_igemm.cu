// 3-D convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces every nn.Conv3d (+ BatchNorm3d + ReLU + residual / FPN top-down add) of the reference path
// (feature_extractor.py:31-68,145-235; anchor.py:177-213), which runs them as separate cuDNN / ATen kernels on
// fp32 NCDHW tensors.
//
// GEMM view:  D[M = voxels, N = Cout] = sum over (tap, Cin)  A[voxel + tap_offset, Cin] * W[tap][Cout, Cin]
//   - M tile  = a brick of bx*by*bz = 128 output voxels.  The A operand of one (tap, 64-channel) k-block is the
//     same brick shifted by the tap offset: ONE 5-D TMA box load {64 ch, bz, by, bx, 1} from the channels-last
//     activation tensor, out-of-bounds voxels zero-filled by TMA (= the convolution's zero padding).  The box
//     lands in shared memory as 128 rows x 128 bytes in the 128B-swizzled K-major layout tcgen05.mma consumes:
//     no im2col buffer, no index arithmetic on the SM.
//   - N tile  = 64 / 128 / 256 output channels; W is pre-packed (tap, Cout, Cin) so a k-block of weights is a
//     3-D TMA box {64, N, 1}.
//   - warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (converged warps, one elect.sync lane issues the async
//     instructions), warps 2-5 = epilogue.  smem ring of 2-8 stages (full/empty mbarriers), accumulators double-buffered
//     in TMEM (2 x N columns) so the epilogue of tile i overlaps the main loop of tile i+1.  Persistent CTAs walk a static
//     tile list that may span up to 4 pyramid levels sharing the same weights (the RPN head on P2..P5 is ONE launch per
//     layer).  Prologue (barriers, TMEM, tensor-map prefetch) runs ahead of griddepcontrol.wait (programmatic dependent launch).
//   - epilogue (fused): + shift (bias / folded BN), + residual (same shape, or nearest-upsampled coarser level for the FPN
//     top-down path), ReLU / GELU, convert to bf16 / fp16 / fp32; 256-bit register stores, or -- TMA_EPI variant of the 1^3
//     layers -- residual tiles prefetched by TMA into shared memory and output tiles stored with cp.async.bulk.tensor.
//   - variants: <256,4,1> <128,6,1> <64,8,1> (long reductions; <64,8,1> also for layers with fewer tiles than SMs),
//     <64,2,3> / <64,2,2,TMA_EPI> (reductions of at most 8 k-blocks), optional split-K (NRPN_SPLITK=1, measured slower).
#include <cstdlib>
#include "common.cuh"
#include "tcgen05.cuh"
#include "conv_internal.cuh"

namespace nrpn {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;       // 16 KB

struct ConvLevelDev {
    int n, xo, yo, zo;
    int bx, by, bz;          // brick (product 128)
    int tx, ty, tz;          // bricks per axis
    int tile_begin;          // first m-tile of this level
    int xr, yr, zr;          // residual extent
    float rsx, rsy, rsz;     // nearest-neighbour scale = in / out
    int ldy, ldr;
    int wide_y, wide_r;      // rows of y / res are 32-byte aligned: 256-bit stores / loads
    int res_tma;             // TMA-epilogue variant: the residual has the output's extent and arrives through maps.r
    void* y;
    const __nv_bfloat16* res;
};

struct ConvDev {
    int n_levels, n_taps, kc_blocks, n_tiles_n, cout, relu, out_fp32, fp16;
    int total_m_tiles;
    int splits, cout_pad;      // split-K: `splits` CTAs share one output tile and reduce through `ws`
    float* ws;                 // fp32 (tiles, splits, 128, BLOCK_N) partial tiles, rewritten by every launch
    unsigned* counters;        // one arrival counter per output tile, zero between launches
    signed char tap[NRPN_CONV_MAX_TAPS][4];
    const float* shift;
    ConvLevelDev lv[NRPN_CONV_MAX_LEVELS];
};

struct ConvMaps {
    CUtensorMap x[NRPN_CONV_MAX_LEVELS];
    CUtensorMap w;
    CUtensorMap y[NRPN_CONV_MAX_LEVELS];      // TMA-epilogue variant only: output tiles are stored with cp.async.bulk.tensor
    CUtensorMap r[NRPN_CONV_MAX_LEVELS];      // ... and same-shape residual tiles are prefetched by the producer warp
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

constexpr int kEpiTileBytes = kBlockM * 64 * 2;      // one 128 x 64 bf16 tile (residual in / output out of the TMA epilogue)

// TMA_EPI (N = 64 tiles of the HBM-bound 1^3 layers, bf16 output, Cout % 64 == 0): the epilogue never touches global memory
// from a register.  The producer warp prefetches the residual tile of a same-shape skip connection into shared memory
// (2-deep ring) together with the operands; the epilogue warps add it from shared memory, write the bf16 tile into a
// swizzled staging buffer and one elected thread stores it with cp.async.bulk.tensor (hardware clips partial bricks).
// With per-thread 16/32-byte global loads and stores the same layers sat at ~45 % of the HBM rate (ncu: lts 45 %, 2-3x the
// compulsory L2 sectors, a DRAM-latency chain per tile).
//
// WS (weight split, the <=1e-3 parity mode of the backbone): every weight is the sum of TWO 16-bit numbers, w = hi + lo with
// lo = round16(w - hi), packed as (taps, 2, CoutPad, Cin).  One TMA box {64, BLOCK_N, 2} brings both planes of a k-block; the
// issuer runs the k-block's MMAs twice on the SAME activation tile (B = hi, then B = lo) into the same accumulator, so the
// weights enter the fp32 accumulation with ~22 significant bits and only the activations carry 16-bit rounding error.
template <int BLOCK_N, int STAGES, int MIN_BLOCKS, bool TMA_EPI = false, bool WS = false>
__global__ void __launch_bounds__(192, MIN_BLOCKS) conv3d_igemm_kernel(const __grid_constant__ ConvMaps maps, const ConvDev P) {
    static_assert(!TMA_EPI || BLOCK_N == 64, "the TMA epilogue stages 128 x 64 tiles");
    constexpr int kBPlane = BLOCK_N * kBlockK * 2;
    constexpr int kBBytes = kBPlane * (WS ? 2 : 1);
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr int RES_SLOTS = (TMA_EPI && WS) ? 1 : 2;              // residual prefetch depth (shared-memory budget of 2 CTAs / SM)
    constexpr int kEpiBytes = TMA_EPI ? (RES_SLOTS + 1) * kEpiTileBytes : 0;      // residual ring + output staging (1)
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;
    const uint32_t kIdesc = P.fp16 ? ptx::make_idesc_f16(kBlockM, BLOCK_N) : ptx::make_idesc_bf16(kBlockM, BLOCK_N);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_res = smem + STAGES * kStageBytes;                 // [2][128 rows][128 B], 128B-swizzled (TMA_EPI)
    uint8_t* smem_out = smem_res + RES_SLOTS * kEpiTileBytes;        // [128 rows][128 B], 128B-swizzled (TMA_EPI)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes + kEpiBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint64_t* rfull_bar = bars + 2 * STAGES + 4;
    uint64_t* rempty_bar = bars + 2 * STAGES + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 128); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&rfull_bar[a], 1); ptx::mbar_init(&rempty_bar[a], 128); }
        ptx::fence_barrier_init();
        for (int l = 0; l < P.n_levels; ++l) ptx::prefetch_tmap(&maps.x[l]);
        if (TMA_EPI) for (int l = 0; l < P.n_levels; ++l) { ptx::prefetch_tmap(&maps.y[l]); ptx::prefetch_tmap(&maps.r[l]); }
        ptx::prefetch_tmap(&maps.w);
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, kTmemCols); ptx::tmem_relinquish(); }
    ptx::pdl_trigger();                 // the next layer may start its own prologue as our CTAs retire
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();                    // everything above overlapped the previous layer's tail; its outputs are visible from here

    const int total_tiles = P.total_m_tiles * P.n_tiles_n * P.splits;     // work items = (output tile, k-split)
    const int kblocks = P.n_taps * P.kc_blocks;
    uint32_t* ticket_slot = tmem_slot + 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (whole warp walks, one elected lane issues)
        {
            const bool leader = ptx::elect_one();
            int stage = 0; uint32_t phase = 0;
            int rslot = 0; uint32_t rphase = 0;
            for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
                const int split = item % P.splits, tile = item / P.splits;
                const int kb0 = (kblocks * split) / P.splits, kb1 = (kblocks * (split + 1)) / P.splits;
                const int n_tile = tile % P.n_tiles_n, m_tile = tile / P.n_tiles_n;
                int l = 0;
#pragma unroll
                for (int i = 1; i < NRPN_CONV_MAX_LEVELS; ++i) if (i < P.n_levels && m_tile >= P.lv[i].tile_begin) l = i;
                const ConvLevelDev& L = P.lv[l];
                int t = m_tile - L.tile_begin;
                const int tiz = t % L.tz; t /= L.tz;
                const int tiy = t % L.ty; t /= L.ty;
                const int tix = t % L.tx; const int nb = t / L.tx;
                const int x0 = tix * L.bx, y0 = tiy * L.by, z0 = tiz * L.bz, n0 = n_tile * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int tap = kb / P.kc_blocks, kc = kb - tap * P.kc_blocks;
                    const int dx = P.tap[tap][0], dy = P.tap[tap][1], dz = P.tap[tap][2];
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (leader) {
                        uint8_t* sa = smem + stage * kStageBytes;
                        uint8_t* sb = sa + kABytes;
                        ptx::mbar_expect_tx(&full_bar[stage], kStageBytes);
                        ptx::tma_load_5d(sa, &maps.x[l], &full_bar[stage], kc * kBlockK, z0 + dz, y0 + dy, x0 + dx, nb);
                        ptx::tma_load_3d(sb, &maps.w, &full_bar[stage], kc * kBlockK, n0, WS ? 2 * tap : tap);   // WS: box depth 2 = hi, lo
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (TMA_EPI && L.res_tma) {                          // residual tile of this output tile, two tiles ahead of its use
                    ptx::mbar_wait(&rempty_bar[rslot], rphase ^ 1u);
                    if (leader) {
                        ptx::mbar_expect_tx(&rfull_bar[rslot], kEpiTileBytes);
                        ptx::tma_load_5d(smem_res + rslot * kEpiTileBytes, &maps.r[l], &rfull_bar[rslot], n0, z0, y0, x0, nb);
                    }
                    __syncwarp();
                    if (++rslot == RES_SLOTS) { rslot = 0; rphase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp walks the warp-uniform loops and waits; one elected lane (elect.sync) issues tcgen05.mma and
        // tcgen05.commit, which keeps descriptors in uniform registers (no per-instruction waterfall loop: at N = 64 an
        // MMA lasts 48 clk and a divergent `if (lane == 0)` issue path costs more than that, profiles/r01_ncu_slab_*.md).
        {
            const bool leader = ptx::elect_one();
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
                const int split = item % P.splits;
                const int nkb = (kblocks * (split + 1)) / P.splits - (kblocks * split) / P.splits;
                ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < nkb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
                    const uint64_t da = ptx::make_desc_sw128(sa);
                    const uint64_t db = ptx::make_desc_sw128(sa + kABytes);
                    if (leader) {
#pragma unroll
                        for (int h = 0; h < (WS ? 2 : 1); ++h) {
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                // advance 16 elements (32 bytes) along K inside the 128-byte swizzle atom: +2 in the >>4 address field
                                ptx::umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(h * (kBPlane >> 4) + 2 * k), kIdesc, (kb | k | h) ? 1u : 0u);
                            }
                        }
                        ptx::umma_commit(&empty_bar[stage]);        // frees the smem slot once the MMAs have read it
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (leader) ptx::umma_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                   // accumulator row == voxel inside the brick
        int acc = 0; uint32_t acc_phase = 0;
        int rslot = 0; uint32_t rphase = 0;
        for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
            const int tile = item / P.splits;
            const int n_tile = tile % P.n_tiles_n, m_tile = tile / P.n_tiles_n;
            int l = 0;
#pragma unroll
            for (int i = 1; i < NRPN_CONV_MAX_LEVELS; ++i) if (i < P.n_levels && m_tile >= P.lv[i].tile_begin) l = i;
            const ConvLevelDev& L = P.lv[l];
            int t = m_tile - L.tile_begin;
            const int tiz = t % L.tz; t /= L.tz;
            const int tiy = t % L.ty; t /= L.ty;
            const int tix = t % L.tx; const int nb = t / L.tx;
            const int zi = row % L.bz, yi = (row / L.bz) % L.by, xi = row / (L.bz * L.by);
            const int gx = tix * L.bx + xi, gy = tiy * L.by + yi, gz = tiz * L.bz + zi;
            const bool valid = gx < L.xo && gy < L.yo && gz < L.zo;
            const size_t vox = (((size_t)nb * L.xo + gx) * L.yo + gy) * L.zo + gz;
            const __nv_bfloat16* rrow = nullptr;
            if (L.res != nullptr && valid) {
                int rx = gx, ry = gy, rz = gz;
                if (L.xr != L.xo || L.yr != L.yo || L.zr != L.zo) {   // F.interpolate(mode='nearest', size=...)
                    rx = min((int)floorf((float)gx * L.rsx), L.xr - 1);
                    ry = min((int)floorf((float)gy * L.rsy), L.yr - 1);
                    rz = min((int)floorf((float)gz * L.rsz), L.zr - 1);
                }
                rrow = L.res + ((((size_t)nb * L.xr + rx) * L.yr + ry) * L.zr + rz) * L.ldr;
            }
            const int n0 = n_tile * BLOCK_N;

            if (P.splits > 1) {
                // ---- split-K (deterministic): every CTA stores its partial tile into its own fp32 slab of the workspace with
                // plain stores; the last CTA to arrive sums the slabs in split order and runs the normal epilogue.
                const int split = item % P.splits;
                float* slab0 = P.ws + ((size_t)tile * P.splits) * (kBlockM * BLOCK_N);          // slabs of this output tile
                float* wrow = slab0 + (size_t)split * (kBlockM * BLOCK_N) + (size_t)row * BLOCK_N;
                ptx::mbar_wait(&tfull_bar[acc], acc_phase);
                ptx::tc_fence_after();
                const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(t_base + (uint32_t)(c * 32), r);
                    ptx::tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            __stcg(reinterpret_cast<float4*>(wrow + c * 32 + j),
                                   make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                    }
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (row == 0) *ticket_slot = atomicAdd(P.counters + tile, 1u);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const bool last = (*ticket_slot == (uint32_t)(P.splits - 1));
                if (last) {
                    __threadfence();
                    if (valid) {
                        const float* rrow0 = slab0 + (size_t)row * BLOCK_N;
                        for (int ch = n0; ch < n0 + BLOCK_N && ch < P.cout; ch += 8) {
                            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                            for (int sp = 0; sp < P.splits; ++sp) {                                  // fixed order => reproducible
                                const float* wp = rrow0 + (size_t)sp * (kBlockM * BLOCK_N) + (ch - n0);
                                const float4 b0 = __ldcg(reinterpret_cast<const float4*>(wp));
                                const float4 b1 = __ldcg(reinterpret_cast<const float4*>(wp + 4));
                                a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
                                a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
                            }
                            const float4 s0 = __ldg(reinterpret_cast<const float4*>(P.shift + ch));
                            const float4 s1 = __ldg(reinterpret_cast<const float4*>(P.shift + ch + 4));
                            float v[8] = {a0.x + s0.x, a0.y + s0.y, a0.z + s0.z, a0.w + s0.w, a1.x + s1.x, a1.y + s1.y, a1.z + s1.z, a1.w + s1.w};
                            if (rrow != nullptr) {
                                const uint4 rv4 = __ldg(reinterpret_cast<const uint4*>(rrow + ch));
                                const uint32_t* rb = reinterpret_cast<const uint32_t*>(&rv4);
#pragma unroll
                                for (int i = 0; i < 4; ++i) { const float2 f = unpack_act2(rb[i], P.fp16); v[2 * i] += f.x; v[2 * i + 1] += f.y; }
                            }
                            if (P.relu == 1) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
                            } else if (P.relu == 2) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752f));
                            }
                            if (P.out_fp32) {
                                float* o = reinterpret_cast<float*>(L.y) + vox * L.ldy + ch;
                                *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                                *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                            } else {
                                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(L.y) + vox * L.ldy + ch;
                                *reinterpret_cast<uint4*>(o) = make_uint4(pack_act2(v[0], v[1], P.fp16), pack_act2(v[2], v[3], P.fp16),
                                                                          pack_act2(v[4], v[5], P.fp16), pack_act2(v[6], v[7], P.fp16));
                            }
                        }
                    }
                    if (row == 0) P.counters[tile] = 0u;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");     // ticket_slot is reused by the next work item
                continue;
            }
            if constexpr (TMA_EPI) {
                // ---- shared-memory epilogue: residual from the TMA-fed ring, output through a swizzled staging tile + TMA store
                const bool res_smem = L.res_tma != 0;
                ptx::mbar_wait(&tfull_bar[acc], acc_phase);
                if (res_smem) ptx::mbar_wait(&rfull_bar[rslot], rphase);
                ptx::tc_fence_after();
                // the staging tile is free once the previous tile's bulk store has finished READING it
                if (threadIdx.x == 64) ptx::bulk_wait_read0();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
                const uint8_t* rsm = smem_res + rslot * kEpiTileBytes + row * 128;
                uint8_t* osm = smem_out + row * 128;
                const int sw = row & 7;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(t_base + (uint32_t)(c * 32), r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int j = c * 4 + g;                     // 16-byte chunk (8 channels) inside the 128-byte row
                        const int ch = n0 + j * 8;
                        const float4 s0 = __ldg(reinterpret_cast<const float4*>(P.shift + ch));
                        const float4 s1 = __ldg(reinterpret_cast<const float4*>(P.shift + ch + 4));
                        float v[8] = {__uint_as_float(r[g * 8 + 0]) + s0.x, __uint_as_float(r[g * 8 + 1]) + s0.y,
                                      __uint_as_float(r[g * 8 + 2]) + s0.z, __uint_as_float(r[g * 8 + 3]) + s0.w,
                                      __uint_as_float(r[g * 8 + 4]) + s1.x, __uint_as_float(r[g * 8 + 5]) + s1.y,
                                      __uint_as_float(r[g * 8 + 6]) + s1.z, __uint_as_float(r[g * 8 + 7]) + s1.w};
                        if (res_smem || rrow != nullptr) {
                            const uint4 rv4 = res_smem ? *reinterpret_cast<const uint4*>(rsm + ((j ^ sw) << 4))
                                                       : __ldg(reinterpret_cast<const uint4*>(rrow + ch));
                            const uint32_t* rb = reinterpret_cast<const uint32_t*>(&rv4);
#pragma unroll
                            for (int i = 0; i < 4; ++i) { const float2 f = unpack_act2(rb[i], P.fp16); v[2 * i] += f.x; v[2 * i + 1] += f.y; }
                        }
                        if (P.relu == 1) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
                        } else if (P.relu == 2) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752f));
                        }
                        *reinterpret_cast<uint4*>(osm + ((j ^ sw) << 4)) =
                            make_uint4(pack_act2(v[0], v[1], P.fp16), pack_act2(v[2], v[3], P.fp16), pack_act2(v[4], v[5], P.fp16), pack_act2(v[6], v[7], P.fp16));
                    }
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tempty_bar[acc]);
                if (res_smem) { ptx::mbar_arrive(&rempty_bar[rslot]); if (++rslot == RES_SLOTS) { rslot = 0; rphase ^= 1u; } }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                ptx::fence_proxy_async();                            // generic-proxy smem writes -> visible to the TMA engine
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 64) {
                    ptx::tma_store_5d(&maps.y[l], smem_out, n0, tiz * L.bz, tiy * L.by, tix * L.bx, nb);
                    ptx::bulk_commit();
                }
                continue;
            }
            // residual rows are fetched one 32-channel chunk ahead (2 x 32 B or 4 x 16 B per thread in flight) so that the DRAM
            // latency of chunk c+1 hides behind the TMEM load + math + stores of chunk c
            uint4 rv[4];
            auto load_res = [&](int c, uint4 (&dst)[4]) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ch = n0 + c * 32 + h * 16;
                    if (rrow != nullptr && L.wide_r && ch + 16 <= P.cout) {
                        ptx::ld_global_nc_v8(rrow + ch, dst[2 * h], dst[2 * h + 1]);
                    } else {
                        dst[2 * h] = (rrow != nullptr && ch < P.cout) ? __ldg(reinterpret_cast<const uint4*>(rrow + ch)) : make_uint4(0u, 0u, 0u, 0u);
                        dst[2 * h + 1] = (rrow != nullptr && ch + 8 < P.cout) ? __ldg(reinterpret_cast<const uint4*>(rrow + ch + 8)) : make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            };
            load_res(0, rv);
            uint4 rn[4];
            if (BLOCK_N == 64) load_res(1, rn);          // N = 64 (HBM-bound 1^3 layers): the whole residual row is in flight before the wait
            ptx::mbar_wait(&tfull_bar[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(t_base + (uint32_t)(c * 32), r);
                if (BLOCK_N != 64 && c + 1 < BLOCK_N / 32) load_res(c + 1, rn);
                ptx::tmem_ld_wait();
                const int ch0 = n0 + c * 32;
                if (valid && ch0 < P.cout) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {            // 2 halves of 16 channels = 32 bytes of bf16
                        const int chh = ch0 + h * 16;
                        if (chh >= P.cout) break;
                        float v[16];
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const int ch = chh + g * 8;
                            if (ch < P.cout) {
                                const float4 s0 = __ldg(reinterpret_cast<const float4*>(P.shift + ch));
                                const float4 s1 = __ldg(reinterpret_cast<const float4*>(P.shift + ch + 4));
                                const uint32_t* rr = r + h * 16 + g * 8;
                                float* vv = v + g * 8;
                                vv[0] = __uint_as_float(rr[0]) + s0.x; vv[1] = __uint_as_float(rr[1]) + s0.y;
                                vv[2] = __uint_as_float(rr[2]) + s0.z; vv[3] = __uint_as_float(rr[3]) + s0.w;
                                vv[4] = __uint_as_float(rr[4]) + s1.x; vv[5] = __uint_as_float(rr[5]) + s1.y;
                                vv[6] = __uint_as_float(rr[6]) + s1.z; vv[7] = __uint_as_float(rr[7]) + s1.w;
                                if (rrow != nullptr) {
                                    const uint32_t* rb = reinterpret_cast<const uint32_t*>(&rv[2 * h + g]);
#pragma unroll
                                    for (int i = 0; i < 4; ++i) { const float2 f = unpack_act2(rb[i], P.fp16); vv[2 * i] += f.x; vv[2 * i + 1] += f.y; }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[g * 8 + i] = 0.0f;
                            }
                        }
                        if (P.relu == 1) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
                        } else if (P.relu == 2) {            // exact (erf) GELU: Swin MLP, feature_extractor.py:635
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752f));
                        }
                        const bool both = chh + 16 <= P.cout;
                        if (P.out_fp32) {
                            float* o = reinterpret_cast<float*>(L.y) + vox * L.ldy + chh;
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                if (chh + g * 8 >= P.cout) break;
                                const uint4 lo = make_uint4(__float_as_uint(v[g * 8 + 0]), __float_as_uint(v[g * 8 + 1]), __float_as_uint(v[g * 8 + 2]), __float_as_uint(v[g * 8 + 3]));
                                const uint4 hi = make_uint4(__float_as_uint(v[g * 8 + 4]), __float_as_uint(v[g * 8 + 5]), __float_as_uint(v[g * 8 + 6]), __float_as_uint(v[g * 8 + 7]));
                                if (L.wide_y) ptx::st_global_v8(o + g * 8, lo, hi);
                                else { *reinterpret_cast<uint4*>(o + g * 8) = lo; *reinterpret_cast<uint4*>(o + g * 8 + 4) = hi; }
                            }
                        } else {
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(L.y) + vox * L.ldy + chh;
                            const uint4 lo = make_uint4(pack_act2(v[0], v[1], P.fp16), pack_act2(v[2], v[3], P.fp16), pack_act2(v[4], v[5], P.fp16), pack_act2(v[6], v[7], P.fp16));
                            const uint4 hi = make_uint4(pack_act2(v[8], v[9], P.fp16), pack_act2(v[10], v[11], P.fp16), pack_act2(v[12], v[13], P.fp16), pack_act2(v[14], v[15], P.fp16));
                            if (both && L.wide_y) ptx::st_global_v8(o, lo, hi);
                            else { *reinterpret_cast<uint4*>(o) = lo; if (both) *reinterpret_cast<uint4*>(o + 8) = hi; }
                        }
                    }
                }
                if (c + 1 < BLOCK_N / 32) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) rv[g] = rn[g];
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    if (TMA_EPI && threadIdx.x == 64) ptx::bulk_wait_all();          // the staging tile must outlive its last bulk store
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, kTmemCols); }
}

// ------------------------------------------------------------------------------------------------ host
template <int BLOCK_N, int STAGES, int MIN_BLOCKS, bool TMA_EPI = false, bool WS = false>
static int launch_conv(const ConvMaps& maps, const ConvDev& P, int total_tiles, cudaStream_t st) {
    constexpr int smem = STAGES * (kABytes + BLOCK_N * kBlockK * 2 * (WS ? 2 : 1)) + (TMA_EPI ? ((WS ? 1 : 2) + 1) * kEpiTileBytes : 0) + 1024 + 256;
    static_assert((smem + 1024) * MIN_BLOCKS <= 228 * 1024, "shared memory budget");
    static_assert(2 * BLOCK_N * MIN_BLOCKS <= 512, "TMEM budget: 512 columns per SM");
    static bool attr_set = false;
    if (!attr_set) {
        NRPN_CUDA_TRY(cudaFuncSetAttribute(conv3d_igemm_kernel<BLOCK_N, STAGES, MIN_BLOCKS, TMA_EPI, WS>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const int slots = num_sms() * MIN_BLOCKS;
    const int grid = total_tiles < slots ? total_tiles : slots;
    NRPN_CUDA_TRY(launch_pdl(conv3d_igemm_kernel<BLOCK_N, STAGES, MIN_BLOCKS, TMA_EPI, WS>, dim3(grid), dim3(192), smem, st, total_tiles <= 2 * slots, maps, P));
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

static void choose_brick(int xo, int yo, int zo, int& bx, int& by, int& bz) {
    long best = -1;
    for (int ez = 0; ez <= 7; ++ez) for (int ey = 0; ey + ez <= 7; ++ey) {
        const int ex = 7 - ez - ey;
        const int cz = 1 << ez, cy = 1 << ey, cx = 1 << ex;
        if (cz > 256 || cy > 256 || cx > 256) continue;
        const long vol = (long)ceil_div(xo, cx) * cx * ceil_div(yo, cy) * cy * ceil_div(zo, cz) * cz;
        // prefer less padding, then compact bricks (smallest halo => best L2 reuse across the filter taps)
        const long score = vol * 4096 + (long)(cx + 2) * (cy + 2) * (cz + 2);
        if (best < 0 || score < best) { best = score; bx = cx; by = cy; bz = cz; }
    }
}

// Split-K: layers with far fewer output tiles than SMs and a long reduction (late ResNet stages: 5x8x8 .. 10x16x16 voxels,
// K up to 13 824) are latency bound on streaming their weights through a handful of CTAs. The kernel can spread the
// k-blocks of each tile over `splits` CTAs (per-split fp32 slabs, last CTA reduces in a fixed order: bit-reproducible).
// MEASURED on B200 (profiles/r01_layers_splitk_*.txt): the single-CTA slab reduction costs more than the split saves
// (L3.c2 512->512 3^3 at 5x8x8: 0.101 ms unsplit, 0.175 ms with 24 splits), so it is OFF unless NRPN_SPLITK=1 is set;
// batching several scenes per launch (bench.py --scenes-per-step) is what feeds these layers instead.
static int choose_splits(int total_tiles, int kblocks, bool short_k) {
    static const bool enabled = [] { const char* e = getenv("NRPN_SPLITK"); return e && e[0] == '1'; }();
    if (!enabled) return 1;
    if (short_k || kblocks < 8 || total_tiles * 2 > num_sms()) return 1;
    int s = num_sms() / total_tiles;
    if (s > kblocks / 4) s = kblocks / 4;
    if (s > 32) s = 32;
    return s < 2 ? 1 : s;
}

struct ConvGeom { int pad_n, cout_pad, block_n, kblocks, m_tiles, n_tiles_n, splits; bool short_k, tma_epi; size_t ws_bytes, counter_bytes; };

static int conv_geometry(const nrpn_conv_desc* d, ConvGeom& g) {
    g.pad_n = nrpn_conv3d_block_n(d->cout);
    g.cout_pad = ceil_div(d->cout, g.pad_n) * g.pad_n;
    g.kblocks = d->n_taps * (d->cin / kBlockK);
    g.short_k = g.kblocks <= 8;
    g.block_n = g.short_k ? 64 : g.pad_n;
    g.n_tiles_n = g.cout_pad / g.block_n;
    int tiles = 0;
    for (int l = 0; l < d->n_levels; ++l) {
        const nrpn_conv_level& S = d->level[l];
        if (S.n < 1 || S.xo < 1 || S.yo < 1 || S.zo < 1) return NRPN_ERR_INVALID;
        int bx, by, bz;
        choose_brick(S.xo, S.yo, S.zo, bx, by, bz);
        tiles += S.n * ceil_div(S.xo, bx) * ceil_div(S.yo, by) * ceil_div(S.zo, bz);
    }
    g.m_tiles = tiles;
    // Late ResNet stages (5x8x8 .. 10x16x16 voxels, K up to 13 824): a handful of 128 x 256 tiles leaves most SMs idle while each
    // tile runs a long reduction.  When even 64-wide tiles fit in one wave, use them: 4x the CTAs, each 128 x 64 MMA lasts 48 clk
    // instead of 128 (same reduction order per output element, bit-identical results).
    if (!g.short_k && g.block_n > 64 && tiles * (g.cout_pad / 64) <= num_sms()) {
        const char* e = getenv("NRPN_CONV_NARROW");
        if (!(e && e[0] == '0')) { g.block_n = 64; g.n_tiles_n = g.cout_pad / 64; }
    }
    g.splits = d->wsplit ? 1 : choose_splits(tiles * g.n_tiles_n, g.kblocks, g.short_k);
    g.counter_bytes = align_up((size_t)tiles * g.n_tiles_n * 4, 256);
    g.ws_bytes = g.splits > 1 ? g.counter_bytes + (size_t)tiles * g.n_tiles_n * g.splits * kBlockM * g.block_n * 4 : 0;
    // shared-memory (TMA) epilogue: the HBM-bound short reductions with whole 64-channel bf16 output tiles
    { const char* e = getenv("NRPN_CONV_TMA_EPI"); g.tma_epi = !(e && e[0] == '0'); }       // A/B switch for tests and profiling
    g.tma_epi = g.tma_epi && g.short_k && g.splits == 1 && !d->out_fp32 && d->cout % 64 == 0;
    for (int l = 0; l < d->n_levels && g.tma_epi; ++l) {
        const nrpn_conv_level& S = d->level[l];
        if (reinterpret_cast<uintptr_t>(S.y) % 16 != 0 || (S.res && reinterpret_cast<uintptr_t>(S.res) % 16 != 0)) g.tma_epi = false;
        // a nearest-upsampled residual (FPN top-down) is gathered per row from registers: the pipelined register epilogue
        // with 3 CTAs per SM is faster for it (measured: lat3+up 0.207 ms vs 0.275 ms)
        if (S.res && (S.xr != S.xo || S.yr != S.yo || S.zr != S.zo)) g.tma_epi = false;
    }
    return NRPN_OK;
}

}  // namespace nrpn

using namespace nrpn;

extern "C" {
#pragma GCC visibility push(default)

size_t nrpn_conv3d_workspace_bytes(const nrpn_conv_desc* d) {
    if (!d || d->cin < 64 || d->cin % 64 != 0 || d->n_levels < 1 || d->n_levels > NRPN_CONV_MAX_LEVELS || d->n_taps < 1) return 0;
    ConvGeom g;
    if (conv_geometry(d, g) != NRPN_OK) return 0;
    return g.ws_bytes;
}

int nrpn_conv3d_block_n(int cout) { return cout <= 64 ? 64 : (cout <= 128 ? 128 : 256); }

const char* nrpn_conv3d_variant(const nrpn_conv_desc* d) {
    if (!d || d->n_taps < 1 || d->n_taps > NRPN_CONV_MAX_TAPS || d->n_levels < 1 || d->n_levels > NRPN_CONV_MAX_LEVELS) return "invalid";
    if (d->cin < 64 || d->cin % 64 != 0 || d->cout < 8 || d->cout % 8 != 0 || (d->stride != 1 && d->stride != 2)) return "unsupported";
    if (!d->wsplit && conv3d_slab_eligible(d)) return "slab<4x16x8,N64>";
    ConvGeom g;
    if (conv_geometry(d, g) != NRPN_OK) return "invalid";
    if (d->wsplit) {
        if (g.short_k) return g.tma_epi ? "igemm<64,2,2,tma-epilogue,wsplit>" : "igemm<64,2,3,wsplit>";
        return g.block_n == 64 ? "igemm<64,6,1,wsplit>" : (g.block_n == 128 ? "igemm<128,4,1,wsplit>" : "igemm<256,2,1,wsplit>");
    }
    if (g.short_k) return g.tma_epi ? "igemm<64,2,2,tma-epilogue>" : "igemm<64,2,3>";
    return g.block_n == 64 ? "igemm<64,8,1>" : (g.block_n == 128 ? "igemm<128,6,1>" : "igemm<256,4,1>");
}

int nrpn_conv3d_fprop(const nrpn_conv_desc* d, nrpn_stream_t stream) {
    if (!d || !d->w || !d->shift) return NRPN_ERR_INVALID;
    if (d->cin < 64 || d->cin % 64 != 0 || d->cout < 8 || d->cout % 8 != 0) return NRPN_ERR_UNSUPPORTED;
    if (d->n_taps < 1 || d->n_taps > NRPN_CONV_MAX_TAPS || d->n_levels < 1 || d->n_levels > NRPN_CONV_MAX_LEVELS)
        return NRPN_ERR_INVALID;
    if (d->stride != 1 && d->stride != 2) return NRPN_ERR_UNSUPPORTED;
    if (d->stride == 2) {       // strided access is expressed through a sub-sampled tensor map: taps must not move
        for (int t = 0; t < d->n_taps; ++t)
            if (d->tap_off[t][0] || d->tap_off[t][1] || d->tap_off[t][2]) return NRPN_ERR_UNSUPPORTED;
    }
    EncodeTiledFn encode = get_encode();
    if (!encode) return NRPN_ERR_NO_DEVICE;
    if (!d->wsplit && conv3d_slab_eligible(d)) return conv3d_slab_launch(d, (cudaStream_t)stream);

    // Weights are padded to nrpn_conv3d_block_n(cout) (a multiple of 64). Long reductions (3^3 taps) use the widest N tile
    // that fits, one CTA per SM, deep smem ring: tensor-pipe bound.  Short reductions (1^3 convs: at most 8 k-blocks) are
    // HBM / epilogue-latency bound: they run N = 64 tiles with 3 CTAs per SM so that 3x more loads/stores are in flight.
    ConvGeom geo;
    { const int rc = conv_geometry(d, geo); if (rc) return rc; }
    const int cout_pad = geo.cout_pad;
    const bool short_k = geo.short_k;
    const int block_n = geo.block_n;
    int splits = geo.splits;
    if (splits > 1 && (!d->workspace || d->workspace_bytes < geo.ws_bytes)) splits = 1;      // no (or too small a) workspace: run unsplit
    ConvMaps maps;
    ConvDev P;
    memset(&P, 0, sizeof(P));
    P.n_levels = d->n_levels; P.n_taps = d->n_taps; P.kc_blocks = d->cin / kBlockK; P.n_tiles_n = cout_pad / block_n;
    P.cout = d->cout; P.relu = d->relu; P.out_fp32 = d->out_fp32; P.shift = d->shift; P.fp16 = d->act_fp16 ? 1 : 0;
    for (int t = 0; t < d->n_taps; ++t) { P.tap[t][0] = d->tap_off[t][0]; P.tap[t][1] = d->tap_off[t][1]; P.tap[t][2] = d->tap_off[t][2]; P.tap[t][3] = 0; }

    int tiles = 0;
    const int s = d->stride;
    for (int l = 0; l < d->n_levels; ++l) {
        const nrpn_conv_level& S = d->level[l];
        ConvLevelDev& L = P.lv[l];
        if (!S.x || !S.y || S.n < 1 || S.xi < 1 || S.yi < 1 || S.zi < 1 || S.xo < 1 || S.yo < 1 || S.zo < 1) return NRPN_ERR_INVALID;
        if (S.ldy < d->cout || S.ldy % 8 != 0) return NRPN_ERR_INVALID;
        if (S.res && (S.ldr < d->cout || S.ldr % 8 != 0 || S.xr < 1 || S.yr < 1 || S.zr < 1)) return NRPN_ERR_INVALID;
        if (s == 2 && (S.xo != (S.xi + 1) / 2 || S.yo != (S.yi + 1) / 2 || S.zo != (S.zi + 1) / 2)) return NRPN_ERR_INVALID;
        L.n = S.n; L.xo = S.xo; L.yo = S.yo; L.zo = S.zo;
        choose_brick(S.xo, S.yo, S.zo, L.bx, L.by, L.bz);
        L.tx = ceil_div(S.xo, L.bx); L.ty = ceil_div(S.yo, L.by); L.tz = ceil_div(S.zo, L.bz);
        L.tile_begin = tiles;
        tiles += S.n * L.tx * L.ty * L.tz;
        L.xr = S.res ? S.xr : S.xo; L.yr = S.res ? S.yr : S.yo; L.zr = S.res ? S.zr : S.zo;
        L.rsx = (float)L.xr / (float)S.xo; L.rsy = (float)L.yr / (float)S.yo; L.rsz = (float)L.zr / (float)S.zo;
        L.ldy = S.ldy; L.ldr = S.ldr; L.y = S.y; L.res = reinterpret_cast<const __nv_bfloat16*>(S.res);
        L.res_tma = (geo.tma_epi && S.res && S.xr == S.xo && S.yr == S.yo && S.zr == S.zo) ? 1 : 0;
        if (geo.tma_epi) {
            // output / residual tensors (C, Z, Y, X, N) with the row pitch ldy / ldr; one box = one 64-channel brick tile
            cuuint32_t ebox[5] = {64, (cuuint32_t)L.bz, (cuuint32_t)L.by, (cuuint32_t)L.bx, 1};
            cuuint32_t eone[5] = {1, 1, 1, 1, 1};
            cuuint64_t ydim[5] = {(cuuint64_t)d->cout, (cuuint64_t)S.zo, (cuuint64_t)S.yo, (cuuint64_t)S.xo, (cuuint64_t)S.n};
            cuuint64_t ystr[4] = {(cuuint64_t)S.ldy * 2, (cuuint64_t)S.zo * S.ldy * 2, (cuuint64_t)S.yo * S.zo * S.ldy * 2,
                                  (cuuint64_t)S.xo * S.yo * S.zo * S.ldy * 2};
            CUresult ry = encode(&maps.y[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, S.y, ydim, ystr, ebox, eone, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (ry != CUDA_SUCCESS) { g_last_cuda_error = (int)ry; return NRPN_ERR_CUDA; }
            if (L.res_tma) {
                cuuint64_t rstr[4] = {(cuuint64_t)S.ldr * 2, (cuuint64_t)S.zo * S.ldr * 2, (cuuint64_t)S.yo * S.zo * S.ldr * 2,
                                      (cuuint64_t)S.xo * S.yo * S.zo * S.ldr * 2};
                CUresult rr = encode(&maps.r[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(S.res), ydim, rstr, ebox, eone,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (rr != CUDA_SUCCESS) { g_last_cuda_error = (int)rr; return NRPN_ERR_CUDA; }
            } else maps.r[l] = maps.y[l];
        }
        const size_t ybytes = (size_t)S.ldy * (d->out_fp32 ? 4 : 2);
        L.wide_y = (ybytes % 32 == 0 && reinterpret_cast<uintptr_t>(S.y) % 32 == 0) ? 1 : 0;
        L.wide_r = (S.res && ((size_t)S.ldr * 2) % 32 == 0 && reinterpret_cast<uintptr_t>(S.res) % 32 == 0) ? 1 : 0;

        // activation tensor map: (C, Z, Y, X, N), optionally sub-sampled by `stride` on the three spatial axes
        const cuuint64_t cin = (cuuint64_t)d->cin;
        cuuint64_t gdim[5] = {cin, (cuuint64_t)ceil_div(S.zi, s), (cuuint64_t)ceil_div(S.yi, s), (cuuint64_t)ceil_div(S.xi, s), (cuuint64_t)S.n};
        cuuint64_t gstr[4] = {cin * 2 * s, (cuuint64_t)S.zi * cin * 2 * s, (cuuint64_t)S.yi * S.zi * cin * 2 * s,
                              (cuuint64_t)S.xi * S.yi * S.zi * cin * 2};
        cuuint32_t box[5] = {(cuuint32_t)kBlockK, (cuuint32_t)L.bz, (cuuint32_t)L.by, (cuuint32_t)L.bx, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&maps.x[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(S.x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return NRPN_ERR_CUDA; }
    }
    for (int l = d->n_levels; l < NRPN_CONV_MAX_LEVELS; ++l) { maps.x[l] = maps.x[0]; maps.y[l] = maps.y[0]; maps.r[l] = maps.r[0]; }
    if (!geo.tma_epi) for (int l = 0; l < NRPN_CONV_MAX_LEVELS; ++l) { maps.y[l] = maps.x[0]; maps.r[l] = maps.x[0]; }
    {
        // wsplit: (taps, 2, CoutPad, cin) -- the hi and lo planes of a tap are consecutive slices, one box of depth 2 fetches both
        const int planes = d->wsplit ? 2 : 1;
        cuuint64_t gdim[3] = {(cuuint64_t)d->cin, (cuuint64_t)cout_pad, (cuuint64_t)d->n_taps * planes};
        cuuint64_t gstr[2] = {(cuuint64_t)d->cin * 2, (cuuint64_t)cout_pad * d->cin * 2};
        cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)block_n, (cuuint32_t)planes};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->w), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return NRPN_ERR_CUDA; }
    }
    P.total_m_tiles = tiles;
    P.splits = splits; P.cout_pad = cout_pad;
    P.counters = reinterpret_cast<unsigned*>(d->workspace);
    P.ws = splits > 1 ? reinterpret_cast<float*>(reinterpret_cast<char*>(d->workspace) + geo.counter_bytes) : nullptr;
    const int total_tiles = tiles * P.n_tiles_n * splits;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->wsplit) {
        if (splits != 1) return NRPN_ERR_UNSUPPORTED;
        if (short_k && geo.tma_epi) return launch_conv<64, 2, 2, true, true>(maps, P, total_tiles, st);
        if (short_k) return launch_conv<64, 2, 3, false, true>(maps, P, total_tiles, st);
        if (block_n == 64) return launch_conv<64, 6, 1, false, true>(maps, P, total_tiles, st);
        if (block_n == 128) return launch_conv<128, 4, 1, false, true>(maps, P, total_tiles, st);
        return launch_conv<256, 2, 1, false, true>(maps, P, total_tiles, st);
    }
    if (short_k && geo.tma_epi && splits == 1) return launch_conv<64, 2, 2, true>(maps, P, total_tiles, st);
    if (short_k) return launch_conv<64, 2, 3>(maps, P, total_tiles, st);
    if (block_n == 64) return launch_conv<64, 8, 1>(maps, P, total_tiles, st);
    if (block_n == 128) return launch_conv<128, 6, 1>(maps, P, total_tiles, st);
    return launch_conv<256, 4, 1>(maps, P, total_tiles, st);
}

#pragma GCC visibility pop
}  // extern "C"
