// Halo-slab implicit-GEMM convolution for the narrow (64-channel) stride-1 layers of the path: the 7^3 stride-2 stem on
// the packed space-to-depth input (feature_extractor.py:163, 4 x 4 x 2 taps of K = 64) and the 3^3 64->64 convolutions
// of ResNet layer1 (feature_extractor.py:31-68).
//
// Why a second kernel: with N = 64 output channels a tap-by-tap brick kernel (conv3d_igemm.cu) re-reads one 16 KB
// activation box AND one 8 KB weight block from L2 for every (tap, 128-voxel tile): 24 KB per 0.5 MFMA, ~8 TB/s of L2
// traffic at the measured 0.66 ms of the stem -- it is L2-bandwidth bound at a quarter of the tensor-pipe rate.
// Here one CTA owns a tile of 4 x-planes x 16 (y) x 8 (z) = 512 output voxels, i.e. FOUR 128 x 64 accumulators in TMEM:
//   - z is the fastest spatial axis, so 8 z-neighbours x 64 channels form exactly one 1024-byte SWIZZLE_128B atom.  For
//     each distinct z offset of the filter (a "phase") the halo slab {z0+dz .. +8, y0+dy_min .. y0+15+dy_max,
//     x0+dx_min .. x0+3+dx_max} is streamed x-plane by x-plane (one TMA box {64 ch, 8 z, Ys y} per plane, out-of-range
//     voxels zero-filled = the convolution padding) into a ring of shared-memory slots;
//   - the A operand of tap (dx, dy) for accumulator a is the 16 consecutive atoms starting at atom (dy - dy_min) of
//     plane (dx - dx_min + a): a plain SWIZZLE_128B K-major descriptor with a different start address -- shifts along
//     y and x are free, only shifts along z need their own copy of the slab;
//   - every 8 KB weight block is used by 16 MMAs (4 accumulators x K 64/16) instead of 4.
// L2 traffic per 128 output voxels drops from 768 KB to ~130 KB (stem) and the layer becomes tensor-pipe bound.
// Roles: warp 0 = slab (A) producer, warp 1 = tcgen05.mma issuer, warp 2 = weight (B) producer, warps 3-6 = epilogue
// (+ shift, ReLU, bf16, 128-byte row stores); accumulators double-buffered (2 x 4 x 64 = all 512 TMEM columns).
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tcgen05.cuh"
#include "conv_internal.cuh"

namespace nrpn {

constexpr int kSlabAcc = 4;              // x-planes (accumulators) per tile
constexpr int kSlabTileY = 16;
constexpr int kSlabTileZ = 8;
constexpr int kSlabMaxSlots = 12;
constexpr int kSlabBStages = 4;
constexpr int kSlabBBytes = 64 * 64 * 2; // one tap of weights: 64 (Cout, padded) x 64 (Cin) bf16
constexpr int kSlabMaxAxis = 8;          // at most 8 distinct offsets per axis
constexpr int kSlabThreads = 224;
constexpr int kSlabSmemLimit = 227 * 1024;

struct SlabDev {
    int n, xo, yo, zo;
    int tx, ty, tz, total_tiles;
    int n_phases, ndx, ndy;
    int dx0, dy0;
    int planes, slots, plane_bytes;
    int cout, relu, ldy, wide_y, fp16;
    signed char dz[kSlabMaxAxis];
    unsigned char dyrel[kSlabMaxAxis];
    unsigned char widx[kSlabMaxAxis][kSlabMaxAxis][kSlabMaxAxis];     // [phase][dx index][dy index] -> tap index of the packed weights
    const float* shift;
    __nv_bfloat16* y;
};

struct SlabMaps { CUtensorMap x, w; };

__device__ __forceinline__ uint32_t slab_pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kSlabThreads, 1) conv3d_slab_kernel(const __grid_constant__ SlabMaps maps, const SlabDev P) {
    const uint32_t kIdesc = P.fp16 ? ptx::make_idesc_f16(128, 64) : ptx::make_idesc_bf16(128, 64);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)P.slots * P.plane_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kSlabBStages * kSlabBBytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + kSlabMaxSlots;
    uint64_t* b_full = bars + 2 * kSlabMaxSlots;
    uint64_t* b_empty = b_full + kSlabBStages;
    uint64_t* tfull = b_empty + kSlabBStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < P.slots; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kSlabBStages; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 128); }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&maps.x);
        ptx::prefetch_tmap(&maps.w);
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::pdl_trigger();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::pdl_wait();                    // prologue overlapped the previous kernel's tail; its outputs are visible from here

    if (warp == 0) {
        // ---------------------------------------------------------------- slab producer: one x-plane per TMA
        {
            const bool leader = ptx::elect_one();
            int slot = 0; uint32_t par = 0;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
                int t = tile;
                const int tiz = t % P.tz; t /= P.tz;
                const int tiy = t % P.ty; t /= P.ty;
                const int tix = t % P.tx; const int nb = t / P.tx;
                const int x0 = tix * kSlabAcc + P.dx0, y0 = tiy * kSlabTileY + P.dy0, z0 = tiz * kSlabTileZ;
                for (int ph = 0; ph < P.n_phases; ++ph) {
                    const int z = z0 + P.dz[ph];
                    for (int p = 0; p < P.planes; ++p) {
                        ptx::mbar_wait(&a_empty[slot], par ^ 1u);
                        if (leader) {
                            ptx::mbar_expect_tx(&a_full[slot], (uint32_t)P.plane_bytes);
                            ptx::tma_load_5d(smem_a + (size_t)slot * P.plane_bytes, &maps.x, &a_full[slot], 0, z, y0, x0 + p, nb);
                        }
                        __syncwarp();
                        if (++slot == P.slots) { slot = 0; par ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ---------------------------------------------------------------- weight producer: one tap (8 KB) per TMA
        {
            const bool leader = ptx::elect_one();
            int stage = 0; uint32_t par = 0;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
                for (int ph = 0; ph < P.n_phases; ++ph)
                    for (int dxi = 0; dxi < P.ndx; ++dxi)
                        for (int dyi = 0; dyi < P.ndy; ++dyi) {
                            ptx::mbar_wait(&b_empty[stage], par ^ 1u);
                            if (leader) {
                                ptx::mbar_expect_tx(&b_full[stage], kSlabBBytes);
                                ptx::tma_load_3d(smem_b + stage * kSlabBBytes, &maps.w, &b_full[stage], 0, 0, (int)P.widx[ph][dxi][dyi]);
                            }
                            __syncwarp();
                            if (++stage == kSlabBStages) { stage = 0; par ^= 1u; }
                        }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        // The whole warp walks the (warp-uniform) loops and waits; one elected lane issues tcgen05.mma / commit.  Keeping
        // the control flow converged lets ptxas hold the descriptors in uniform registers: at N = 64 an MMA lasts only
        // 48 clk, so the issue path (not the tensor pipe) is what limits this kernel if every MMA needs a waterfall loop.
        const bool leader = ptx::elect_one();
        const uint32_t sa0 = ptx::smem_u32(smem_a), sb0 = ptx::smem_u32(smem_b);
        int a_slot = 0; uint32_t a_par = 0;          // ring position of plane 0 of the current phase
        int b_stage = 0; uint32_t b_par = 0;
        int buf = 0; uint32_t buf_par = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
            ptx::mbar_wait(&tempty[buf], buf_par ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_base = tmem_base + (uint32_t)(buf * kSlabAcc * 64);
            uint32_t accumulate = 0;
            for (int ph = 0; ph < P.n_phases; ++ph) {
                int waited = 0;
                for (int dxi = 0; dxi < P.ndx; ++dxi) {
                    while (waited <= dxi + kSlabAcc - 1) {          // planes become visible in ring order
                        int s = a_slot + waited; uint32_t pr = a_par;
                        if (s >= P.slots) { s -= P.slots; pr ^= 1u; }
                        ptx::mbar_wait(&a_full[s], pr);
                        ++waited;
                    }
                    ptx::tc_fence_after();
                    uint64_t plane_desc[kSlabAcc];
#pragma unroll
                    for (int a = 0; a < kSlabAcc; ++a) {
                        int s = a_slot + dxi + a;
                        if (s >= P.slots) s -= P.slots;
                        plane_desc[a] = ptx::make_desc_sw128(sa0 + (uint32_t)(s * P.plane_bytes));
                    }
                    for (int dyi = 0; dyi < P.ndy; ++dyi) {
                        const uint64_t yoff = (uint64_t)P.dyrel[dyi] * 64u;          // dy atoms of 1024 B, in the >>4 address field
                        ptx::mbar_wait(&b_full[b_stage], b_par);
                        ptx::tc_fence_after();
                        const uint64_t db = ptx::make_desc_sw128(sb0 + (uint32_t)(b_stage * kSlabBBytes));
                        if (leader) {
#pragma unroll
                            for (int a = 0; a < kSlabAcc; ++a) {
                                const uint64_t da = plane_desc[a] + yoff;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::umma_bf16(d_base + (uint32_t)(a * 64), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc,
                                                   (accumulate | (uint32_t)k) ? 1u : 0u);
                            }
                            ptx::umma_commit(&b_empty[b_stage]);
                        }
                        __syncwarp();
                        accumulate = 1u;
                        if (++b_stage == kSlabBStages) { b_stage = 0; b_par ^= 1u; }
                    }
                    // plane p is last read by (dx index min(p, ndx-1)): release what this dx index finished with
                    const int p_lo = dxi, p_hi = (dxi == P.ndx - 1) ? P.planes - 1 : dxi;
                    for (int p = p_lo; p <= p_hi; ++p) {
                        int s = a_slot + p;
                        if (s >= P.slots) s -= P.slots;
                        if (leader) ptx::umma_commit(&a_empty[s]);
                    }
                    __syncwarp();
                }
                a_slot += P.planes;
                if (a_slot >= P.slots) { a_slot -= P.slots; a_par ^= 1u; }
            }
            if (leader) ptx::umma_commit(&tfull[buf]);
            __syncwarp();
            if (++buf == 2) { buf = 0; buf_par ^= 1u; }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (warps 3..6)
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                   // accumulator row = (y, z) inside the tile, z fastest
        const int yi = row >> 3, zi = row & 7;
        int buf = 0; uint32_t buf_par = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
            int t = tile;
            const int tiz = t % P.tz; t /= P.tz;
            const int tiy = t % P.ty; t /= P.ty;
            const int tix = t % P.tx; const int nb = t / P.tx;
            const int gy = tiy * kSlabTileY + yi, gz = tiz * kSlabTileZ + zi;
            const bool valid_yz = gy < P.yo && gz < P.zo;
            ptx::mbar_wait(&tfull[buf], buf_par);
            ptx::tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kSlabAcc * 64);
#pragma unroll 1
            for (int a = 0; a < kSlabAcc; ++a) {
                uint32_t r0[32], r1[32];
                ptx::tmem_ld_32x32(t_base + (uint32_t)(a * 64), r0);
                ptx::tmem_ld_32x32(t_base + (uint32_t)(a * 64 + 32), r1);
                ptx::tmem_ld_wait();
                const int gx = tix * kSlabAcc + a;
                if (valid_yz && gx < P.xo) {
                    __nv_bfloat16* o = P.y + ((((size_t)nb * P.xo + gx) * P.yo + gy) * P.zo + gz) * P.ldy;
#pragma unroll
                    for (int h = 0; h < 4; ++h) {                       // 16 channels = 32 bytes per step
                        const int chh = h * 16;
                        if (chh >= P.cout) break;
                        uint4 pk[2];
#pragma unroll
                        for (int gg = 0; gg < 2; ++gg) {
                            const int g = h * 2 + gg, ch = g * 8;
                            const uint32_t* r = (g < 4) ? (r0 + g * 8) : (r1 + (g - 4) * 8);
                            float v[8];
                            if (ch < P.cout) {
                                const float4 s0 = __ldg(reinterpret_cast<const float4*>(P.shift + ch));
                                const float4 s1 = __ldg(reinterpret_cast<const float4*>(P.shift + ch + 4));
                                v[0] = __uint_as_float(r[0]) + s0.x; v[1] = __uint_as_float(r[1]) + s0.y; v[2] = __uint_as_float(r[2]) + s0.z;
                                v[3] = __uint_as_float(r[3]) + s0.w; v[4] = __uint_as_float(r[4]) + s1.x; v[5] = __uint_as_float(r[5]) + s1.y;
                                v[6] = __uint_as_float(r[6]) + s1.z; v[7] = __uint_as_float(r[7]) + s1.w;
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = 0.0f;
                            }
                            if (P.relu == 1) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
                            } else if (P.relu == 2) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752f));
                            }
                            pk[gg] = make_uint4(pack_act2(v[0], v[1], P.fp16), pack_act2(v[2], v[3], P.fp16), pack_act2(v[4], v[5], P.fp16), pack_act2(v[6], v[7], P.fp16));
                        }
                        const bool both = chh + 16 <= P.cout;
                        if (both && P.wide_y) ptx::st_global_v8(o + chh, pk[0], pk[1]);
                        else { *reinterpret_cast<uint4*>(o + chh) = pk[0]; if (both) *reinterpret_cast<uint4*>(o + chh + 8) = pk[1]; }
                    }
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty[buf]);
            if (++buf == 2) { buf = 0; buf_par ^= 1u; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------ host
// Decomposes the tap table into {z offsets} x {contiguous x offsets} x {y offsets}; false if it is not such a full box.
static bool slab_plan(const nrpn_conv_desc* d, SlabDev& P) {
    { const char* e = getenv("NRPN_CONV_SLAB"); if (e && e[0] == '0') return false; }   // A/B switch for tests and profiling
    if (d->cin != 64 || d->cout > 64 || d->cout % 8 != 0 || d->stride != 1 || d->n_levels != 1 || d->out_fp32) return false;
    if (d->n_taps < 8 || d->n_taps > NRPN_CONV_MAX_TAPS) return false;
    const nrpn_conv_level& S = d->level[0];
    if (S.res != nullptr) return false;
    if (S.xi < 1 || S.yi < 1 || S.zi < 1 || S.xo < 1 || S.yo < 1 || S.zo < 1) return false;   // input may be larger (packed stem: Z/2 + 1 rows)
    int vz[kSlabMaxAxis], vx[kSlabMaxAxis], vy[kSlabMaxAxis], nz = 0, nx = 0, ny = 0;
    auto insert = [](int* v, int& n, int val) -> bool {
        int i = 0;
        while (i < n && v[i] < val) ++i;
        if (i < n && v[i] == val) return true;
        if (n == kSlabMaxAxis) return false;
        for (int j = n; j > i; --j) v[j] = v[j - 1];
        v[i] = val; ++n;
        return true;
    };
    for (int t = 0; t < d->n_taps; ++t)
        if (!insert(vx, nx, d->tap_off[t][0]) || !insert(vy, ny, d->tap_off[t][1]) || !insert(vz, nz, d->tap_off[t][2])) return false;
    if (nx * ny * nz != d->n_taps) return false;
    if (vx[nx - 1] - vx[0] != nx - 1) return false;                     // slab planes are consecutive in x
    memset(P.widx, 0xFF, sizeof(P.widx));
    for (int t = 0; t < d->n_taps; ++t) {
        int iz = 0, ix = 0, iy = 0;
        while (vz[iz] != d->tap_off[t][2]) ++iz;
        while (vx[ix] != d->tap_off[t][0]) ++ix;
        while (vy[iy] != d->tap_off[t][1]) ++iy;
        if (P.widx[iz][ix][iy] != 0xFF) return false;                   // duplicate tap
        P.widx[iz][ix][iy] = (unsigned char)t;
    }
    const int ys = kSlabTileY + vy[ny - 1] - vy[0];
    if (ys > 256 || ys - kSlabTileY > 255) return false;
    P.n_phases = nz; P.ndx = nx; P.ndy = ny; P.dx0 = vx[0]; P.dy0 = vy[0];
    for (int i = 0; i < nz; ++i) P.dz[i] = (signed char)vz[i];
    for (int i = 0; i < ny; ++i) P.dyrel[i] = (unsigned char)(vy[i] - vy[0]);
    P.plane_bytes = ys * 1024;
    P.planes = kSlabAcc + nx - 1;
    const int avail = kSlabSmemLimit - 1024 - kSlabBStages * kSlabBBytes - 512;
    int slots = avail / P.plane_bytes;
    if (slots > kSlabMaxSlots) slots = kSlabMaxSlots;
    if (slots < P.planes + 1) return false;                            // the ring must hold a whole phase plus one plane in flight
    P.slots = slots;
    return true;
}

bool conv3d_slab_eligible(const nrpn_conv_desc* d) {
    SlabDev P;
    return slab_plan(d, P);
}

int conv3d_slab_launch(const nrpn_conv_desc* d, cudaStream_t st) {
    SlabDev P;
    memset(&P, 0, sizeof(P));
    if (!slab_plan(d, P)) return NRPN_ERR_UNSUPPORTED;
    const nrpn_conv_level& S = d->level[0];
    if (!S.x || !S.y || S.n < 1 || S.xo < 1 || S.yo < 1 || S.zo < 1) return NRPN_ERR_INVALID;
    if (S.ldy < d->cout || S.ldy % 8 != 0) return NRPN_ERR_INVALID;
    EncodeTiledFn encode = get_encode();
    if (!encode) return NRPN_ERR_NO_DEVICE;
    P.n = S.n; P.xo = S.xo; P.yo = S.yo; P.zo = S.zo;
    P.tx = ceil_div(S.xo, kSlabAcc); P.ty = ceil_div(S.yo, kSlabTileY); P.tz = ceil_div(S.zo, kSlabTileZ);
    P.total_tiles = S.n * P.tx * P.ty * P.tz;
    P.fp16 = d->act_fp16 ? 1 : 0;
    P.cout = d->cout; P.relu = d->relu; P.ldy = S.ldy; P.shift = d->shift; P.y = reinterpret_cast<__nv_bfloat16*>(S.y);
    P.wide_y = (((size_t)S.ldy * 2) % 32 == 0 && reinterpret_cast<uintptr_t>(S.y) % 32 == 0) ? 1 : 0;

    SlabMaps maps;
    {   // activations (C, Z, Y, X, N): one box = one x-plane of the slab, {64 ch, 8 z, Ys y}
        cuuint64_t gdim[5] = {64, (cuuint64_t)S.zi, (cuuint64_t)S.yi, (cuuint64_t)S.xi, (cuuint64_t)S.n};
        cuuint64_t gstr[4] = {128, (cuuint64_t)S.zi * 128, (cuuint64_t)S.yi * S.zi * 128, (cuuint64_t)S.xi * S.yi * S.zi * 128};
        cuuint32_t box[5] = {64, (cuuint32_t)kSlabTileZ, (cuuint32_t)(P.plane_bytes / 1024), 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&maps.x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(S.x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return NRPN_ERR_CUDA; }
    }
    {   // weights (Cin, CoutPad = 64, taps)
        cuuint64_t gdim[3] = {64, 64, (cuuint64_t)d->n_taps};
        cuuint64_t gstr[2] = {128, 64 * 128};
        cuuint32_t box[3] = {64, 64, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->w), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return NRPN_ERR_CUDA; }
    }
    const int smem = 1024 + P.slots * P.plane_bytes + kSlabBStages * kSlabBBytes + 512;
    static int smem_set = 0;
    if (smem > smem_set) {
        NRPN_CUDA_TRY(cudaFuncSetAttribute(conv3d_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    const int grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
    NRPN_CUDA_TRY(launch_pdl(conv3d_slab_kernel, dim3(grid), dim3(kSlabThreads), smem, st, P.total_tiles <= 2 * num_sms(), maps, P));
    NRPN_LAUNCH_CHECK();
    return NRPN_OK;
}

}  // namespace nrpn
