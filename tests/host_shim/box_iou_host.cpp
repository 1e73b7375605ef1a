// Host build of nerf_rpn_b200/csrc/box_iou.cuh (TEST ONLY): the CUDA device functions are compiled as
// plain C++ (with -ffp-contract=off, so __f*_rn == separately rounded ops) to check their logic against
// the oracle on GPU-less boxes. The product never uses this file.
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static int g_shim_iou_mode = 0;
#define NRPN_IOU_MODE g_shim_iou_mode
extern "C" void shim_set_iou_mode(int m) { g_shim_iou_mode = m; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
using std::isfinite;
#define NRPN_HOST_SHIM 1
namespace nrpn { template <typename T> inline T ceil_div(T a, T b) { return (a + b - 1) / b; } }
#define NRPN_SKIP_COMMON 1
#include "../../nerf_rpn_b200/csrc/box_iou.cuh"

extern "C" void shim_iou_pairs(const float* a, const float* b, int n, int box_dim, int cull, float* out) {
    for (int i = 0; i < n; ++i) {
        if (box_dim == 7) {
            nrpn::ObbPrep pa, pb;
            nrpn::obb_prepare(a + (size_t)i * 7, pa);
            nrpn::obb_prepare(b + (size_t)i * 7, pb);
            out[i] = nrpn::iou3d_obb(pa, pb, cull != 0);
        } else out[i] = nrpn::iou3d_aabb(a + (size_t)i * 6, b + (size_t)i * 6);
    }
}

#include "../../nerf_rpn_b200/csrc/rpn_decode.cuh"
extern "C" void shim_decode(const float* anchors, const float* deltas, int n, int rotated, float* out) {
    for (int i = 0; i < n; ++i) {
        if (rotated) nrpn::decode_obb(anchors + (size_t)i * 6, deltas + (size_t)i * 8, out + (size_t)i * 7);
        else nrpn::decode_aabb(anchors + (size_t)i * 6, deltas + (size_t)i * 6, out + (size_t)i * 6);
    }
}
extern "C" void shim_sigmoid(const float* x, int n, float* out) { for (int i = 0; i < n; ++i) out[i] = nrpn::sigmoid_ref(x[i]); }
