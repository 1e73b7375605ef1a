// Microbenchmark: issue rate of tcgen05.mma (kind::f16, cta_group::1, both operands from shared memory, SWIZZLE_128B K-major)
// as a function of the instruction shape.  One CTA per SM, one issuing thread, operands never change (no TMA in the loop).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I nerf_rpn_b200/csrc tools/micro/mma_floor.cu -o build/mma_floor
#include <cstdio>
#include "tcgen05.cuh"
using namespace nrpn;

template <int M, int N>
__global__ void __launch_bounds__(128, 1) mma_floor_kernel(long long* out, int iters, int same_a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (4 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    ptx::fence_proxy_async();
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(M, N);
        const uint32_t sa = ptx::smem_u32(smem), sb = sa + 4 * 16384;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {                     // 4 A tiles (e.g. 4 accumulators x 1 tap), 4 k-steps each
                const uint64_t da = ptx::make_desc_sw128(sa + (same_a ? 0 : a * 16384));
                const uint64_t db = ptx::make_desc_sw128(sb);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::umma_bf16(tmem + (uint32_t)((a * N) % 512), da + 2 * k, db + 2 * k, idesc, 1u);
            }
        }
        ptx::umma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int M, int N>
static void run(const char* name, int same_a) {
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    const int smem = 4 * 16384 + 32768 + 1024;
    cudaFuncSetAttribute(mma_floor_kernel<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int grid : {1, 148}) {
        mma_floor_kernel<M, N><<<grid, 128, smem>>>(d, iters, same_a);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double per = (double)mx / (iters * 16.0);
        printf("%-28s grid %3d  %7.1f clk/MMA  %6.1f%% of 4096 FMA/clk/SM  (%s)\n", name, grid, per,
               100.0 * (double)M * N * 16 / per / 4096.0, cudaGetErrorString(e));
    }
    cudaFree(d);
}

int main() {
    run<128, 256>("M128 N256 distinct A", 0);
    run<128, 128>("M128 N128 distinct A", 0);
    run<128, 64>("M128 N64  distinct A", 0);
    run<128, 64>("M128 N64  same A", 1);
    run<128, 32>("M128 N32  distinct A", 0);
    run<64, 256>("M64  N256 distinct A", 0);
    run<64, 128>("M64  N128 distinct A", 0);
    run<64, 64>("M64  N64  distinct A", 0);
    return 0;
}
