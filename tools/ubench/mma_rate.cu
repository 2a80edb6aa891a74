// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16) for the shapes the attention/GEMM kernels use.
// One CTA per SM (optionally two), one issuing thread, operands are whatever is in smem/TMEM (values irrelevant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ucod_dpl_b200/csrc tools/ubench/mma_rate.cu -o /tmp/mma_rate
#include "common.cuh"
#include <stdlib.h>
using namespace ucod;

__device__ __forceinline__ uint64_t desc_mn(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3FFF);
    d |= (uint64_t)(16384 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// mode 0: SS N=256 ; 1: SS N=128 ; 2: SS N=64 ; 3: TS N=64 (B MN-major) ; 4: alternate 4x SS N=128 + 8x TS N=64
__global__ void __launch_bounds__(128, 1) mma_rate(int mode, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
        const uint32_t id256 = umma_idesc_bf16(128, 256), id128 = umma_idesc_bf16(128, 128), id64 = umma_idesc_bf16(128, 64);
        const uint32_t id64t = id64 | (1u << 16);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (mode == 0) { for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id256, 1); }
            else if (mode == 1) { for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1); }
            else if (mode == 2) { for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id64, 1); }
            else if (mode == 3) { for (int k = 0; k < 4; ++k) mma_ts(tm + 192, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1); }
            else {
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1);
                for (int k = 0; k < 8; ++k) mma_ts(tm + 192, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1);
            }
        }
        long long t1 = clock64();   // issue time
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();   // completion time
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2000;
    const char* names[] = {"SS N=256 (x4)", "SS N=128 (x4)", "SS N=64 (x4)", "TS N=64 MN-B (x4)", "4xSS128 + 8xTS64"};
    const int per[] = {4, 4, 4, 4, 12};
    for (int grid : {1, 148}) for (int mode = 0; mode < 5; ++mode) {
        mma_rate<<<grid, 128, 100 * 1024>>>(mode, iters, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("grid %3d %-20s issue %.1f clk/mma  complete %.1f clk/mma  (%s)\n", grid, names[mode],
               (double)h[0] / iters / per[mode], (double)h[1] / iters / per[mode], cudaGetErrorString(e));
    }
    return 0;
}
