// Microbenchmark: issue / completion rate of tcgen05.mma (kind::f16) for the shapes the attention/GEMM kernels use.
// One CTA per SM, one issuing thread, operands are whatever is in smem/TMEM (values irrelevant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ucod_dpl_b200/csrc tools/ubench/mma_rate.cu -o tools/ubench/bin/mma_rate
// Round 2 adds: K-major B for the TS form, alternating accumulators (is the N = 64 P.V chain latency- or
// throughput-bound?), N = 128 TS, the attention tile mix with ping-pong S / O accumulators, and cta_group::2 pairs
// (M = 256 over two SMs) for both MMA forms.
#include "common.cuh"
#include <stdlib.h>
using namespace ucod;

__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo = 16384) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo >> 4) << 16;
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
__device__ __forceinline__ void mma_ss2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts2(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

enum {
    M_SS256 = 0, M_SS128, M_SS64, M_TS64_MN, M_MIX, M_TS64_K, M_TS64_MN_ALT, M_TS128_MN, M_SS64_ALT, M_MIX_PP,
    M_TS64_MN_4ACC, M_COUNT
};

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
        const uint32_t id64t = id64 | (1u << 16), id128t = id128 | (1u << 16);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            switch (mode) {
            case M_SS256: for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id256, 1); break;
            case M_SS128: for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1); break;
            case M_SS64: for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id64, 1); break;
            case M_TS64_MN: for (int k = 0; k < 4; ++k) mma_ts(tm + 192, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1); break;
            case M_MIX:
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1);
                for (int k = 0; k < 8; ++k) mma_ts(tm + 192, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1);
                break;
            case M_TS64_K: for (int k = 0; k < 4; ++k) mma_ts(tm + 192, tm + 128 + k * 8, umma_desc_kmajor_sw128(b + k * 32), id64, 1); break;
            case M_TS64_MN_ALT: for (int k = 0; k < 4; ++k) mma_ts(tm + 192 + (k & 1) * 64, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1); break;
            case M_TS128_MN: for (int k = 0; k < 4; ++k) mma_ts(tm + 256, tm + 128 + k * 8, desc_mn(b + k * 2048), id128t, 1); break;
            case M_SS64_ALT: for (int k = 0; k < 4; ++k) umma_bf16_ss(tm + (k & 1) * 64, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id64, 1); break;
            case M_MIX_PP: {
                // two q-tiles sharing K/V: S0, PV1, S1, PV0 (accumulators S0 0, S1 128, O0 256, O1 320; P aliases S)
                const int pp = i & 1;
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tm + pp * 128, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1);
                for (int k = 0; k < 8; ++k) mma_ts(tm + 256 + (pp ^ 1) * 64, tm + (pp ^ 1) * 128 + k * 8, desc_mn(b + k * 2048), id64t, 1);
                break;
            }
            case M_TS64_MN_4ACC: for (int k = 0; k < 4; ++k) mma_ts(tm + 192 + k * 64, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1); break;
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

// two co-resident CTAs per SM (256 TMEM columns each, like the attention kernel): each issues the attention tile mix
// mode 0: 4x SS N=128 then 8x TS N=64 ; mode 1: SS only ; mode 2: TS only ; mode 3: interleaved S,T,T,S,T,T,...
__global__ void __launch_bounds__(128, 2) mma_rate_2cta(int mode, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, 256); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
        const uint32_t id128 = umma_idesc_bf16(128, 128), id64t = umma_idesc_bf16(128, 64) | (1u << 16);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (mode == 0 || mode == 1)
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1);
            if (mode == 0 || mode == 2)
                for (int k = 0; k < 8; ++k) mma_ts(tm + 192, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1);
            if (mode == 3)
                for (int k = 0; k < 4; ++k) {
                    umma_bf16_ss(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1);
                    mma_ts(tm + 192, tm + 128 + 2 * k * 8, desc_mn(b + 2 * k * 2048), id64t, 1);
                    mma_ts(tm + 192, tm + 128 + (2 * k + 1) * 8, desc_mn(b + (2 * k + 1) * 2048), id64t, 1);
                }
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

// ---- CTA pairs (cta_group::2): the leader issues M = 256 MMAs over both SMs ----
__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// mode 0: SS M=256 N=128 ; 1: SS M=256 N=256 ; 2: TS M=256 N=64 MN-B ; 3: TS M=256 N=128 MN-B ; 4: mix 4x SS(256x128) + 8x TS(256x64)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_rate_pair(int mode, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t rank = cl_rank();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    cl_sync();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    cl_sync();
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
        // M = 256: N is the full N of the pair; each CTA supplies N/2 rows of B
        const uint32_t id128 = umma_idesc_bf16(256, 128), id256 = umma_idesc_bf16(256, 256);
        const uint32_t id64t = umma_idesc_bf16(256, 64) | (1u << 16), id128t = umma_idesc_bf16(256, 128) | (1u << 16);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            switch (mode) {
            case 0: for (int k = 0; k < 4; ++k) mma_ss2(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1); break;
            case 1: for (int k = 0; k < 4; ++k) mma_ss2(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id256, 1); break;
            case 2: for (int k = 0; k < 4; ++k) mma_ts2(tm + 256, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1); break;
            case 3: for (int k = 0; k < 4; ++k) mma_ts2(tm + 256, tm + 128 + k * 8, desc_mn(b + k * 2048), id128t, 1); break;
            default:
                for (int k = 0; k < 4; ++k) mma_ss2(tm, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), id128, 1);
                for (int k = 0; k < 8; ++k) mma_ts2(tm + 256, tm + 128 + k * 8, desc_mn(b + k * 2048), id64t, 1);
                break;
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    cl_sync();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(mma_rate_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2000;
    const char* names[] = {"SS N=256 (x4)", "SS N=128 (x4)", "SS N=64 (x4)", "TS N=64 MN-B (x4)", "4xSS128 + 8xTS64",
                           "TS N=64 K-major B", "TS N=64 MN-B 2 acc", "TS N=128 MN-B", "SS N=64 2 acc",
                           "pingpong 4xSS128+8xTS64", "TS N=64 MN-B 4 acc"};
    const int per[] = {4, 4, 4, 4, 12, 4, 4, 4, 4, 12, 4};
    for (int grid : {1, 148}) for (int mode = 0; mode < M_COUNT; ++mode) {
        mma_rate<<<grid, 128, 100 * 1024>>>(mode, iters, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("grid %3d %-26s issue %.1f clk/mma  complete %.1f clk/mma  (%s)\n", grid, names[mode],
               (double)h[0] / iters / per[mode], (double)h[1] / iters / per[mode], cudaGetErrorString(e));
    }
    cudaFuncSetAttribute(mma_rate_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char* cnames[] = {"attention mix 4SS+8TS", "SS N=128 only (x4)", "TS N=64 only (x8)", "interleaved S,T,T"};
    const int cper[] = {12, 4, 8, 12};
    for (int grid : {148, 296}) for (int mode = 0; mode < 4; ++mode) {
        mma_rate_2cta<<<grid, 128, 100 * 1024>>>(mode, iters, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%d CTA/SM %-24s complete %.1f clk/mma per CTA -> %.0f clk per tile per SM (%s)\n", grid / 148, cnames[mode],
               (double)h[1] / iters / cper[mode], (double)h[1] / iters / (grid / 148) * (mode == 0 || mode == 3 ? 1.0 : 0.0),
               cudaGetErrorString(e));
    }
    const char* pnames[] = {"pair SS 256x128", "pair SS 256x256", "pair TS 256x64 MN-B", "pair TS 256x128 MN-B",
                            "pair 4xSS(256x128)+8xTS(256x64)"};
    const int pper[] = {4, 4, 4, 4, 12};
    for (int grid : {2, 148}) for (int mode = 0; mode < 5; ++mode) {
        mma_rate_pair<<<grid, 128, 100 * 1024>>>(mode, iters, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("grid %3d %-34s issue %.1f clk/mma  complete %.1f clk/mma  (%s)\n", grid, pnames[mode],
               (double)h[0] / iters / pper[mode], (double)h[1] / iters / pper[mode], cudaGetErrorString(e));
    }
    return 0;
}
