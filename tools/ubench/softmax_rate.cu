// Microbenchmark: throughput of the attention softmax inner loop (one thread per query row, 128 scores per tile)
// without any MMA / barrier machinery: TMEM load of S -> row max -> exp2 -> row sum -> bf16 pack -> TMEM store of P.
// Variants: scalar math vs packed f32x2 (FFMA2 / FADD2), and POLY of every 8 exponentials evaluated on the FMA pipe
// (Cody-Waite + degree-3 polynomial, also packed) instead of MUFU.EX2.  Reports clk per 128x128 tile per warp for
// 1 and 2 softmax warps per SM sub-partition (the attention kernel runs 2: two q-tiles in flight per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ucod_dpl_b200/csrc tools/ubench/softmax_rate.cu -o tools/ubench/bin/softmax_rate
#include "common.cuh"
#include <stdlib.h>
using namespace ucod;

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

// 2^x for a pair on the FMA pipe; x <= 8.  (relative error 7.5e-5, well below bf16 resolution)
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& p0, float& p1) {
    x0 = fmaxf(x0, -126.f);
    x1 = fmaxf(x1, -126.f);
    const uint64_t magic = pk2(12582912.f, 12582912.f), nmagic = pk2(-12582912.f, -12582912.f);
    const uint64_t x = pk2(x0, x1);
    const uint64_t t = fadd2(x, magic);
    const uint64_t r = fadd2(t, nmagic);
    float r0, r1; up2(r, r0, r1);
    const uint64_t f = fadd2(x, pk2(-r0, -r1));
    uint64_t p = ffma2(pk2(0.05517121031880379f, 0.05517121031880379f), f, pk2(0.24261027574539185f, 0.24261027574539185f));
    p = ffma2(p, f, pk2(0.6932609677314758f, 0.6932609677314758f));
    p = ffma2(p, f, pk2(0.9999281167984009f, 0.9999281167984009f));
    float q0, q1, t0, t1; up2(p, q0, q1); up2(t, t0, t1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
__device__ __forceinline__ float ex2_poly1(float x) {
    x = fmaxf(x, -126.f);
    const float t = x + 12582912.f;
    const float f = x - (t - 12582912.f);
    float p = fmaf(0.05517121031880379f, f, 0.24261027574539185f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999281167984009f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// POLY: pairs out of every 8 pairs (16 elements) that go to the FMA pipe; PACKED: f32x2 math
template <int POLY, bool PACKED, bool SUM = true, bool MAX = true>
__global__ void __launch_bounds__(256, 1) softmax_rate(int iters, float scale, long long* out, float* sink) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = warp >> 2, quarter = warp & 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tm_s = tm + wg * 128, tm_p = tm + 256 + wg * 64;
    {   // plausible scores in TMEM
        uint32_t init[32];
        for (int c = 0; c < 4; ++c) {
            for (int i = 0; i < 32; ++i) init[i] = __float_as_uint(((lane * 37 + i * 11 + c * 5) % 97) * 0.11f - 5.f);
            tmem_st32(tm_s + lane_off + c * 32, init);
        }
        tmem_wait_st();
    }
    __syncthreads();
    float m_ref = 0.f, l_run = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t u[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tm_s + lane_off + c * 32, reinterpret_cast<uint32_t(&)[32]>(u[32 * c]));
        tmem_wait_ld();
        float mx[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            mx[c] = __uint_as_float(u[32 * c]);
            if constexpr (MAX) {
#pragma unroll
                for (int i = 1; i < 32; i += 2)
                    mx[c] = max3(mx[c], __uint_as_float(u[32 * c + i]), __uint_as_float(u[32 * c + (i + 1 < 32 ? i + 1 : i)]));
            }
        }
        const float ms = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale;
        float alpha = 1.f;
        if (it == 0) m_ref = ms;
        else if (ms > m_ref + 8.f) { alpha = ex2a(m_ref - ms); m_ref = ms; }
        const float neg_m = -m_ref;
        uint32_t pk[64];
        if constexpr (PACKED) {
            const uint64_t sc2 = pk2(scale, scale), nm2 = pk2(neg_m, neg_m);
            uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const uint64_t x = ffma2(pk2(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1])), sc2, nm2);
                float x0, x1, p0, p1;
                up2(x, x0, x1);
                if ((i & 7) < POLY) ex2_poly2(x0, x1, p0, p1);
                else { p0 = ex2a(x0); p1 = ex2a(x1); }
                if constexpr (SUM) ls[i & 3] = fadd2(ls[i & 3], pk2(p0, p1));
                pk[i] = pack_bf16x2(p0, p1);
            }
            float a0, a1, b0, b1;
            up2(fadd2(fadd2(ls[0], ls[1]), fadd2(ls[2], ls[3])), a0, a1);
            (void)b0; (void)b1;
            l_run = l_run * alpha + (a0 + a1);
        } else {
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const float x0 = fmaf(__uint_as_float(u[2 * i]), scale, neg_m);
                const float x1 = fmaf(__uint_as_float(u[2 * i + 1]), scale, neg_m);
                float p0, p1;
                if ((i & 7) < POLY) { p0 = ex2_poly1(x0); p1 = ex2_poly1(x1); }
                else { p0 = ex2a(x0); p1 = ex2a(x1); }
                ls[i & 3] += p0 + p1;
                pk[i] = pack_bf16x2(p0, p1);
            }
            l_run = l_run * alpha + (ls[0] + ls[1]) + (ls[2] + ls[3]);
        }
        tmem_st32(tm_p + lane_off, reinterpret_cast<uint32_t(&)[32]>(pk[0]));
        tmem_st32(tm_p + lane_off + 32, reinterpret_cast<uint32_t(&)[32]>(pk[32]));
        tmem_wait_st();
    }
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = l_run + m_ref;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int POLY, bool PACKED, bool SUM = true, bool MAX = true>
void run(const char* name, long long* d, float* sink) {
    for (int threads : {128, 256}) {
        softmax_rate<POLY, PACKED, SUM, MAX><<<148, threads>>>(2000, 0.18f, d, sink);
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        const double per_tile = (double)h / 2000;
        printf("%-28s %d softmax warps/SMSP: %.0f clk per tile per warp -> %.0f clk per tile per SM (%s)\n", name,
               threads / 128, per_tile, per_tile / (threads / 128), cudaGetErrorString(e));
    }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    float* sink; cudaMalloc(&sink, 148 * 256 * 4);
    run<0, false>("scalar, all MUFU", d, sink);
    run<0, true>("packed, all MUFU", d, sink);
    run<1, true>("packed, 1/8 poly", d, sink);
    run<2, true>("packed, 2/8 poly", d, sink);
    run<3, true>("packed, 3/8 poly", d, sink);
    run<4, true>("packed, 4/8 poly", d, sink);
    run<2, true, false>("packed, 2/8 poly, no sum", d, sink);
    run<3, true, false>("packed, 3/8 poly, no sum", d, sink);
    run<2, true, true, false>("packed, 2/8 poly, no max", d, sink);
    run<2, true, false, false>("packed, 2/8 poly, no sum/max", d, sink);
    run<0, false, false, false>("scalar, all MUFU, no sum/max", d, sink);
    run<2, false>("scalar, 2/8 poly", d, sink);
    run<4, false>("scalar, 4/8 poly", d, sink);
    return 0;
}
