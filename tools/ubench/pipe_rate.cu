// Microbenchmark: issue throughput of the instruction kinds the attention softmax loop is made of (per SM
// sub-partition, clk per warp instruction), alone and in the pairings that matter (MUFU + FMA-pipe work).
// 8 independent dependency chains per thread, 1 / 2 / 4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ucod_dpl_b200/csrc tools/ubench/pipe_rate.cu -o tools/ubench/bin/pipe_rate
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define CH 8
#define INNER 16

enum { OP_FFMA = 0, OP_FFMA2, OP_FADD, OP_FADD2, OP_MUFU, OP_FMNMX3, OP_F2FP, OP_LEA, OP_MUFU_FFMA, OP_MUFU_FFMA2,
       OP_MUFU_F2FP_FFMA2, OP_FMNMX, OP_FMUL2, OP_COUNT };

template <int OP>
__global__ void __launch_bounds__(512, 1) pipe_rate(int iters, long long* out, float* sink, float seed) {
    float a[CH], b[CH];
    uint64_t p[CH];
    uint32_t u[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        a[i] = seed + i + threadIdx.x * 1e-3f;
        b[i] = seed * 0.5f + i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
        u[i] = threadIdx.x + i;
    }
    const float c0 = seed * 1.0001f, c1 = seed * 0.25f;
    uint64_t pc0, pc1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pc0) : "f"(c0), "f"(c0));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pc1) : "f"(c1), "f"(c1));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (OP == OP_FFMA) a[i] = fmaf(a[i], c0, c1);
                if (OP == OP_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc0), "l"(pc1));
                if (OP == OP_FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc0));
                if (OP == OP_FADD) a[i] = a[i] + c1;
                if (OP == OP_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc1));
                if (OP == OP_MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                if (OP == OP_FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c1));
                if (OP == OP_FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
                if (OP == OP_F2FP) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
                if (OP == OP_LEA) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 23; add.u32 %0, t, %1; }" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
                if (OP == OP_MUFU_FFMA) {
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                    b[i] = fmaf(b[i], c0, c1);
                }
                if (OP == OP_MUFU_FFMA2) {
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc0), "l"(pc1));
                }
                if (OP == OP_MUFU_F2FP_FFMA2) {
                    // the per-pair mix of the all-MUFU softmax: 2 MUFU + 1 FFMA2 + 1 FADD2 + 1 F2FP + 1 FMNMX3
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc0), "l"(pc1));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[(i + 1) % CH]) : "l"(pc1));
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i]));
                    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[(i + 3) % CH]) : "f"(b[i]), "f"(c1));
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        s += a[i] + b[i] + lo + hi + __uint_as_float(u[i]);
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, long long* d, float* sink) {
    printf("%-34s", name);
    for (int threads : {128, 256, 512}) {
        const int iters = 200;
        pipe_rate<OP><<<148, threads>>>(iters, d, sink, 1.0f);
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double n = (double)iters * INNER * CH * per_iter * (threads / 128);  // warp instructions per sub-partition
        printf("  %dw/SMSP %.2f clk/instr", threads / 128, (double)h / n);
    }
    printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    float* sink; cudaMalloc(&sink, 148 * 512 * 4);
    run<OP_FFMA>("FFMA", 1, d, sink);
    run<OP_FFMA2>("FFMA2", 1, d, sink);
    run<OP_FMUL2>("FMUL2", 1, d, sink);
    run<OP_FADD>("FADD", 1, d, sink);
    run<OP_FADD2>("FADD2", 1, d, sink);
    run<OP_MUFU>("MUFU.EX2", 1, d, sink);
    run<OP_FMNMX3>("FMNMX3", 1, d, sink);
    run<OP_FMNMX>("FMNMX", 1, d, sink);
    run<OP_F2FP>("F2FP.BF16 pack", 1, d, sink);
    run<OP_LEA>("SHL+ADD (LEA)", 1, d, sink);
    run<OP_MUFU_FFMA>("MUFU + FFMA (per pair)", 2, d, sink);
    run<OP_MUFU_FFMA2>("MUFU + FFMA2 (per pair)", 2, d, sink);
    run<OP_MUFU_F2FP_FFMA2>("softmax mix (6 instrs)", 6, d, sink);
    return 0;
}
