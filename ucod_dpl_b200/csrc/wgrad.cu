// Weight gradient of the decoder's 1x1 `decoupling` conv on the tensor cores (first-stage training step,
// engine/runner/loop_UCOD_DPL.py:148-184 -> autograd of models/modules/DBA.py:35):
//
//     dW[c, k] = sum_t dD[t, c] * X[t, k]          c < 128 output channels, k < dim input features, t < T tokens
//
// A contraction over the TOKEN index of two token-major tensors: neither operand has the reduction index contiguous.
//   * dD is small (T x 128): `decoder_bwd_pack_kernel` (decoder.cu) fuses the last step of the backward
//     (dD = A - r^2 t E) with a transpose to dDt bf16 [128, Tpad] (K-major A operand) and writes per-block column
//     sums for db into the buffers `wgrad_plan` lays out.
//   * X (the cached backbone keys, T x dim bf16, the only HBM-sized read) is consumed in place as an MN-MAJOR B operand:
//     a TMA box of 64 tokens x 64 features lands as 64 rows of 128 bytes, which is exactly the tcgen05 MN-major SW128
//     layout (the same trick the attention kernel uses for V).
// Split-K over the tokens: CTA (n, s) accumulates its token range for feature tile n (256 wide) in TMEM (128 x 256 fp32)
// and writes a partial tile; `wgrad_reduce_kernel` adds the partials in a fixed order, so the result is bit-reproducible
// (round 1: CUDA-core FFMA kernel with float atomics, 211 us for 16 images).
#include "wgrad.cuh"

#include "prof.cuh"

namespace ucod {

namespace {

constexpr int WG_BM = 128, WG_BN = 256, WG_BK = 64, WG_STAGES = 4;
constexpr int WG_A_BYTES = WG_BM * WG_BK * 2;            // 16 KB: dDt tile [128 x 64], K-major
constexpr int WG_BBOX_BYTES = WG_BK * 64 * 2;            // 8 KB: X box [64 tokens x 64 features]
constexpr int WG_B_BYTES = (WG_BN / 64) * WG_BBOX_BYTES;  // 32 KB
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 256 + 1024;
constexpr int WG_THREADS = 192;

__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
    wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         float* __restrict__ partial, int dim, int n_kblocks, int kb_per_split) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + WG_STAGES * WG_A_BYTES;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* bar_empty = bar_full + WG_STAGES;
    uint64_t* bar_done = bar_empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * WG_BN, split = blockIdx.y;
    const int kb0 = split * kb_per_split;
    const int kb1 = min(n_kblocks, kb0 + kb_per_split);  // host guarantees kb0 < kb1

    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, WG_BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait_parked(&bar_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&bar_full[stage], WG_STAGE_BYTES);
                tma_load_2d(sA + stage * WG_A_BYTES, &tm_a, &bar_full[stage], kb * WG_BK, 0);
#pragma unroll
                for (int c = 0; c < WG_BN / 64; ++c)
                    tma_load_2d(sB + stage * WG_B_BYTES + c * WG_BBOX_BYTES, &tm_b, &bar_full[stage], n0 + c * 64,
                                kb * WG_BK);
                if (++stage == WG_STAGES) stage = 0, phase ^= 1;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            constexpr uint32_t idesc = umma_idesc_bf16(WG_BM, WG_BN) | (1u << 16);  // B operand MN-major
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait_parked(&bar_full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + stage * WG_A_BYTES);
                const uint32_t b_addr = smem_u32(sB + stage * WG_B_BYTES);
#pragma unroll
                for (int k = 0; k < WG_BK / 16; ++k)
                    umma_bf16_ss(tmem_acc, umma_desc_kmajor_sw128(a_addr + k * 32),
                                 desc_mnmajor_sw128(b_addr + k * 2048, WG_BBOX_BYTES), idesc, (kb > kb0) || (k > 0));
                umma_commit(&bar_empty[stage]);
                if (++stage == WG_STAGES) stage = 0, phase ^= 1;
            }
            umma_commit(bar_done);
        }
    } else {
        // ===== epilogue: TMEM -> this split's partial tile =====
        const int q = warp & 3;
        const int c = q * 32 + lane;  // output channel == TMEM lane
        mbar_wait(bar_done, 0);
        tc_fence_after();
        float* dst = partial + ((size_t)split * WG_BM + c) * dim + n0;
#pragma unroll 1
        for (int ch = 0; ch < WG_BN / 32; ++ch) {
            uint32_t v[32];
            tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + ch * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(dst + ch * 32)[i] =
                    make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                __uint_as_float(v[4 * i + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, WG_BN);
}

// dW = sum of the split partials, db = sum of the block column sums — both in a fixed order
__global__ void __launch_bounds__(256)
    wgrad_reduce_kernel(const float* __restrict__ partial, int n_splits, float* __restrict__ dW, int n_elem,
                        const float* __restrict__ bpart, int n_blocks, float* __restrict__ db) {
    // one element per thread (4x the CTAs of a float4 mapping: the sum over the splits is a serial chain per element,
    // so the kernel is latency-bound and wants threads, not wide loads); 8 independent loads in flight per thread
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_elem) {
        float s = 0.f;
        int k = 0;
        for (; k + 8 <= n_splits; k += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = partial[(size_t)(k + u) * n_elem + i];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; k < n_splits; ++k) s += partial[(size_t)k * n_elem + i];
        dW[i] = s;
    }
    // db: the column sums of the 64-row blocks.  One CTA, two threads per channel, 32 independent loads in flight per
    // thread (a plain serial loop over the ~340 blocks was a 20 us dependent-latency chain and set the kernel's time);
    // fixed order: even blocks, odd blocks, then even + odd
    if (blockIdx.x == 0) {
        __shared__ float s_half[128];
        const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
        float s = 0.f;
        int k = half;
        for (; k + 62 < n_blocks; k += 64) {
            float v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] = bpart[(size_t)(k + 2 * u) * 128 + c];
#pragma unroll
            for (int u = 0; u < 32; ++u) s += v[u];
        }
        for (; k < n_blocks; k += 2) s += bpart[(size_t)k * 128 + c];
        if (half == 1) s_half[c] = s;
        __syncthreads();
        if (half == 0) db[c] = s + s_half[c];
    }
}

int wg_splits(int n_kblocks, int n_tiles, int* kb_per_split) {
    int target = (2 * device_sm_count()) / n_tiles;  // ~2 CTAs' worth of work items per SM-row of tiles
    if (target > device_sm_count() / n_tiles) target = device_sm_count() / n_tiles;  // one wave
    if (target < 1) target = 1;
    int kps = ceil_div(n_kblocks, target);
    if (kps < 2 && n_kblocks >= 2) kps = 2;
    *kb_per_split = kps;
    return ceil_div(n_kblocks, kps);
}

}  // namespace

size_t wgrad_workspace_bytes(int T, int dim) {
    const int n_kblocks = ceil_div(T, WG_BK);
    int kps = 1;
    const int splits = wg_splits(n_kblocks, dim / WG_BN, &kps);
    const size_t t_pad = (size_t)n_kblocks * WG_BK;
    return 128 * t_pad * 2 + (size_t)splits * 128 * dim * 4 + (size_t)n_kblocks * 128 * 4 + 4096;
}

int wgrad_plan(int T, int dim, void* workspace, size_t ws_bytes, WgradPlan* plan) {
    UCOD_REQUIRE(dim % WG_BN == 0, "wgrad: dim must be a multiple of %d", WG_BN);
    UCOD_REQUIRE(ws_bytes >= wgrad_workspace_bytes(T, dim), "wgrad: workspace too small");
    UCOD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "wgrad: workspace must be 256-byte aligned");
    plan->n_kblocks = ceil_div(T, WG_BK);
    plan->t_pad = plan->n_kblocks * WG_BK;
    plan->splits = wg_splits(plan->n_kblocks, dim / WG_BN, &plan->kb_per_split);
    uint8_t* p = static_cast<uint8_t*>(workspace);
    plan->dDt = reinterpret_cast<__nv_bfloat16*>(p);
    p += (128 * (size_t)plan->t_pad * 2 + 255) / 256 * 256;
    plan->partial = reinterpret_cast<float*>(p);
    p += (size_t)plan->splits * 128 * dim * 4;
    plan->bpart = reinterpret_cast<float*>(p);
    return 0;
}

int wgrad_contract(const WgradPlan& plan, const void* keys_bf16, int T, int dim, float* dW, float* db,
                   cudaStream_t stream) {
    CUtensorMap ta, tb;
    if (int rc = make_tmap_2d_bf16(&ta, plan.dDt, 128, (uint64_t)plan.t_pad, (uint64_t)plan.t_pad, 128, 64)) return rc;
    if (int rc = make_tmap_2d_bf16(&tb, keys_bf16, (uint64_t)T, (uint64_t)dim, (uint64_t)dim, 64, 64)) return rc;
    static bool configured = false;
    if (!configured) {
        UCOD_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        configured = true;
    }
    {
        ProfScope ps(KC_GEMM, stream, 2.0 * 128 * (double)dim * T);
        wgrad_tcgen05_kernel<<<dim3(dim / WG_BN, plan.splits), WG_THREADS, WG_SMEM, stream>>>(
                ta, tb, plan.partial, dim, plan.n_kblocks, plan.kb_per_split);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    {
        const int n_elem = 128 * dim;
        ProfScope ps(KC_DECODER, stream, (double)plan.splits * n_elem * 4);
        wgrad_reduce_kernel<<<ceil_div(n_elem, 256), 256, 0, stream>>>(plan.partial, plan.splits, dW, n_elem,
                                                                           plan.bpart, plan.n_kblocks, db);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
