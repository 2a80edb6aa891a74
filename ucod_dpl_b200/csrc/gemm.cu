// Persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   D[M,N] = A[M,K] * W[N,K]^T      A, W: bf16, K contiguous ("TN");  accumulate fp32 in TMEM.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM allocator), warps 2..5 = epilogue.
// Tile 128 x BN x 64, STAGES-deep smem ring (mbarrier full/empty), two TMEM accumulator stages so the
// epilogue of tile i overlaps the main loop of tile i+1. One CTA per SM, static tile striding with the
// N index fastest so that CTAs running concurrently share the same A row-block through L2.
//
// Replaces (reference, all library-dispatched): HF Dinov2/ViT `nn.Linear` query/key/value/dense/fc1/fc2
// (transformers modeling_dinov2.py:153-235,348-387), patch-embed Conv2d (:38-117), the decoder's
// `decoupling` 1x1 conv (models/modules/DBA.py:13,35) and the CORAL CSF projections (models/modules/mlp.py:116-148).
#include "gemm.cuh"
#include "prof.cuh"

namespace ucod {

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128;
    static constexpr int BK = 64;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
    static constexpr int TMEM_COLS = 2 * BN;                                    // two accumulator stages
    static constexpr int THREADS = 192;
};

// ------------------------------------------------------------------------------------------------
// Epilogue: one thread owns one accumulator row; called per 32-column chunk.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load32(const float* p, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i + 0] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void store32_bf16(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 t;
        t.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
        t.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        t.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
        t.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        reinterpret_cast<uint4*>(dst)[i] = t;
    }
}
__device__ __forceinline__ void store32_f32(float* dst, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const GemmEpi& ep, int m, int n0, int N, float (&v)[32]) {
    if (ep.bias != nullptr) {
        float b[32];
        load32(ep.bias + n0, b);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += b[i];
    }
    if constexpr (MODE == EPI_BIAS_BF16) {
        store32_bf16(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)m * ep.ld_out + n0, v);
    } else if constexpr (MODE == EPI_BIAS_GELU_BF16) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        store32_bf16(reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)m * ep.ld_out + n0, v);
    } else if constexpr (MODE == EPI_BIAS_F32) {
        store32_f32(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ld_out + n0, v);
    } else if constexpr (MODE == EPI_RESID_F32) {
        float* x = reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ld_out + n0;
        if (ep.scale != nullptr) {
            float s[32];
            load32(ep.scale + n0, s);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= s[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 t = reinterpret_cast<const float4*>(x)[i];
            v[4 * i] += t.x, v[4 * i + 1] += t.y, v[4 * i + 2] += t.z, v[4 * i + 3] += t.w;
        }
        store32_f32(x, v);
    } else if constexpr (MODE == EPI_PATCH) {
        const int P = ep.tokens;
        const int b = m / P, p = m - b * P;
        float pe[32];
        load32(ep.pos + (size_t)(1 + p) * N + n0, pe);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += pe[i];
        store32_f32(reinterpret_cast<float*>(ep.out) + ((size_t)b * (P + 1) + 1 + p) * ep.ld_out + n0, v);
    } else if constexpr (MODE == EPI_KEYS) {
        const int T = ep.tokens;
        const int b = m / T, t = m - b * T;
        if (t < ep.skip) return;
        const size_t row = (size_t)b * (T - ep.skip) + (t - ep.skip);
        if (ep.out != nullptr) store32_f32(reinterpret_cast<float*>(ep.out) + row * ep.ld_out + n0, v);
        if (ep.out2 != nullptr) store32_bf16(reinterpret_cast<__nv_bfloat16*>(ep.out2) + row * ep.ld_out + n0, v);
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int MODE>
__global__ void __launch_bounds__(192, 1)
    gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                             int M, int N, int K, const GemmEpi ep) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* bar_empty = bar_full + STAGES;
    uint64_t* bar_tfull = bar_empty + STAGES;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = (M + Cfg::BM - 1) / Cfg::BM;
    const int n_tiles = N / BN;
    const int total_tiles = m_tiles * n_tiles;
    const int k_blocks = (K + Cfg::BK - 1) / Cfg::BK;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&tmap_a);
            tma_prefetch_desc(&tmap_b);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * Cfg::BM;
                const int n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bar_full[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(sA + stage * Cfg::A_BYTES, &tmap_a, &bar_full[stage], kb * Cfg::BK, m0);
                    tma_load_2d(sB + stage * Cfg::B_BYTES, &tmap_b, &bar_full[stage], kb * Cfg::BK, n0);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(Cfg::BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&bar_tempty[as], aphase ^ 1);  // epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&bar_full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < Cfg::BK / 16; ++k) {
                        umma_bf16_ss(tmem_d, umma_desc_kmajor_sw128(a_addr + k * 32),
                                     umma_desc_kmajor_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                    }
                    umma_commit(&bar_empty[stage]);  // frees the smem slot when these MMAs retire
                    if (kb == k_blocks - 1) umma_commit(&bar_tfull[as]);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else {
        // ===================== Epilogue warps (TMEM -> regs -> global) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may touch
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / n_tiles) * Cfg::BM;
            const int n0 = (tile % n_tiles) * BN;
            mbar_wait(&bar_tfull[as], aphase);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                tmem_wait_ld();
                if (m < M) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    epilogue_chunk<MODE>(ep, m, n0 + c * 32, N, v);
                }
            }
            tc_fence_before();
            mbar_arrive(&bar_tempty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Host launcher
// ------------------------------------------------------------------------------------------------
template <int BN, int MODE>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpi& ep,
                       cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    auto kern = gemm_bf16_tcgen05_kernel<BN, MODE>;
    static bool configured = false;
    if (!configured) {
        UCOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int tiles = ceil_div(M, Cfg::BM) * (N / BN);
    const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
    {
        ProfScope ps(KC_GEMM, stream, 2.0 * M * N * K);
        kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, M, N, K, ep);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int BN>
static int launch_mode(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpi& ep,
                       cudaStream_t s) {
    switch (ep.mode) {
        case EPI_BIAS_BF16: return launch_inst<BN, EPI_BIAS_BF16>(ta, tb, M, N, K, ep, s);
        case EPI_BIAS_GELU_BF16: return launch_inst<BN, EPI_BIAS_GELU_BF16>(ta, tb, M, N, K, ep, s);
        case EPI_RESID_F32: return launch_inst<BN, EPI_RESID_F32>(ta, tb, M, N, K, ep, s);
        case EPI_PATCH: return launch_inst<BN, EPI_PATCH>(ta, tb, M, N, K, ep, s);
        case EPI_BIAS_F32: return launch_inst<BN, EPI_BIAS_F32>(ta, tb, M, N, K, ep, s);
        case EPI_KEYS: return launch_inst<BN, EPI_KEYS>(ta, tb, M, N, K, ep, s);
        default: set_last_error("launch_gemm_bf16: unknown epilogue mode %d", ep.mode); return 1;
    }
}

int launch_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpi& ep,
                     cudaStream_t stream) {
    UCOD_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    UCOD_REQUIRE(N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
    UCOD_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && K % 8 == 0, "gemm: lda/ldw/K must be multiples of 8");
    const int BN = (N % 256 == 0) ? 256 : 128;
    CUtensorMap ta, tb;
    if (int rc = make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64)) return rc;
    if (int rc = make_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)BN, 64)) return rc;
    return BN == 256 ? launch_mode<256>(ta, tb, M, N, K, ep, stream) : launch_mode<128>(ta, tb, M, N, K, ep, stream);
}

}  // namespace ucod
