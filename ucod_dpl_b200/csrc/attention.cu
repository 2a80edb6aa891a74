// Fused multi-head attention forward (non-causal, head_dim 64) on tcgen05 / TMEM.
//
//   ctx[b, t, h*64:(h+1)*64] = softmax_j(q[b,h,t,:] . k[b,h,j,:] * scale) @ v[b,h,j,:]
//
// One CTA = one (batch*head, 128-query tile).  Warp 0: TMA producer (Q once; K_j and V^T_j double
// buffered).  Warp 1: single-thread MMA issuer: S = Q K_j^T (128x128x64) into TMEM cols [0,128),
// O += P_j V_j (128x64x128) into TMEM cols [128,192).  Warps 2..5: one thread per query row —
// online softmax straight out of TMEM (no cross-thread shuffles), P written as bf16 into a
// 128B-swizzled smem tile that the second MMA consumes, O rescaled in TMEM only when the row max moved.
// S_{j+1} is issued before P_j V_j so the tensor pipe works while the softmax warps run; two CTAs fit per SM.
//
// V is consumed as V^T [bh, 64, Tpad] (written by the QKV GEMM epilogue) so that both MMAs read
// K-major operands.  Keys >= T in the last tile are masked to -inf (their V^T columns are zero).
//
// Replaces: HF Dinov2SelfAttention / ViTSelfAttention eager+sdpa paths (transformers
// modeling_dinov2.py:153-235) reached from data/utils/feature_extractor.py:49-59.
#include "attention.cuh"
#include "prof.cuh"

namespace ucod {

namespace {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;     // head dim
constexpr int SQ_BYTES = ATT_BM * ATT_D * 2;        // 16 KB
constexpr int SK_BYTES = ATT_BN * ATT_D * 2;        // 16 KB
constexpr int SV_BYTES = ATT_D * ATT_BN * 2;        // 16 KB (two [64 x 64] K-blocks)
constexpr int SP_BYTES = ATT_BM * ATT_BN * 2;       // 32 KB (two [128 x 64] K-blocks)
constexpr int ATT_SMEM = SQ_BYTES + 2 * SK_BYTES + 2 * SV_BYTES + SP_BYTES + 256 + 1024;
constexpr int ATT_TMEM_COLS = 256;                  // S: [0,128)  O: [128,192)
constexpr int ATT_THREADS = 192;

__global__ void __launch_bounds__(ATT_THREADS, 2)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                         const __grid_constant__ CUtensorMap tm_vt, __nv_bfloat16* __restrict__ ctx, int T, int H,
                         float scale_log2e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + SQ_BYTES;
    uint8_t* sV = sK + 2 * SK_BYTES;
    uint8_t* sP = sV + 2 * SV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + SP_BYTES);
    uint64_t* bar_q = bars;            // 1
    uint64_t* bar_kv_full = bars + 1;  // 2
    uint64_t* bar_kv_empty = bars + 3; // 2
    uint64_t* bar_s = bars + 5;        // S tile ready in TMEM
    uint64_t* bar_p = bars + 6;        // P tile ready in smem (128 arrivals)
    uint64_t* bar_pv = bars + 7;       // P V MMA retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_BM;
    const int bh = blockIdx.y;
    const int n_tiles = (T + ATT_BN - 1) / ATT_BN;

    if (threadIdx.x == 0) {
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_kv_full[i], 1);
            mbar_init(&bar_kv_empty[i], 1);
        }
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 128);
        mbar_init(bar_pv, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, ATT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + 128;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&tm_q);
            tma_prefetch_desc(&tm_k);
            tma_prefetch_desc(&tm_vt);
            mbar_arrive_expect_tx(bar_q, SQ_BYTES);
            tma_load_3d(sQ, &tm_q, bar_q, 0, q0, bh);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                const uint32_t use = (uint32_t)(j >> 1);
                mbar_wait(&bar_kv_empty[st], (use & 1) ^ 1);
                mbar_arrive_expect_tx(&bar_kv_full[st], SK_BYTES + SV_BYTES);
                tma_load_3d(sK + st * SK_BYTES, &tm_k, &bar_kv_full[st], 0, j * ATT_BN, bh);
                tma_load_3d(sV + st * SV_BYTES, &tm_vt, &bar_kv_full[st], j * ATT_BN, 0, bh);
                tma_load_3d(sV + st * SV_BYTES + SV_BYTES / 2, &tm_vt, &bar_kv_full[st], j * ATT_BN + 64, 0, bh);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, ATT_BN);
            constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, ATT_D);
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t p_addr = smem_u32(sP);
            auto issue_s = [&](int j) {
                const uint32_t k_addr = smem_u32(sK + (j & 1) * SK_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)
                    umma_bf16_ss(tmem_s, umma_desc_kmajor_sw128(q_addr + k * 32),
                                 umma_desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0);
                umma_commit(bar_s);
            };
            mbar_wait(bar_q, 0);
            mbar_wait(&bar_kv_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                mbar_wait(bar_p, j & 1);  // softmax has consumed S_j and published P_j
                tc_fence_after();
                if (j + 1 < n_tiles) {
                    mbar_wait(&bar_kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
                    tc_fence_after();
                    issue_s(j + 1);
                }
                const uint32_t v_addr = smem_u32(sV + st * SV_BYTES);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_o, umma_desc_kmajor_sw128(p_addr + kb * (SP_BYTES / 2) + k * 32),
                                     umma_desc_kmajor_sw128(v_addr + kb * (SV_BYTES / 2) + k * 32), idesc_o,
                                     (j | kb | k) != 0);
                }
                umma_commit(&bar_kv_empty[st]);
                umma_commit(bar_pv);
            }
        }
    } else {
        // ===================== softmax / correction / epilogue: one thread per query row ==============
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        uint8_t* p_row = sP + (r >> 3) * 1024 + (r & 7) * 128;
        const int rx = r & 7;
        float m_run = -INFINITY, l_run = 0.f;

        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(bar_s, j & 1);
            tc_fence_after();
            const int valid = T - j * ATT_BN;  // keys valid in this tile (>= 1)
            // ---- pass 1: row max ----
            float m_tile = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t u[32];
                tmem_ld32(tmem_s + lane_off + c * 32, u);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float s = (c * 32 + i < valid) ? __uint_as_float(u[i]) : -INFINITY;
                    m_tile = fmaxf(m_tile, s);
                }
            }
            const float m_new = fmaxf(m_run, m_tile);
            const float alpha = exp2f((m_run - m_new) * scale_log2e);  // 0 when m_run = -inf
            const float m_scaled = m_new * scale_log2e;

            // previous P V must have retired before P is overwritten / O is rescaled
            if (j > 0) {
                mbar_wait(bar_pv, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, m_new > m_run)) {
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t u[32];
                        tmem_ld32(tmem_o + lane_off + c * 32, u);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) u[i] = __float_as_uint(__uint_as_float(u[i]) * alpha);
                        tmem_st32(tmem_o + lane_off + c * 32, u);
                    }
                    tmem_wait_st();
                }
            }
            // ---- pass 2: probabilities -> bf16 -> swizzled smem ----
            float l_tile = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t u[32];
                tmem_ld32(tmem_s + lane_off + c * 32, u);
                tmem_wait_ld();
                float p[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float s = __uint_as_float(u[i]);
                    const float e = exp2f(fmaf(s, scale_log2e, -m_scaled));
                    p[i] = (c * 32 + i < valid) ? e : 0.f;
                    l_tile += p[i];
                }
                uint8_t* dst = p_row + (c >> 1) * (SP_BYTES / 2);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int ci = (c & 1) * 4 + g;  // 16-byte chunk index inside the 128-byte row
                    uint4 t;
                    t.x = pack_bf16x2(p[8 * g + 0], p[8 * g + 1]);
                    t.y = pack_bf16x2(p[8 * g + 2], p[8 * g + 3]);
                    t.z = pack_bf16x2(p[8 * g + 4], p[8 * g + 5]);
                    t.w = pack_bf16x2(p[8 * g + 6], p[8 * g + 7]);
                    *reinterpret_cast<uint4*>(dst + ((ci ^ rx) << 4)) = t;
                }
            }
            l_run = l_run * alpha + l_tile;
            m_run = m_new;
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(bar_p);
        }

        // ---- epilogue: O / l -> bf16 ctx ----
        mbar_wait(bar_pv, (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        const int t = q0 + r;
        const int b = bh / H, h = bh - b * H;
        __nv_bfloat16* out = ctx + ((size_t)b * T + t) * (size_t)(H * ATT_D) + h * ATT_D;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t u[32];
            tmem_ld32(tmem_o + lane_off + c * 32, u);
            tmem_wait_ld();
            if (t < T) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(u[8 * g + 0]) * inv_l, __uint_as_float(u[8 * g + 1]) * inv_l);
                    v.y = pack_bf16x2(__uint_as_float(u[8 * g + 2]) * inv_l, __uint_as_float(u[8 * g + 3]) * inv_l);
                    v.z = pack_bf16x2(__uint_as_float(u[8 * g + 4]) * inv_l, __uint_as_float(u[8 * g + 5]) * inv_l);
                    v.w = pack_bf16x2(__uint_as_float(u[8 * g + 6]) * inv_l, __uint_as_float(u[8 * g + 7]) * inv_l);
                    reinterpret_cast<uint4*>(out + c * 32)[g] = v;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ATT_TMEM_COLS);
}

}  // namespace

int launch_attention_d64(const void* q, const void* k, const void* vt, void* ctx, int B, int H, int T, int Tpad,
                         float scale, cudaStream_t stream) {
    UCOD_REQUIRE(B > 0 && H > 0 && T > 0 && Tpad >= T && Tpad % 8 == 0, "attention: bad geometry B=%d H=%d T=%d Tpad=%d",
                 B, H, T, Tpad);
    const uint64_t BH = (uint64_t)B * H;
    CUtensorMap tq, tk, tv;
    if (int rc = make_tmap_3d_bf16(&tq, q, BH, (uint64_t)T, 64, 64, (uint64_t)T * 64, ATT_BM, 64)) return rc;
    if (int rc = make_tmap_3d_bf16(&tk, k, BH, (uint64_t)T, 64, 64, (uint64_t)T * 64, ATT_BN, 64)) return rc;
    if (int rc = make_tmap_3d_bf16(&tv, vt, BH, 64, (uint64_t)Tpad, (uint64_t)Tpad, (uint64_t)Tpad * 64, 64, 64))
        return rc;
    static bool configured = false;
    if (!configured) {
        UCOD_CHECK_CUDA(
            cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(T, ATT_BM), (unsigned)BH);
    {
        ProfScope ps(KC_ATTENTION, stream, 4.0 * B * H * (double)T * T * ATT_D);
        attention_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(
            tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(ctx), T, H, scale * 1.4426950408889634f);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
