// Fused multi-head attention forward (non-causal) on tcgen05 / TMEM, head_dim 64 (ViT-B) or 128 (CORAL CSF,
// head_dim 96 zero-padded).
//
//   ctx[b, t, h*D:(h+1)*D] = softmax_j(q[b,t,h,:] . k[b,j,h,:] * scale) @ v[b,j,h,:]
//
// Q, K, V are read in place from token-major activations ([batch, tokens, ld] bf16, head h at columns h*D —
// e.g. the three column blocks of the fused QKV GEMM output) through 3-D TMA maps; nothing is transposed or
// re-laid-out in HBM.  One CTA = one (batch, head, 128-query tile); two CTAs are resident per SM (D = 64) so one
// CTA's softmax overlaps the other's MMAs.
//
// CTA = 256 threads = two warpgroups:
//   WG0: warp 0 = TMA producer (Q once; K_j / V_j double buffered, separately released), warp 1 = single-thread
//        S issuer (+ TMEM allocator), warp 2 = single-thread P V issuer; registers trimmed with setmaxnreg.dec (80).
//   WG1: 4 softmax warps, one thread per query row (= TMEM lane), registers raised with setmaxnreg.inc (176) so
//        the whole 128-wide S row lives in registers (S is read from TMEM exactly once).
// Measured on B200 (tools/att_timeline.py, tools/ubench/mma_rate.cu): the softmax warps are busy ~85 % of a tile
// period (exp phase bound by the 16/clk/SM MUFU rate shared by the two resident CTAs); dependent N=64 P.V MMAs
// retire every ~74 clk (latency-bound chain), S MMAs every 67 clk.  Tried and rejected (measured slower): a
// TMA L2 prefetch of K/V, the MMA issuer in the high warp ids, 25 % of the exponentials as an FMA-pipe polynomial
// (UCOD_ATT_POLY_EVERY), two softmax threads per row (8 softmax warps, shared-memory max/sum exchange), and
// staggering the two co-resident CTAs by half a tile period, a persistent work-item loop, and a fused two-pipeline
// CTA with an explicit MUFU-phase token (no gain: see DESIGN.md section 4 for the measurements).
// TMEM: S fp32 [0,128) | P bf16-packed [128,192) | O fp32 [192,192+D).
//   S_j  = Q K_j^T                (SS MMA, both operands K-major SW128 tiles)
//   P_j  = exp2(S_j*c - m_ref)    (written back to TMEM as packed bf16; never touches shared memory)
//   O   += P_j V_j                (TS MMA: A = P from TMEM, B = V tile consumed MN-major straight from [keys, D])
// Online softmax with a lazily updated reference maximum: O / l are rescaled only when the row maximum grows by
// more than 2^8 (exact result after the final O / l normalisation; rescales become rare after the first tiles).
// S_{j+1} is issued as soon as the softmax warps have pulled S_j into registers, so the tensor pipe computes the
// next scores while the exponentials of the current tile are evaluated.
//
// Replaces: HF Dinov2SelfAttention / ViTSelfAttention eager+sdpa paths (transformers modeling_dinov2.py:153-235)
// reached from data/utils/feature_extractor.py:49-59, and nn.MultiheadAttention in models/modules/mlp.py:134-148.
#include "attention.cuh"
#include "prof.cuh"

#include <stdlib.h>

namespace ucod {

#ifdef UCOD_ATT_TIMELINE
__device__ long long g_att_tl[6][16][8];  // roles: 0 producer, 1 issuer, 2..5 softmax warps 4..7
#define TL(role, j, slot) do { if (tl_on && (j) < 16) g_att_tl[(role) == 2 ? 2 + (warp & 3) : (role)][j][slot] = clock64() - tl_t0; } while (0)
#else
#define TL(role, j, slot) do {} while (0)
#endif

namespace {

template <int D>
struct AttCfg {
    static constexpr int BM = 128;      // queries per CTA
    static constexpr int BN = 128;      // keys per tile
    static constexpr int NKB = D / 64;  // 64-column (128-byte) blocks per head
    static constexpr int BLK_BYTES = 128 * 64 * 2;  // one [128 x 64] bf16 TMA box
    static constexpr int SQ_BYTES = NKB * BLK_BYTES;
    static constexpr int SK_BYTES = NKB * BLK_BYTES;
    static constexpr int SV_BYTES = NKB * BLK_BYTES;
    static constexpr int STAGES = 2;
#ifndef UCOD_ATT_SMEM_PAD
#define UCOD_ATT_SMEM_PAD 0  // experiment knob: extra dynamic shared memory per CTA (120000 forces one CTA per SM)
#endif
    static constexpr int SMEM = SQ_BYTES + STAGES * (SK_BYTES + SV_BYTES) + 256 + 1024 + UCOD_ATT_SMEM_PAD;
    static constexpr int TM_S = 0, TM_P = 128, TM_O = 192;
    static constexpr int TMEM_COLS = (D == 64) ? 256 : 512;
    static constexpr int THREADS = 256;
    static constexpr int CTAS_PER_SM = (D == 64) ? 2 : 1;
};

// MN-major operand tile, 128-byte swizzle: rows (k index) are 128 B = 64 consecutive MN elements, 8 rows form one
// 1024 B swizzle atom (SBO), further 64-element MN chunks are `lbo_bytes` apart.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp2 on the FMA pipe (Cody-Waite split + degree-3 minimax of 2^f on [-0.5, 0.5], relative error 7.5e-5 — far
// below the bf16 resolution of P).  Used for a fixed fraction of the probabilities so that the 16-lane MUFU unit,
// which bounds the softmax phase at head_dim 64, is not the only exponential engine.  x <= 8 on this path.
// Measured pipe rates per SM sub-partition (tools/ubench/pipe_rate.cu): MUFU.EX2 8 clk per warp instruction,
// FFMA / FADD 1, FFMA2 / FADD2 2 (no arithmetic gain, half the issue slots), F2FP pack 2, FMNMX 1, FMNMX3 2.
// One polynomial exponential costs 6 clk of FMA pipe (3 adds + 3 FMAs) against 8 clk of MUFU, so the two engines
// balance at ~3/8 of the elements on the FMA pipe (tools/ubench/softmax_rate.cu: 1218 -> 1005 clk per tile per SM).
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// 2^x for a pair of values, packed f32x2 arithmetic
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& p0, float& p1) {
    x0 = fmaxf(x0, -126.f);
    x1 = fmaxf(x1, -126.f);
    const uint64_t x = pack2f(x0, x1);
    const uint64_t t = fadd2(x, pack2f(12582912.f, 12582912.f));  // 1.5 * 2^23: round-to-nearest integer in the mantissa
    const uint64_t r = fadd2(t, pack2f(-12582912.f, -12582912.f));
    float r0, r1;
    unpack2f(r, r0, r1);
    const uint64_t f = fadd2(x, pack2f(-r0, -r1));  // in [-0.5, 0.5]
    uint64_t p = ffma2(pack2f(0.05517121031880379f, 0.05517121031880379f), f,
                       pack2f(0.24261027574539185f, 0.24261027574539185f));
    p = ffma2(p, f, pack2f(0.6932609677314758f, 0.6932609677314758f));
    p = ffma2(p, f, pack2f(0.9999281167984009f, 0.9999281167984009f));
    float q0, q1, t0, t1;
    unpack2f(p, q0, q1);
    unpack2f(t, t0, t1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
#ifndef UCOD_ATT_POLY8
#define UCOD_ATT_POLY8 2  // pairs out of every 8 (16 elements) whose exponentials go to the FMA pipe (0 = all on MUFU)
#endif
// Non-blocking phase test: issued early, its result is consumed later, so the ~100 clk barrier round trip overlaps
// arithmetic instead of sitting between two phases of a softmax warp's tile.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
#ifndef UCOD_ATT_EARLY
#define UCOD_ATT_EARLY 1  // early (hidden) barrier polls in the softmax warps
#endif
#ifndef UCOD_ATT_MAXCH
#define UCOD_ATT_MAXCH 8  // independent chains of the row maximum
#endif
template <int N>
__device__ __forceinline__ void reg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

template <int D>
__global__ void __launch_bounds__(256, AttCfg<D>::CTAS_PER_SM)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                         const __grid_constant__ CUtensorMap tm_v, __nv_bfloat16* __restrict__ ctx, int Tq, int Tk,
                         int H, int ld_ctx, float scale_log2e, const int* __restrict__ kv_map,
                         const int* __restrict__ batch_dev) {
    using C = AttCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + C::SQ_BYTES;
    uint8_t* sV = sK + C::STAGES * C::SK_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + C::STAGES * C::SV_BYTES);
    uint64_t* bar_q = bars;           // Q tile landed
    uint64_t* bar_kfull = bars + 1;   // [2]
    uint64_t* bar_kempty = bars + 3;  // [2]
    uint64_t* bar_vfull = bars + 5;   // [2]
    uint64_t* bar_vempty = bars + 7;  // [2]
    uint64_t* bar_s = bars + 9;       // S_j complete in TMEM
    uint64_t* bar_sfree = bars + 10;  // S_j pulled into registers by all 128 rows
    uint64_t* bar_p = bars + 11;      // P_j published in TMEM (128 arrivals)
    uint64_t* bar_pv = bars + 12;     // P_j V_j retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * C::BM;
    // (batch, head) pairs are walked from the end: the QKV GEMM finished with the last images (their rows are the
    // ones still in L2) and the out-projection GEMM starts with the first images' ctx rows
    const int bh = (int)gridDim.y - 1 - (int)blockIdx.y;
    const int b = bh / H, h = bh - b * H;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;  // device-side batch count: whole CTA leaves
    const int col0 = h * D;
    const int n_tiles = (Tk + C::BN - 1) / C::BN;
#ifdef UCOD_ATT_TIMELINE
    const bool tl_on = (blockIdx.x == (gridDim.x > 5 ? 5u : gridDim.x - 1) && blockIdx.y == 300) && (lane == 0);
    const long long tl_t0 = clock64();
#endif

    if (threadIdx.x == 0) {
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_kfull[i], 1);
            mbar_init(&bar_kempty[i], 1);
            mbar_init(&bar_vfull[i], 1);
            mbar_init(&bar_vempty[i], 1);
        }
        mbar_init(bar_s, 1);
        mbar_init(bar_sfree, 128);
        mbar_init(bar_p, 128);
        mbar_init(bar_pv, 1);
        fence_mbar_init();
        // The first loads go out right here (thread 0 is the producer lane): their DRAM/L2 latency overlaps the TMEM
        // allocation, the CTA-wide sync and the register re-partitioning instead of following them (~1000 clk / CTA).
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        const int bkv0 = kv_map != nullptr ? __ldg(kv_map + b) : b;
        mbar_arrive_expect_tx(bar_q, C::SQ_BYTES);
#pragma unroll
        for (int kb = 0; kb < C::NKB; ++kb) tma_load_3d(sQ + kb * C::BLK_BYTES, &tm_q, bar_q, col0 + kb * 64, q0, b);
        mbar_arrive_expect_tx(&bar_kfull[0], C::SK_BYTES);
#pragma unroll
        for (int kb = 0; kb < C::NKB; ++kb)
            tma_load_3d(sK + kb * C::BLK_BYTES, &tm_k, &bar_kfull[0], col0 + kb * 64, 0, bkv0);
        mbar_arrive_expect_tx(&bar_vfull[0], C::SV_BYTES);
#pragma unroll
        for (int kb = 0; kb < C::NKB; ++kb)
            tma_load_3d(sV + kb * C::BLK_BYTES, &tm_v, &bar_vfull[0], col0 + kb * 64, 0, bkv0);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base + C::TM_S;
    const uint32_t tmem_p = tmem_base + C::TM_P;
    const uint32_t tmem_o = tmem_base + C::TM_O;

    if (warp < 4) {
        reg_dec<80>();
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer (tile 0 was issued before the CTA-wide sync) =====================
            const int bkv = kv_map != nullptr ? __ldg(kv_map + b) : b;
            for (int j = 1; j < n_tiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = ((uint32_t)(j >> 1) & 1) ^ 1;
                mbar_wait_parked(&bar_kempty[st], ph);
                TL(0, j, 0);
                mbar_arrive_expect_tx(&bar_kfull[st], C::SK_BYTES);
#pragma unroll
                for (int kb = 0; kb < C::NKB; ++kb)
                    tma_load_3d(sK + st * C::SK_BYTES + kb * C::BLK_BYTES, &tm_k, &bar_kfull[st], col0 + kb * 64,
                                j * C::BN, bkv);
                mbar_wait_parked(&bar_vempty[st], ph);
                TL(0, j, 1);
                mbar_arrive_expect_tx(&bar_vfull[st], C::SV_BYTES);
#pragma unroll
                for (int kb = 0; kb < C::NKB; ++kb)
                    tma_load_3d(sV + st * C::SV_BYTES + kb * C::BLK_BYTES, &tm_v, &bar_vfull[st], col0 + kb * 64,
                                j * C::BN, bkv);
            }
        } else if (warp == 1 && lane == 0) {
            // ===================== S = Q K^T issuer =====================
            // Two issuing lanes (this one and the P V issuer in warp 2): one tcgen05.mma takes 50-60 clk to issue and
            // blocks behind the co-resident CTA's MMAs, so a single lane issuing S_{j+1} and then P_j V_j spent
            // ~1500 clk per tile in issue plus four barrier wake-ups — as long as the softmax warps' own tile period
            // (timeline, round 2).  S and P V only share read-only operands, so they need no mutual ordering.
            constexpr uint32_t idesc_s = umma_idesc_bf16(C::BM, C::BN);
            const uint32_t q_addr = smem_u32(sQ);
            // the last tile only computes the 32-column chunks that hold valid keys (S with N = tail_cols, P V with
            // tail_cols / 16 K-steps); TMA zero-fills the key rows past Tk inside the last chunk
            const int tail_cols = ((Tk - (n_tiles - 1) * C::BN + 31) / 32) * 32;
            const uint32_t idesc_s_tail = umma_idesc_bf16(C::BM, tail_cols);
            mbar_wait_parked(bar_q, 0);
            for (int j = 0; j < n_tiles; ++j) {
                mbar_wait_parked(&bar_kfull[j & 1], (j >> 1) & 1);
                TL(1, j, 0);
                if (j > 0) mbar_wait_parked(bar_sfree, (j - 1) & 1);  // S_{j-1} is in the softmax warps' registers
                TL(1, j, 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sK + (j & 1) * C::SK_BYTES);
                const uint32_t idesc = j == n_tiles - 1 ? idesc_s_tail : idesc_s;
#pragma unroll
                for (int kb = 0; kb < C::NKB; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_s, umma_desc_kmajor_sw128(q_addr + kb * C::BLK_BYTES + k * 32),
                                     umma_desc_kmajor_sw128(k_addr + kb * C::BLK_BYTES + k * 32), idesc,
                                     (kb | k) != 0);
                umma_commit(&bar_kempty[j & 1]);
                umma_commit(bar_s);
                TL(1, j, 2);
            }
        } else if (warp == 2 && lane == 0) {
            // ===================== O += P V issuer =====================
            constexpr uint32_t idesc_o = umma_idesc_bf16(C::BM, D) | (1u << 16);  // B operand MN-major
            const int tail_cols = ((Tk - (n_tiles - 1) * C::BN + 31) / 32) * 32;
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                mbar_wait_parked(&bar_vfull[st], (j >> 1) & 1);
                TL(1, j, 3);
                mbar_wait_parked(bar_p, j & 1);  // P_j published
                TL(1, j, 4);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sV + st * C::SV_BYTES);
                const int ksteps = j == n_tiles - 1 ? tail_cols / 16 : C::BN / 16;
#pragma unroll
                for (int k = 0; k < C::BN / 16; ++k)
                    if (k < ksteps)
                        umma_bf16_ts(tmem_o, tmem_p + k * 8, umma_desc_mnmajor_sw128(v_addr + k * 2048, C::BLK_BYTES),
                                     idesc_o, (j | k) != 0);
                umma_commit(&bar_vempty[st]);
                umma_commit(bar_pv);
                TL(1, j, 5);
            }
        }
    } else {
        // ===================== softmax / correction / epilogue: one thread per query row ==============
        reg_inc<176>();
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        float m_ref = 0.f, l_run = 0.f;
        bool s_ready = false;  // S_j already seen complete by the early poll of the previous tile

        for (int j = 0; j < n_tiles - 1; ++j) {
            TL(2, j, 0);
            if (!s_ready) mbar_wait(bar_s, j & 1);
            TL(2, j, 1);
            tc_fence_after();
            uint32_t u[128];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32(tmem_s + lane_off + c * 32, reinterpret_cast<uint32_t(&)[32]>(u[32 * c]));
            tmem_wait_ld();
            TL(2, j, 2);
            tc_fence_before();
            mbar_arrive(bar_sfree);
            // ---- row maximum (UCOD_ATT_MAXCH independent chains of 3-input maxima) ----
            constexpr int MC = UCOD_ATT_MAXCH, ML = 128 / MC;
            float mx[MC];
#pragma unroll
            for (int c = 0; c < MC; ++c) {
                mx[c] = __uint_as_float(u[ML * c]);
#pragma unroll
                for (int i = 1; i < ML; i += 2)
                    mx[c] = fmaxf(mx[c], fmaxf(__uint_as_float(u[ML * c + i]),
                                               __uint_as_float(u[ML * c + (i + 1 < ML ? i + 1 : i)])));
            }
#pragma unroll
            for (int w = MC / 2; w > 0; w >>= 1)
#pragma unroll
                for (int c = 0; c < w; ++c) mx[c] = fmaxf(mx[c], mx[c + w]);
            const float ms = mx[0] * scale_log2e;
            float alpha = 1.f;
            bool grow = false;
            if (j == 0) {
                m_ref = ms;
            } else if (ms > m_ref + 8.f) {
                alpha = ex2_approx(m_ref - ms);
                m_ref = ms;
                grow = true;
            }
            TL(2, j, 3);
            // ---- probabilities -> packed bf16 (registers) ----
            const float neg_m = -m_ref;
            uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};
            uint32_t pk[64];
            const uint64_t sc2 = pack2f(scale_log2e, scale_log2e), nm2 = pack2f(neg_m, neg_m);
            bool pv_done = (j == 0);
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                if (UCOD_ATT_EARLY && i == 40 && j > 0) pv_done = mbar_test(bar_pv, (j - 1) & 1);
                const uint64_t x = ffma2(pack2f(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1])), sc2, nm2);
                float x0, x1, p0, p1;
                unpack2f(x, x0, x1);
                if ((i & 7) < UCOD_ATT_POLY8) {
                    ex2_poly2(x0, x1, p0, p1);
                } else {
                    p0 = ex2_approx(x0);
                    p1 = ex2_approx(x1);
                }
                ls[i & 3] = fadd2(ls[i & 3], pack2f(p0, p1));
                pk[i] = pack_bf16x2(p0, p1);
            }
            float ls_lo, ls_hi;
            unpack2f(fadd2(fadd2(ls[0], ls[1]), fadd2(ls[2], ls[3])), ls_lo, ls_hi);
            l_run = l_run * alpha + (ls_lo + ls_hi);
            TL(2, j, 7);
            // ---- the previous P V must have retired before P is overwritten / O is rescaled ----
            if (j > 0) {
                if (!pv_done) mbar_wait(bar_pv, (j - 1) & 1);
                TL(2, j, 4);
                tc_fence_after();
                if (__any_sync(0xffffffffu, grow)) {
#pragma unroll 1
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld32(tmem_o + lane_off + c * 32, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tmem_o + lane_off + c * 32, o);
                    }
                }
            }
            tmem_st32(tmem_p + lane_off, reinterpret_cast<uint32_t(&)[32]>(pk[0]));
            tmem_st32(tmem_p + lane_off + 32, reinterpret_cast<uint32_t(&)[32]>(pk[32]));
            TL(2, j, 5);
            // S_{j+1} was issued when this tile's scores had been pulled into registers: poll it while the P store drains
            if (UCOD_ATT_EARLY) s_ready = mbar_test(bar_s, (j + 1) & 1);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_p);
            TL(2, j, 6);
        }

        {
            // ---- last tile: only the nch 32-column chunks that hold valid keys are loaded, exponentiated and
            //      published (warp-uniform guards; everything past `valid` is -inf -> probability 0) ----
            const int j = n_tiles - 1;
            const int valid = Tk - j * C::BN;  // keys valid in this tile (>= 1)
            const int nch = (valid + 31) >> 5;
            if (!s_ready) mbar_wait(bar_s, j & 1);
            tc_fence_after();
            uint32_t u[128];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nch) tmem_ld32(tmem_s + lane_off + c * 32, reinterpret_cast<uint32_t(&)[32]>(u[32 * c]));
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(bar_sfree);
#pragma unroll
            for (int i = 0; i < 128; ++i)
                if (i >= valid) u[i] = 0xff800000u;  // -inf (also the chunks that were not loaded)
            float mx = __uint_as_float(u[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nch) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(u[32 * c + i]));
                }
            const float ms = mx * scale_log2e;
            float alpha = 1.f;
            bool grow = false;
            if (j == 0) {
                m_ref = ms;
            } else if (ms > m_ref + 8.f) {
                alpha = ex2_approx(m_ref - ms);
                m_ref = ms;
                grow = true;
            }
            const float neg_m = -m_ref;
            float ls = 0.f;
            uint32_t pk[64];
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nch) {
#pragma unroll
                    for (int i = 16 * c; i < 16 * c + 16; ++i) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(u[2 * i]), scale_log2e, neg_m));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(u[2 * i + 1]), scale_log2e, neg_m));
                        ls += p0 + p1;
                        pk[i] = pack_bf16x2(p0, p1);
                    }
                }
            l_run = l_run * alpha + ls;
            if (j > 0) {
                mbar_wait(bar_pv, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, grow)) {
#pragma unroll 1
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld32(tmem_o + lane_off + c * 32, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tmem_o + lane_off + c * 32, o);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nch) tmem_st16(tmem_p + lane_off + c * 16, reinterpret_cast<uint32_t(&)[16]>(pk[16 * c]));
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar_p);
        }

        // ---- epilogue: O / l -> bf16 ctx ----
        mbar_wait(bar_pv, (n_tiles - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        const int t = q0 + r;
        __nv_bfloat16* out = ctx + ((size_t)b * Tq + t) * (size_t)ld_ctx + h * D;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_o + lane_off + c * 32, o);
            tmem_wait_ld();
            if (t < Tq) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(o[8 * g + 0]) * inv_l, __uint_as_float(o[8 * g + 1]) * inv_l);
                    v.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv_l, __uint_as_float(o[8 * g + 3]) * inv_l);
                    v.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv_l, __uint_as_float(o[8 * g + 5]) * inv_l);
                    v.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv_l, __uint_as_float(o[8 * g + 7]) * inv_l);
                    reinterpret_cast<uint4*>(out + c * 32)[g] = v;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------
// A few query rows on CUDA cores: when tokens_q = n * 128 + (1..4) (a ViT's 256 patches + CLS), a whole 128-row
// tile for the leftover rows would cost as much as a full one; this kernel does them instead, one CTA per
// (row, batch*head): fp32 scores in shared memory, block max / sum, then the probability-weighted V sum.
// It is bound by re-reading K and V once (HBM/L2): T = 257, 256 images: tiles 0.169 ms + rows 0.058 ms, against
// 0.257 ms with a third q-tile.
// ------------------------------------------------------------------------------------------------
constexpr int kRowsMax = 4;

template <int D>
__global__ void __launch_bounds__(128)
    attention_rows_kernel(const __nv_bfloat16* __restrict__ q, int ld_q, const __nv_bfloat16* __restrict__ k,
                          const __nv_bfloat16* __restrict__ v, int ld_kv, __nv_bfloat16* __restrict__ ctx, int ld_ctx,
                          int Tq, int Tk, int H, int row0, float scale, const int* __restrict__ kv_map,
                          const int* __restrict__ batch_dev) {
    extern __shared__ float sm_rows[];
    float* sq = sm_rows;          // [D]
    float* part = sq + D;         // [128 / (D/2)][D] = 256 partial sums
    float* red = part + 256;      // [4]
    float* sc = red + 4;          // [Tk]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = row0 + blockIdx.x;
    const int b = blockIdx.y / H, h = blockIdx.y - b * H;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    const int bkv = kv_map != nullptr ? __ldg(kv_map + b) : b;
    const __nv_bfloat16* qr = q + ((size_t)b * Tq + row) * (size_t)ld_q + h * D;
    const __nv_bfloat16* kb = k + (size_t)bkv * Tk * (size_t)ld_kv + h * D;
    const __nv_bfloat16* vb = v + (size_t)bkv * Tk * (size_t)ld_kv + h * D;
    if (tid < D) sq[tid] = __bfloat162float(qr[tid]);
    __syncthreads();
    // scores
    float mx = -INFINITY;
#pragma unroll 2
    for (int t = tid; t < Tk; t += 128) {
        const uint4* kr = reinterpret_cast<const uint4*>(kb + (size_t)t * ld_kv);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
            const uint4 w = __ldg(kr + c);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
                acc = fmaf(f.x, sq[c * 8 + 2 * e], acc);
                acc = fmaf(f.y, sq[c * 8 + 2 * e + 1], acc);
            }
        }
        acc *= scale;
        sc[t] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int t = tid; t < Tk; t += 128) {
        const float p = __expf(sc[t] - mx);
        sc[t] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);
    __syncthreads();
    // ctx[d] = sum_t p[t] * V[t][d]: thread = (dim pair, key slice)
    constexpr int DP = D / 2, SL = 128 / DP;
    const int dp = tid % DP, sl = tid / DP;
    float a0 = 0.f, a1 = 0.f;
    const __nv_bfloat16* vcol = vb + 2 * dp;
    int t = sl;
    for (; t + 7 * SL < Tk; t += 8 * SL) {  // 8 independent loads in flight per thread
        uint32_t w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            w[u] = __ldg(reinterpret_cast<const unsigned int*>(vcol + (size_t)(t + u * SL) * ld_kv));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[u]));
            const float p = sc[t + u * SL];
            a0 = fmaf(p, f.x, a0);
            a1 = fmaf(p, f.y, a1);
        }
    }
    for (; t < Tk; t += SL) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vcol + (size_t)t * ld_kv));
        const float p = sc[t];
        a0 = fmaf(p, f.x, a0);
        a1 = fmaf(p, f.y, a1);
    }
    part[sl * D + 2 * dp] = a0;
    part[sl * D + 2 * dp + 1] = a1;
    __syncthreads();
    if (tid < DP) {
        __nv_bfloat16* out = ctx + ((size_t)b * Tq + row) * (size_t)ld_ctx + h * D;
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < SL; ++s2) {
            o0 += part[s2 * D + 2 * tid];
            o1 += part[s2 * D + 2 * tid + 1];
        }
        *reinterpret_cast<__nv_bfloat162*>(out + 2 * tid) = __floats2bfloat162_rn(o0 * inv, o1 * inv);
    }
}

template <int D>
int launch_inst(const AttentionArgs& a, cudaStream_t stream) {
    using C = AttCfg<D>;
    CUtensorMap tq, tk, tv;
    const uint64_t cols = (uint64_t)a.heads * D;
    if (int rc = make_tmap_3d_bf16(&tq, a.q, (uint64_t)a.batch, (uint64_t)a.tokens_q, cols, (uint64_t)a.ld_q,
                                   (uint64_t)a.tokens_q * a.ld_q, C::BM, 64))
        return rc;
    if (int rc = make_tmap_3d_bf16(&tk, a.k, (uint64_t)(a.kv_batch > 0 ? a.kv_batch : a.batch), (uint64_t)a.tokens_kv, cols, (uint64_t)a.ld_kv,
                                   (uint64_t)a.tokens_kv * a.ld_kv, C::BN, 64))
        return rc;
    if (int rc = make_tmap_3d_bf16(&tv, a.v, (uint64_t)(a.kv_batch > 0 ? a.kv_batch : a.batch), (uint64_t)a.tokens_kv, cols, (uint64_t)a.ld_kv,
                                   (uint64_t)a.tokens_kv * a.ld_kv, C::BN, 64))
        return rc;
    auto kern = attention_fwd_kernel<D>;
    static bool configured = false;
    if (!configured) {
        UCOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured = true;
    }
    // a query tail of 1..kRowsMax rows goes to the CUDA-core row kernel instead of a whole extra q-tile
    const int q_tail = a.tokens_q % C::BM;
    const size_t rows_smem = ((size_t)D + 256 + 4 + a.tokens_kv) * sizeof(float);
    const bool split_tail = a.tokens_q > C::BM && q_tail > 0 && q_tail <= kRowsMax && rows_smem <= 48 * 1024;
    dim3 grid((unsigned)(split_tail ? a.tokens_q / C::BM : ceil_div(a.tokens_q, C::BM)), (unsigned)(a.batch * a.heads));
    {
        ProfScope ps(KC_ATTENTION, stream,
                     a.batch_dev ? 0.0 : 4.0 * a.batch * a.heads * (double)a.tokens_q * a.tokens_kv * a.head_dim_real);
        kern<<<grid, C::THREADS, C::SMEM, stream>>>(tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(a.ctx), a.tokens_q,
                                                     a.tokens_kv, a.heads, a.ld_ctx, a.scale * 1.4426950408889634f,
                                                     a.kv_batch_map, a.batch_dev);
    }
    if (split_tail) {
        ProfScope ps(KC_ATTENTION, stream, 0.0);  // its flops are counted with the tile kernel above
        attention_rows_kernel<D><<<dim3((unsigned)q_tail, grid.y), 128, rows_smem, stream>>>(
                static_cast<const __nv_bfloat16*>(a.q), a.ld_q, static_cast<const __nv_bfloat16*>(a.k),
                static_cast<const __nv_bfloat16*>(a.v), a.ld_kv, reinterpret_cast<__nv_bfloat16*>(a.ctx), a.ld_ctx,
            a.tokens_q, a.tokens_kv, a.heads, a.tokens_q - q_tail, a.scale, a.kv_batch_map, a.batch_dev);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

#ifdef UCOD_ATT_TIMELINE
extern "C" int ucod_debug_att_timeline(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_att_tl, sizeof(long long) * 6 * 16 * 8);
}
#endif

int launch_attention(const AttentionArgs& a, cudaStream_t stream) {
    UCOD_REQUIRE(a.q && a.k && a.v && a.ctx, "attention: null pointer");
    UCOD_REQUIRE(a.batch > 0 && a.heads > 0 && a.tokens_q > 0 && a.tokens_kv > 0,
                 "attention: bad geometry batch=%d heads=%d Tq=%d Tk=%d", a.batch, a.heads, a.tokens_q, a.tokens_kv);
    UCOD_REQUIRE(a.head_dim == 64 || a.head_dim == 128, "attention: head_dim %d not supported (64 or 128)", a.head_dim);
    UCOD_REQUIRE(a.ld_q % 8 == 0 && a.ld_kv % 8 == 0 && a.ld_ctx % 8 == 0 && a.ld_q >= a.heads * a.head_dim &&
                     a.ld_kv >= a.heads * a.head_dim && a.ld_ctx >= a.heads * a.head_dim,
                 "attention: row pitches must be multiples of 8 and cover heads*head_dim");
    UCOD_REQUIRE(((uintptr_t)a.q | (uintptr_t)a.k | (uintptr_t)a.v | (uintptr_t)a.ctx) % 16 == 0,
                 "attention: pointers must be 16-byte aligned");
    return a.head_dim == 64 ? launch_inst<64>(a, stream) : launch_inst<128>(a, stream);
}

}  // namespace ucod
