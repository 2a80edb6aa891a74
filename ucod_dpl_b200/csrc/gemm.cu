// Persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   D[M,N] = A[M,K] * W[N,K]^T      A, W: bf16, K contiguous ("TN");  accumulate fp32 in TMEM.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM allocator), warps 2..5 = epilogue.
// Tile 128 x BN x 64, STAGES-deep smem ring (mbarrier full/empty), two TMEM accumulator stages so the
// epilogue of tile i overlaps the main loop of tile i+1. One CTA per SM, static tile striding with the
// N index fastest so that CTAs running concurrently share the same A row-block through L2.
// Epilogue: TMEM -> registers (+bias from a shared-memory tile, GELU) -> 128B-swizzled staging box in shared
// memory -> TMA bulk tensor store (bf16 outputs) or TMA reduce-add (fp32 residual stream: x += acc + bias is
// performed by the L2, the SM never reads x).  Patch-embed / key / decoder outputs use direct row stores.
//
// Replaces (reference, all library-dispatched): HF Dinov2/ViT `nn.Linear` query/key/value/dense/fc1/fc2
// (transformers modeling_dinov2.py:153-235,348-387), patch-embed Conv2d (:38-117), the decoder's
// `decoupling` 1x1 conv (models/modules/DBA.py:13,35) and the CORAL CSF projections (models/modules/mlp.py:116-148).
#include "gemm.cuh"
#include "prof.cuh"

#include <stdlib.h>
#include <string.h>

namespace ucod {

#ifdef UCOD_GEMM_TIMELINE
__device__ long long g_gemm_tl[2][32][4];
#endif

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128;
    static constexpr int BK = 64;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_BYTES = BM * 128;  // one staged output box: 128 rows x 128 B (64 bf16 or 32 fp32 columns)
    static constexpr int BIAS_BYTES = BN * 4;
    static constexpr int BAR_BYTES = 128;
    // +1024: manual alignment slack
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * OUT_BYTES + BIAS_BYTES + BAR_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
    static constexpr int THREADS = 192;
};
static_assert(GemmCfg<256>::SMEM_BYTES <= 227 * 1024, "GEMM shared memory budget exceeded");

// Modes whose output tile is staged in shared memory and written with TMA (bulk tensor store / reduce-add).
__host__ __device__ constexpr bool epi_uses_tma(int mode) {
    return mode == EPI_BIAS_BF16 || mode == EPI_BIAS_GELU_BF16 || mode == EPI_RESID_F32;
}
__host__ __device__ constexpr int epi_box_cols(int mode) { return mode == EPI_RESID_F32 ? 32 : 64; }

// ------------------------------------------------------------------------------------------------
// Epilogue helpers
// ------------------------------------------------------------------------------------------------
// erf-GELU through erfc: 0.5*erfc(|x|/sqrt(2)) = 2^q(|x|), q a degree-5 fit on |x| in [0, 5.94] with the 1/sqrt(2)
// folded into the coefficients; beyond the fit range q keeps falling monotonically (checked up to |x| = 42), so
// 2^q underflows to 0 without a clamp.  |gelu error| < 1e-6, far below the bf16 output resolution.
// x > 0: x - x*e, x <= 0: x*e.  One MUFU + 5 FFMA (|x| as an operand modifier) + 3 FMA-pipe instructions.
__device__ __forceinline__ float gelu_fast(float x) {
    const float t = fabsf(x);
    float q = -0.0004910549614578485f;
    q = fmaf(q, t, 0.007219184655696154f);
    q = fmaf(q, t, -0.05219459533691406f);
    q = fmaf(q, t, -0.45955371856689453f);
    q = fmaf(q, t, -1.1510077714920044f);
    q = fmaf(q, t, -1.0000038146972656f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
    const float r = x * e;
    return x > 0.f ? x - r : r;
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
__device__ __forceinline__ void load32(const float* p, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        v[4 * i + 0] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
    }
}
// acc += bias (bias tile lives in shared memory; every thread reads the same address -> broadcast)
__device__ __forceinline__ void add_bias32(const float* sbias, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = reinterpret_cast<const float4*>(sbias)[i];
        v[4 * i + 0] += t.x, v[4 * i + 1] += t.y, v[4 * i + 2] += t.z, v[4 * i + 3] += t.w;
    }
}
// TMA store / reduce-add of one staged [128 rows x 128 B] box (shared -> global), bulk-group completion.
__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(desc)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* desc, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(desc)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// direct (non-TMA) output modes: one thread owns one row, 32 consecutive columns per call
template <int MODE>
__device__ __forceinline__ void epilogue_direct(const GemmEpi& ep, int m, int n0, int N, float (&v)[32]) {
    if constexpr (MODE == EPI_BIAS_F32) {
        store32_f32(reinterpret_cast<float*>(ep.out) + (size_t)m * ep.ld_out + n0, v);
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
                             const __grid_constant__ CUtensorMap tmap_out, int M, int N, int K, const GemmEpi ep) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint8_t* sOut = smem + STAGES * Cfg::STAGE_BYTES;  // 2 x OUT_BYTES, 1024-aligned
    float* sBias = reinterpret_cast<float*>(sOut + 2 * Cfg::OUT_BYTES);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::BIAS_BYTES);
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

    if (ep.m_dev != nullptr) {  // device-side row count (capacity M)
        const long long md = (long long)__ldg(ep.m_dev) * ep.m_per;
        M = md < (long long)M ? (int)(md < 0 ? 0 : md) : M;
    }
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
                    mbar_wait_parked(&bar_empty[stage], phase ^ 1);
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
#ifdef UCOD_GEMM_TIMELINE
                const long long tl0 = clock64();
                long long tl_full = 0;
#endif
                mbar_wait_parked(&bar_tempty[as], aphase ^ 1);  // epilogue has drained this accumulator stage
#ifdef UCOD_GEMM_TIMELINE
                const long long tl1 = clock64();
#endif
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
#ifdef UCOD_GEMM_TIMELINE
                    const long long tw = clock64();
#endif
                    mbar_wait_parked(&bar_full[stage], phase);
#ifdef UCOD_GEMM_TIMELINE
                    tl_full += clock64() - tw;
#endif
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
#ifdef UCOD_GEMM_TIMELINE
                if (blockIdx.x == 7 && it < 32) {
                    g_gemm_tl[0][it][0] = tl1 - tl0;          // wait for the epilogue (tempty)
                    g_gemm_tl[0][it][1] = tl_full;            // wait for TMA (sum over k-blocks)
                    g_gemm_tl[0][it][2] = clock64() - tl0;    // whole tile (issue side)
                }
#endif
            }
        }
    } else {
        // ===================== Epilogue warps: TMEM -> registers -> (smem -> TMA) | global =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;  // 0..127
        const bool leader = (et == 0);
        if (leader && epi_uses_tma(MODE)) tma_prefetch_desc(&tmap_out);
        uint32_t sub_count = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / n_tiles) * Cfg::BM;
            const int n0 = (tile % n_tiles) * BN;
#ifdef UCOD_GEMM_TIMELINE
            const long long te0 = clock64();
#endif
            // bias tile -> shared (the previous tile's readers are all past their last barrier)
            for (int i = et; i < BN; i += 128) sBias[i] = ep.bias != nullptr ? __ldg(ep.bias + n0 + i) : 0.f;
            mbar_wait_parked(&bar_tfull[as], aphase);
#ifdef UCOD_GEMM_TIMELINE
            const long long te1 = clock64();
#endif
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
            if constexpr (epi_uses_tma(MODE)) {
                constexpr int BOX = epi_box_cols(MODE);  // columns per staged box
                constexpr int NSUB = BN / BOX;
#pragma unroll 1
                for (int sidx = 0; sidx < NSUB; ++sidx, ++sub_count) {
                    uint8_t* stage_out = sOut + (sub_count & 1) * Cfg::OUT_BYTES;
                    if (leader) bulk_wait_read<1>();  // the store that last used this buffer has read it
                    epi_bar_sync();                   // buffer free for everyone (+ bias tile visible)
                    uint8_t* srow = stage_out + row * 128;
                    const int rx = row & 7;
#pragma unroll
                    for (int c = 0; c < BOX / 32; ++c) {
                        uint32_t r[32];
                        tmem_ld32(taddr + sidx * BOX + c * 32, r);
                        tmem_wait_ld();
                        if (sidx == NSUB - 1 && c == BOX / 32 - 1) {  // accumulator stage fully read
                            tc_fence_before();
                            mbar_arrive(&bar_tempty[as]);
                        }
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        add_bias32(sBias + sidx * BOX + c * 32, v);
                        if constexpr (MODE == EPI_RESID_F32) {
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                *reinterpret_cast<float4*>(srow + ((g ^ rx) << 4)) =
                                    make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                        } else {
                            if constexpr (MODE == EPI_BIAS_GELU_BF16) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 t;
                                t.x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
                                t.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
                                t.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
                                t.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
                                *reinterpret_cast<uint4*>(srow + (((c * 4 + g) ^ rx) << 4)) = t;
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    epi_bar_sync();
                    if (leader) {
                        if constexpr (MODE == EPI_RESID_F32)
                            tma_reduce_add_2d(&tmap_out, stage_out, n0 + sidx * BOX, m0);
                        else
                            tma_store_2d(&tmap_out, stage_out, n0 + sidx * BOX, m0);
                        bulk_commit();
                    }
                }
            } else {
                epi_bar_sync();  // bias tile visible
                const int m = m0 + row;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
                    tmem_wait_ld();
                    if (c == BN / 32 - 1) {
                        tc_fence_before();
                        mbar_arrive(&bar_tempty[as]);
                    }
                    if (m < M) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        add_bias32(sBias + c * 32, v);
                        epilogue_direct<MODE>(ep, m, n0 + c * 32, N, v);
                    }
                }
                epi_bar_sync();  // everyone is done with the bias tile before the next one is written
            }
#ifdef UCOD_GEMM_TIMELINE
            if (blockIdx.x == 7 && it < 32 && leader) {
                g_gemm_tl[1][it][0] = te1 - te0;          // epilogue waiting for the accumulator
                g_gemm_tl[1][it][1] = clock64() - te1;    // epilogue busy
            }
#endif
        }
        if (leader && epi_uses_tma(MODE)) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// 2-CTA kernel (cta_group::2): a CTA pair (cluster of two SMs of one TPC) owns one 256 x 256 tile.  Each CTA stages
// its own 128 rows of A and its own half (128 of the 256 N-rows) of B, the leader issues M = 256 MMAs that read both
// CTAs' shared memory, and every CTA keeps the accumulator rows of its 128 output rows in its own TMEM.  Per CTA and
// k-block the MMA reads 32 KB instead of 48 KB and TMA writes 32 KB instead of 48 KB: shared-memory bandwidth —
// the limiter of the 1-CTA 128x256 tile (measured: 96 B/clk MMA reads + 96 B/clk TMA writes on a 128 B/clk port)
// — drops to 64 + 64 B/clk, and the ring is six stages deep.
// Barriers: the leader's `full` barrier collects expect_tx arrivals and TMA bytes from both CTAs (cta_group::2
// loads signal the leader); `empty` and `tfull` are released in both CTAs by multicast tcgen05.commit; the leader's
// `tempty` collects one arrival per epilogue warp of both CTAs.
// ------------------------------------------------------------------------------------------------
struct Gemm2Cfg {
    static constexpr int BM = 128;   // rows per CTA (256 per pair)
    static constexpr int BN = 256;
    static constexpr int BK = 64;
    static constexpr int STAGES = 5;
    static constexpr int OUT_BUFS = 4;  // staged output boxes in flight: a TMA store queues behind the ring's loads
    static constexpr int A_BYTES = BM * BK * 2;        // 16 KB
    static constexpr int B_BYTES = (BN / 2) * BK * 2;  // 16 KB (this CTA's half of the B tile)
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_BYTES = BM * 128;
    static constexpr int BIAS_BYTES = BN * 4;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BUFS * OUT_BYTES + BIAS_BYTES + BAR_BYTES + 1024;
    static constexpr int TMEM_COLS = 512;
    static constexpr int THREADS = 192;
};
static_assert(Gemm2Cfg::SMEM_BYTES <= 227 * 1024, "2-CTA GEMM shared memory budget exceeded");

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* desc, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once all prior MMAs of the pair retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
    gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                              const __grid_constant__ CUtensorMap tmap_out, int M, int N, int K, const GemmEpi ep) {
    using Cfg = Gemm2Cfg;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BN = Cfg::BN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint8_t* sOut = smem + STAGES * Cfg::STAGE_BYTES;
    float* sBias = reinterpret_cast<float*>(sOut + Cfg::OUT_BUFS * Cfg::OUT_BYTES);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::BIAS_BYTES);
    uint64_t* bar_empty = bar_full + STAGES;
    uint64_t* bar_tfull = bar_empty + STAGES;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool is_leader_cta = (rank == 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bar_full[i], 2);   // one expect_tx arrival per CTA of the pair (used in the leader only)
            mbar_init(&bar_empty[i], 1);  // multicast commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_tfull[i], 1);   // multicast commit
            mbar_init(&bar_tempty[i], 8);  // 4 epilogue warps x 2 CTAs (used in the leader only)
        }
        fence_mbar_init();
    }
    cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / TMA signal
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (ep.m_dev != nullptr) {  // device-side row count (capacity M)
        const long long md = (long long)__ldg(ep.m_dev) * ep.m_per;
        M = md < (long long)M ? (int)(md < 0 ? 0 : md) : M;
    }
    const int m_pairs = (M + 2 * Cfg::BM - 1) / (2 * Cfg::BM);
    const int n_tiles = N / BN;
    const int total_tiles = m_pairs * n_tiles;
    const int k_blocks = (K + Cfg::BK - 1) / Cfg::BK;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            tma_prefetch_desc(&tmap_a);
            tma_prefetch_desc(&tmap_b);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                const int m0 = (tile / n_tiles) * 2 * Cfg::BM + (int)rank * Cfg::BM;
                const int n0 = (tile % n_tiles) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_parked(&bar_empty[stage], phase ^ 1);
                    const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[stage]), 0);
                    mbar_arrive_expect_tx_cluster(full_leader, Cfg::STAGE_BYTES);
                    tma_load_2d_pair(sA + stage * Cfg::A_BYTES, &tmap_a, full_leader, kb * Cfg::BK, m0);
                    tma_load_2d_pair(sB + stage * Cfg::B_BYTES, &tmap_b, full_leader, kb * Cfg::BK, n0);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && is_leader_cta) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * Cfg::BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait_parked(&bar_tempty[as], aphase ^ 1);  // both CTAs' epilogues have drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait_parked(&bar_full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < Cfg::BK / 16; ++k) {
                        umma_bf16_ss_pair(tmem_d, umma_desc_kmajor_sw128(a_addr + k * 32),
                                          umma_desc_kmajor_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                    }
                    umma_commit_pair(&bar_empty[stage]);
                    if (kb == k_blocks - 1) umma_commit_pair(&bar_tfull[as]);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else {
        // ===================== Epilogue warps (both CTAs) =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;
        const bool leader = (et == 0);
        if (leader && epi_uses_tma(MODE)) tma_prefetch_desc(&tmap_out);
        uint32_t sub_count = 0;
        int it = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / n_tiles) * 2 * Cfg::BM + (int)rank * Cfg::BM;
            const int n0 = (tile % n_tiles) * BN;
            const uint32_t tempty_leader = mapa_u32(smem_u32(&bar_tempty[as]), 0);
            for (int i = et; i < BN; i += 128) sBias[i] = ep.bias != nullptr ? __ldg(ep.bias + n0 + i) : 0.f;
            mbar_wait_parked(&bar_tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
            if constexpr (epi_uses_tma(MODE) && MODE != EPI_RESID_F32) {
                // bf16 outputs: 64-column boxes = two 32-column TMEM chunks.  The TMEM load of the next chunk is issued
                // before the current one is processed, so its latency (long while the tensor pipe is writing the other
                // accumulator stage) overlaps the bias / GELU / pack work instead of preceding it.
                constexpr int BOX = epi_box_cols(MODE);
                constexpr int NSUB = BN / BOX;
                static_assert(BOX == 64, "two chunks per box");
                auto process = [&](const uint32_t (&r)[32], const float* bias, uint8_t* srow, int rx, int c) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    add_bias32(bias, v);
                    if constexpr (MODE == EPI_BIAS_GELU_BF16) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 t;
                        t.x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
                        t.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
                        t.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
                        t.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
                        *reinterpret_cast<uint4*>(srow + (((c * 4 + g) ^ rx) << 4)) = t;
                    }
                };
                uint32_t rA[32], rB[32];
                tmem_ld32(taddr, rA);
#pragma unroll 1
                for (int sidx = 0; sidx < NSUB; ++sidx, ++sub_count) {
                    uint8_t* stage_out = sOut + (sub_count % Cfg::OUT_BUFS) * Cfg::OUT_BYTES;
                    if (leader) bulk_wait_read<Cfg::OUT_BUFS - 1>();
                    epi_bar_sync();
                    uint8_t* srow = stage_out + row * 128;
                    const int rx = row & 7;
                    tmem_wait_ld();
                    tmem_ld_consume32(rA);
                    tmem_ld32(taddr + sidx * BOX + 32, rB);
                    process(rA, sBias + sidx * BOX, srow, rx, 0);
                    tmem_wait_ld();
                    tmem_ld_consume32(rB);
                    if (sidx == NSUB - 1) {  // accumulator stage fully read
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty_leader);
                    } else {
                        tmem_ld32(taddr + (sidx + 1) * BOX, rA);
                    }
                    process(rB, sBias + sidx * BOX + 32, srow, rx, 1);
                    fence_proxy_async_smem();
                    epi_bar_sync();
                    if (leader) {
                        tma_store_2d(&tmap_out, stage_out, n0 + sidx * BOX, m0);
                        bulk_commit();
                    }
                }
            } else if constexpr (epi_uses_tma(MODE)) {
                constexpr int BOX = epi_box_cols(MODE);
                constexpr int NSUB = BN / BOX;
#pragma unroll 1
                for (int sidx = 0; sidx < NSUB; ++sidx, ++sub_count) {
                    uint8_t* stage_out = sOut + (sub_count % Cfg::OUT_BUFS) * Cfg::OUT_BYTES;
                    if (leader) bulk_wait_read<Cfg::OUT_BUFS - 1>();
                    epi_bar_sync();
                    uint8_t* srow = stage_out + row * 128;
                    const int rx = row & 7;
#pragma unroll
                    for (int c = 0; c < BOX / 32; ++c) {
                        uint32_t r[32];
                        tmem_ld32(taddr + sidx * BOX + c * 32, r);
                        tmem_wait_ld();
                        if (sidx == NSUB - 1 && c == BOX / 32 - 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(tempty_leader);
                        }
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        add_bias32(sBias + sidx * BOX + c * 32, v);
                        if constexpr (MODE == EPI_RESID_F32) {
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                *reinterpret_cast<float4*>(srow + ((g ^ rx) << 4)) =
                                    make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                        } else {
                            if constexpr (MODE == EPI_BIAS_GELU_BF16) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 t;
                                t.x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
                                t.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
                                t.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
                                t.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
                                *reinterpret_cast<uint4*>(srow + (((c * 4 + g) ^ rx) << 4)) = t;
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    epi_bar_sync();
                    if (leader) {
                        if constexpr (MODE == EPI_RESID_F32)
                            tma_reduce_add_2d(&tmap_out, stage_out, n0 + sidx * BOX, m0);
                        else
                            tma_store_2d(&tmap_out, stage_out, n0 + sidx * BOX, m0);
                        bulk_commit();
                    }
                }
            } else {
                epi_bar_sync();
                const int m = m0 + row;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
                    tmem_wait_ld();
                    if (c == BN / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty_leader);
                    }
                    if (m < M) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        add_bias32(sBias + c * 32, v);
                        epilogue_direct<MODE>(ep, m, n0 + c * 32, N, v);
                    }
                }
                epi_bar_sync();
            }
        }
        if (leader && epi_uses_tma(MODE)) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer may still signal our barriers / read our shared memory until it is done too
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)Cfg::TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Host launcher
// ------------------------------------------------------------------------------------------------
template <int BN, int MODE>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpi& ep,
                       cudaStream_t stream) {
    CUtensorMap tout;
    memset(&tout, 0, sizeof(tout));
    if (epi_uses_tma(MODE)) {
        UCOD_REQUIRE(ep.out != nullptr && ep.ld_out >= N, "gemm: output pointer / pitch missing");
        const int esz = MODE == EPI_RESID_F32 ? 4 : 2;
        UCOD_REQUIRE(((size_t)ep.ld_out * esz) % 16 == 0, "gemm: output row pitch must be a multiple of 16 bytes");
        if (int rc = make_tmap_2d(&tout, ep.out, esz, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ld_out, 128,
                                  (uint32_t)epi_box_cols(MODE)))
            return rc;
    }
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
        ProfScope ps(KC_GEMM, stream, ep.m_dev ? 0.0 : 2.0 * M * N * K);  // device-side row count: the caller books the work
        kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, tout, M, N, K, ep);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int MODE>
static int launch_pair_inst(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpi& ep,
                            cudaStream_t stream) {
    using Cfg = Gemm2Cfg;
    CUtensorMap tout;
    memset(&tout, 0, sizeof(tout));
    if (epi_uses_tma(MODE)) {
        UCOD_REQUIRE(ep.out != nullptr && ep.ld_out >= N, "gemm: output pointer / pitch missing");
        const int esz = MODE == EPI_RESID_F32 ? 4 : 2;
        UCOD_REQUIRE(((size_t)ep.ld_out * esz) % 16 == 0, "gemm: output row pitch must be a multiple of 16 bytes");
        if (int rc = make_tmap_2d(&tout, ep.out, esz, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ld_out, 128,
                                  (uint32_t)epi_box_cols(MODE)))
            return rc;
    }
    auto kern = gemm2_bf16_tcgen05_kernel<MODE>;
    static int max_clusters = 0;
    if (max_clusters == 0) {
        UCOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * device_sm_count());
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = device_sm_count() / 2;
        (void)cudaGetLastError();
        max_clusters = n;
    }
    const int tiles = ceil_div(M, 2 * Cfg::BM) * (N / Cfg::BN);
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    {
        ProfScope ps(KC_GEMM, stream, ep.m_dev ? 0.0 : 2.0 * M * N * K);  // device-side row count: the caller books the work
        kern<<<2 * clusters, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, tout, M, N, K, ep);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpi& ep,
                       cudaStream_t s) {
    switch (ep.mode) {
        case EPI_BIAS_BF16: return launch_pair_inst<EPI_BIAS_BF16>(ta, tb, M, N, K, ep, s);
        case EPI_BIAS_GELU_BF16: return launch_pair_inst<EPI_BIAS_GELU_BF16>(ta, tb, M, N, K, ep, s);
        case EPI_RESID_F32: return launch_pair_inst<EPI_RESID_F32>(ta, tb, M, N, K, ep, s);
        case EPI_PATCH: return launch_pair_inst<EPI_PATCH>(ta, tb, M, N, K, ep, s);
        case EPI_BIAS_F32: return launch_pair_inst<EPI_BIAS_F32>(ta, tb, M, N, K, ep, s);
        case EPI_KEYS: return launch_pair_inst<EPI_KEYS>(ta, tb, M, N, K, ep, s);
        default: set_last_error("launch_gemm_bf16: unknown epilogue mode %d", ep.mode); return 1;
    }
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

#ifdef UCOD_GEMM_TIMELINE
extern "C" int ucod_debug_gemm_timeline(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_gemm_tl, sizeof(long long) * 2 * 32 * 4);
}
#endif

int launch_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpi& ep,
                     cudaStream_t stream) {
    UCOD_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    UCOD_REQUIRE(N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
    UCOD_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && K % 8 == 0, "gemm: lda/ldw/K must be multiples of 8");
    const int BN = (N % 256 == 0) ? 256 : 128;
    CUtensorMap ta, tb;
    if (int rc = make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64)) return rc;
    static const bool pair_disabled = getenv("UCOD_GEMM_NO_PAIR") != nullptr;
    if (BN == 256 && M > 256 && !pair_disabled) {  // CTA-pair kernel: 256 x 256 tiles, B halves of 128 rows
        if (int rc = make_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 128, 64)) return rc;
        return launch_pair(ta, tb, M, N, K, ep, stream);
    }
    if (int rc = make_tmap_2d_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)BN, 64)) return rc;
    return BN == 256 ? launch_mode<256>(ta, tb, M, N, K, ep, stream) : launch_mode<128>(ta, tb, M, N, K, ep, stream);
}

}  // namespace ucod
