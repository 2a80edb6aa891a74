// Fixed-strategy pseudo-label scoring and small-component cleanup.
//
//  * pseudo_label_score: `compute_img_bkg_seg` (data/utils/found_bkg_mask.py:4-85).  The reference builds the
//    full [P,P] cosine matrix per image and reads ONE row of it; here one CTA per image computes the head
//    sparsity weights beta, the least-attended reference patch, and that single row: each key row is read once
//    (coalesced 128-bit loads, warp-shuffle reductions).  HBM-bound: P*768 key elements + 12*P attention values.
//  * refine_small_components: `refine_post_process` (generate_pseudo_label.py:30-67): 8-connected components of
//    the (<= 32x32) mask in shared memory, OpenCV label order (first 2x2 block in block-raster order), then the
//    reference's sequential "flip isolated components with area < 4" rule.  Integer work, bit-exact.
#include "pseudo_label.cuh"

#include "prof.cuh"

namespace ucod {

namespace {

constexpr int PL_THREADS = 512;
constexpr int PL_MAX_HEADS = 16;

__device__ __forceinline__ float block_sum(float v, float* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}

template <typename TK>
__device__ __forceinline__ float key_at(const TK* p, int i);
template <>
__device__ __forceinline__ float key_at<float>(const float* p, int i) { return p[i]; }
template <>
__device__ __forceinline__ float key_at<__nv_bfloat16>(const __nv_bfloat16* p, int i) { return __bfloat162float(p[i]); }

// attn [B, nh, P] fp32 ; keys [B, P, nh*64] (TK) ; outputs cos [B,P] fp32, bkg [B,P] u8, ref_idx [B] i32,
// gmax: device scalar (ordered-int encoded) receiving max over the launch of (1 - cos).
// MODE 0: everything in one launch, one CTA per image (4-byte scratch; the legacy entry point).
// MODE 1 + MODE 2 (the fast path): the prologue (weights, reference patch, normalised reference descriptor) is a chain of
// block-wide reductions that touches 12 KB of attention per image, while the cosine row streams the whole key block.
// With both in one CTA per image the stream waits for the chain and runs at one CTA's memory parallelism (0.27-0.38
// of HBM with bf16 keys).  MODE 1 runs the chain for all images at once (one small CTA per image) and leaves
// refvec [B, C] / beta [B, 16] in the scratch; MODE 2 is a pure streaming kernel over (image, slab of PL_SLAB patches).
constexpr int PL_SLAB = 64;
template <typename TK, int MODE>
__global__ void __launch_bounds__(PL_THREADS)
    pseudo_label_score_kernel(const float* __restrict__ attn, const TK* __restrict__ keys, float* __restrict__ cos_out,
                              uint8_t* __restrict__ bkg_out, int* __restrict__ ref_out, int* __restrict__ gmax,
                              int nh, int P, float th_bkg, float epsilon, int apply_weights,
                              float* __restrict__ refvec, float* __restrict__ betas) {
    extern __shared__ float sm[];
    float* s_att = sm;                                  // nh * P (MODE 0 / 1)
    float* s_ref = MODE == 2 ? sm : s_att + nh * P;     // nh * 64 (normalised, beta-weighted reference descriptor)
    __shared__ float red[PL_THREADS / 32];
    __shared__ float s_beta[PL_MAX_HEADS];
    __shared__ int s_cnt[PL_MAX_HEADS];
    __shared__ float s_minv[PL_THREADS / 32];
    __shared__ int s_mini[PL_THREADS / 32];
    __shared__ int s_refidx;

    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int C = nh * 64;
    const float* att_b = attn + (size_t)b * nh * P;
    const TK* keys_b = keys + (size_t)b * P * C;

    if constexpr (MODE != 2) {
        // ---- threshold = mean attention; Q_h = fraction of patches above it; beta_h ----
        float acc = 0.f;
        for (int i = threadIdx.x; i < nh * P; i += blockDim.x) {
            const float a = att_b[i];
            s_att[i] = a;
            acc += a;
        }
        if (threadIdx.x < PL_MAX_HEADS) s_cnt[threadIdx.x] = 0;
        const float thr = block_sum(acc, red) / (float)(nh * P);
        for (int h = 0; h < nh; ++h) {
            int c = 0;
            for (int p = threadIdx.x; p < P; p += blockDim.x) c += s_att[h * P + p] > thr ? 1 : 0;
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0 && c) atomicAdd(&s_cnt[h], c);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int h = 0; h < nh; ++h) tot += (float)s_cnt[h] / (float)P + epsilon;
            // apply_weights = False (found_bkg_mask.py:44-47,64-65): neither the descriptors nor the attention sum are weighted
            for (int h = 0; h < nh; ++h) s_beta[h] = apply_weights ? logf(tot / ((float)s_cnt[h] / (float)P + epsilon)) : 1.f;
        }
        __syncthreads();

        // ---- reference patch = argmin_p sum_h att[h,p] * beta[h] (first index on ties) ----
        float best = INFINITY;
        int besti = 0x7fffffff;
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            float s = 0.f;
            for (int h = 0; h < nh; ++h) s += s_att[h * P + p] * s_beta[h];
            if (s < best) best = s, besti = p;  // p increases per thread -> keeps the first minimum
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ov < best || (ov == best && oi < besti)) best = ov, besti = oi;
        }
        if (lane == 0) s_minv[warp] = best, s_mini[warp] = besti;
        __syncthreads();
        if (threadIdx.x == 0) {
            float bv = s_minv[0];
            int bi = s_mini[0];
            for (int i = 1; i < nw; ++i)
                if (s_minv[i] < bv || (s_minv[i] == bv && s_mini[i] < bi)) bv = s_minv[i], bi = s_mini[i];
            s_refidx = bi;
            ref_out[b] = bi;
        }
        __syncthreads();
        const int ref = s_refidx;

        // ---- normalised reference descriptor ----
        {
            const TK* kr = keys_b + (size_t)ref * C;
            float q = 0.f;
            for (int i = threadIdx.x; i < C; i += blockDim.x) {
                const float v = key_at<TK>(kr, i) * s_beta[i >> 6];
                s_ref[i] = v;
                q += v * v;
            }
            const float nrm = fmaxf(sqrtf(block_sum(q, red)), 1e-12f);
            for (int i = threadIdx.x; i < C; i += blockDim.x) s_ref[i] /= nrm;
            __syncthreads();
        }

        if constexpr (MODE == 1) {  // publish for the streaming kernel
            for (int i = threadIdx.x; i < C; i += blockDim.x) refvec[(size_t)b * C + i] = s_ref[i];
            if (threadIdx.x < PL_MAX_HEADS) betas[b * PL_MAX_HEADS + threadIdx.x] = threadIdx.x < nh ? s_beta[threadIdx.x] : 0.f;
            return;
        }
    }
    // ---- one warp per patch: cos(ref, p); 128-bit key loads (8 bf16 / 4 fp32 per lane and step).
    // Two patches per iteration with all their loads issued before the first reduction: with one patch in flight a
    // warp alternates between a DRAM round trip and two dependent shuffle trees, and the kernel took the same time
    // for bf16 keys as for fp32 ones (0.27 vs 0.60 of the HBM roofline: latency-, not bandwidth-bound). ----
    constexpr int EPV = 16 / sizeof(TK);  // elements per 16-byte vector (a vector never straddles a 64-wide head)
    constexpr int VMAX = (PL_MAX_HEADS * 64 / EPV + 31) / 32;  // vectors per lane and patch (upper bound)
    const int nvec = C / EPV;
    float wmax = -INFINITY;
    const int p_begin = MODE == 2 ? (int)blockIdx.y * PL_SLAB : 0;
    const int p_end = MODE == 2 ? min(P, p_begin + PL_SLAB) : P;
    auto load_pair = [&](int p0, uint4 (&raw)[2][VMAX]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = p0 + u < p_end ? p0 + u : p0;
            const uint4* kr = reinterpret_cast<const uint4*>(keys_b + (size_t)p * C);
#pragma unroll
            for (int i = 0; i < VMAX; ++i) {
                const int v = lane + 32 * i;
                raw[u][i] = v < nvec ? __ldg(kr + v) : make_uint4(0u, 0u, 0u, 0u);
            }
        }
    };
    auto finish_pair = [&](int p0, const uint4 (&raw)[2][VMAX]) {
        float dot[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < VMAX; ++i) {
            const int v = lane + 32 * i;
            if (v < nvec) {
                const float beta = s_beta[(v * EPV) >> 6];
                const float* rf = s_ref + v * EPV;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    float x[EPV];
                    if constexpr (sizeof(TK) == 2) {
                        const uint32_t wds[4] = {raw[u][i].x, raw[u][i].y, raw[u][i].z, raw[u][i].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            x[2 * e] = __uint_as_float(wds[e] << 16);
                            x[2 * e + 1] = __uint_as_float(wds[e] & 0xffff0000u);
                        }
                    } else {
                        x[0] = __uint_as_float(raw[u][i].x), x[1] = __uint_as_float(raw[u][i].y);
                        x[2] = __uint_as_float(raw[u][i].z), x[3] = __uint_as_float(raw[u][i].w);
                    }
#pragma unroll
                    for (int e = 0; e < EPV; ++e) {
                        const float val = x[e] * beta;
                        dot[u] += val * rf[e];
                        q[u] += val * val;
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // four interleaved shuffle trees
            dot[0] += __shfl_xor_sync(0xffffffffu, dot[0], o);
            dot[1] += __shfl_xor_sync(0xffffffffu, dot[1], o);
            q[0] += __shfl_xor_sync(0xffffffffu, q[0], o);
            q[1] += __shfl_xor_sync(0xffffffffu, q[1], o);
        }
        if (lane < 2 && p0 + lane < p_end) {
            const float c = (lane == 0 ? dot[0] : dot[1]) / fmaxf(sqrtf(lane == 0 ? q[0] : q[1]), 1e-12f);
            cos_out[(size_t)b * P + p0 + lane] = c;
            bkg_out[(size_t)b * P + p0 + lane] = c > th_bkg ? 1 : 0;
            wmax = fmaxf(wmax, 1.f - c);
        }
    };
    if constexpr (MODE == 2) {
        // the first pair's loads are in flight while the reference descriptor arrives from L2
        int p0 = p_begin + 2 * warp;
        uint4 raw[2][VMAX];
        if (p0 < p_end) load_pair(p0, raw);
        for (int i = threadIdx.x; i < C; i += blockDim.x) s_ref[i] = refvec[(size_t)b * C + i];
        if (threadIdx.x < PL_MAX_HEADS) s_beta[threadIdx.x] = betas[b * PL_MAX_HEADS + threadIdx.x];
        __syncthreads();
        for (bool first = true; p0 < p_end; p0 += 2 * nw, first = false) {
            if (!first) load_pair(p0, raw);
            finish_pair(p0, raw);
        }
    } else {
        for (int p0 = p_begin + 2 * warp; p0 < p_end; p0 += 2 * nw) {
            uint4 raw[2][VMAX];
            load_pair(p0, raw);
            finish_pair(p0, raw);
        }
    }
    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, 1));
    if (lane == 0 && gmax != nullptr && wmax > -INFINITY) {
        // float max through an order-preserving int encoding
        int enc = __float_as_int(wmax);
        enc = enc >= 0 ? enc : enc ^ 0x7fffffff;
        atomicMax(gmax, enc);
    }
}

// sim[b,p] = (1 - cos) / (max + 1e-10) * (1 - bkg)      (found_bkg_mask.py:81-85; `max` is launch-global)
__global__ void pseudo_label_sim_kernel(const float* __restrict__ cos_in, const uint8_t* __restrict__ bkg,
                                        const int* __restrict__ gmax, float* __restrict__ sim, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int enc = *gmax;
    enc = enc >= 0 ? enc : enc ^ 0x7fffffff;
    const float mx = __int_as_float(enc);
    sim[i] = ((1.f - cos_in[i]) / (mx + 1e-10f)) * (bkg[i] ? 0.f : 1.f);
}

// ------------------------------------------------------------------------------------------------
// refine_post_process: one CTA per mask, everything in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int RF_MAX_PIX = 1024;

__global__ void __launch_bounds__(256)
    refine_small_components_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W,
                                   int area_threshold) {
    __shared__ uint8_t s_mask[RF_MAX_PIX];     // original mask (non-zero = fg)
    __shared__ uint8_t s_ref[RF_MAX_PIX];      // progressively refined mask
    __shared__ int s_lab[RF_MAX_PIX];          // root = min raster index of the component, -1 = bg
    __shared__ int s_area[RF_MAX_PIX], s_x0[RF_MAX_PIX], s_x1[RF_MAX_PIX], s_y0[RF_MAX_PIX], s_y1[RF_MAX_PIX],
        s_key[RF_MAX_PIX];
    __shared__ int s_changed;
    const int n = H * W;
    const uint8_t* src = in + (size_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint8_t v = src[i];
        s_mask[i] = v;
        s_ref[i] = v;
        s_lab[i] = v ? i : -1;
        s_area[i] = 0;
        s_x0[i] = W, s_x1[i] = -1, s_y0[i] = H, s_y1[i] = -1;
        s_key[i] = 0x7fffffff;
    }
    __syncthreads();
    // min-label propagation over the 8-neighbourhood until a fixed point
    while (true) {
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (s_lab[i] < 0) continue;
            const int y = i / W, x = i - y * W;
            int best = s_lab[i];
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    const int l = s_lab[yy * W + xx];
                    if (l >= 0 && l < best) best = l;
                }
            if (best < s_lab[i]) {
                // pointer-jump one step to speed convergence
                const int bb = s_lab[best];
                s_lab[i] = (bb >= 0 && bb < best) ? bb : best;
                s_changed = 1;
            }
        }
        __syncthreads();
        if (!s_changed) break;
        __syncthreads();
    }
    // per-component stats, keyed by root
    const int bw = (W + 1) / 2;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = s_lab[i];
        if (r < 0) continue;
        const int y = i / W, x = i - y * W;
        atomicAdd(&s_area[r], 1);
        atomicMin(&s_x0[r], x);
        atomicMax(&s_x1[r], x);
        atomicMin(&s_y0[r], y);
        atomicMax(&s_y1[r], y);
        atomicMin(&s_key[r], (y >> 1) * bw + (x >> 1));
    }
    __syncthreads();
    // sequential rule, components visited in OpenCV label order (increasing first-block key)
    if (threadIdx.x == 0) {
        int last_key = -1;
        while (true) {
            int r = -1, rk = 0x7fffffff;
            for (int i = 0; i < n; ++i)
                if (s_lab[i] == i && s_area[i] < area_threshold && s_key[i] > last_key && s_key[i] < rk)
                    r = i, rk = s_key[i];
            if (r < 0) break;
            last_key = rk;
            const int x = s_x0[r], y = s_y0[r], w = s_x1[r] - s_x0[r] + 1, h = s_y1[r] - s_y0[r] + 1;
            const int xs = max(x - 1, 0), ys = max(y - 1, 0), xe = min(x + w + 1, W), ye = min(y + h + 1, H);
            const int comp_label = s_ref[(y + h / 2) * W + (x + w / 2)];
            const int opposite = (1 - comp_label) & 0xff;
            bool all = true;
            for (int yy = ys; yy < ye && all; ++yy)
                for (int xx = xs; xx < xe; ++xx) {
                    const int i = yy * W + xx;
                    if (s_lab[i] == r) continue;  // the component's own pixels are not part of the ring
                    if (s_ref[i] != opposite) {
                        all = false;
                        break;
                    }
                }
            if (all)
                for (int yy = y; yy < y + h; ++yy)
                    for (int xx = x; xx < x + w; ++xx)
                        if (s_lab[yy * W + xx] == r) s_ref[yy * W + xx] = (uint8_t)opposite;
        }
    }
    __syncthreads();
    uint8_t* dst = out + (size_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = s_ref[i];
}

}  // namespace

size_t pseudo_label_scratch_bytes(int B, int nh) {
    return 256 + (size_t)B * ((size_t)nh * 64 + PL_MAX_HEADS) * sizeof(float);
}

template <typename TK>
static int launch_score(const float* attn_cls, const TK* keys, int B, int nh, int P, float th_bkg, float epsilon,
                        float* cos_out, uint8_t* bkg_out, int* ref_out, int* scratch, size_t scratch_bytes,
                        cudaStream_t stream, int apply_weights, double bytes) {
    const size_t smem = ((size_t)nh * P + (size_t)nh * 64) * sizeof(float);
    UCOD_REQUIRE(smem <= 200 * 1024, "pseudo_label_score: %d patches do not fit in shared memory", P);
    if (scratch != nullptr && scratch_bytes >= pseudo_label_scratch_bytes(B, nh)) {
        float* refvec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + 256);
        float* betas = refvec + (size_t)B * nh * 64;
        auto k1 = pseudo_label_score_kernel<TK, 1>;
        auto k2 = pseudo_label_score_kernel<TK, 2>;
        if (smem > 48 * 1024)
            UCOD_CHECK_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ProfScope ps(KC_PSEUDO, stream, bytes);
        k1<<<B, 256, smem, stream>>>(attn_cls, keys, cos_out, bkg_out, ref_out, scratch, nh, P, th_bkg, epsilon,
                                     apply_weights, refvec, betas);
        k2<<<dim3(B, ceil_div(P, PL_SLAB)), 256, (size_t)nh * 64 * sizeof(float), stream>>>(
                attn_cls, keys, cos_out, bkg_out, ref_out, scratch, nh, P, th_bkg, epsilon, apply_weights, refvec, betas);
    } else {
        auto kern = pseudo_label_score_kernel<TK, 0>;
        if (smem > 48 * 1024)
            UCOD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ProfScope ps(KC_PSEUDO, stream, bytes);
        kern<<<B, PL_THREADS, smem, stream>>>(attn_cls, keys, cos_out, bkg_out, ref_out, scratch, nh, P, th_bkg, epsilon,
                                              apply_weights, nullptr, nullptr);
    }
    return 0;
}

int pseudo_label_score(const float* attn_cls, const void* keys, int keys_bf16, int B, int nh, int P, float th_bkg,
                       float epsilon, float* cos_out, uint8_t* bkg_out, int* ref_out, float* sim_out, int* scratch,
                       cudaStream_t stream, int apply_weights, size_t scratch_bytes) {
    UCOD_REQUIRE(attn_cls && keys && cos_out && bkg_out && ref_out, "pseudo_label_score: null argument");
    UCOD_REQUIRE(B > 0 && P > 0 && nh > 0 && nh <= PL_MAX_HEADS, "pseudo_label_score: bad geometry (heads <= 16)");
    UCOD_REQUIRE(sim_out == nullptr || scratch != nullptr, "pseudo_label_score: sim_map needs the 4-byte scratch");
    if (scratch) {
        // 0x80808080 decodes (ordered-int encoding) to about -3.39e38: below any real 1 - cos
        UCOD_CHECK_CUDA(cudaMemsetAsync(scratch, 0x80, sizeof(int), stream));
    }
    const double bytes = (double)B * P * nh * 64 * (keys_bf16 ? 2 : 4) + (double)B * nh * P * 4 + (double)B * P * 5;
    if (keys_bf16) {
        if (int rc = launch_score(attn_cls, static_cast<const __nv_bfloat16*>(keys), B, nh, P, th_bkg, epsilon, cos_out,
                                  bkg_out, ref_out, scratch, scratch_bytes, stream, apply_weights, bytes))
            return rc;
    } else {
        if (int rc = launch_score(attn_cls, static_cast<const float*>(keys), B, nh, P, th_bkg, epsilon, cos_out, bkg_out,
                                  ref_out, scratch, scratch_bytes, stream, apply_weights, bytes))
            return rc;
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    if (sim_out) {
        const int n = B * P;
        ProfScope ps(KC_PSEUDO, stream, (double)n * 9);
        pseudo_label_sim_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(cos_out, bkg_out, scratch, sim_out, n);
        UCOD_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

int refine_small_components(const uint8_t* mask_in, uint8_t* mask_out, int B, int H, int W, int area_threshold,
                            cudaStream_t stream) {
    UCOD_REQUIRE(mask_in && mask_out && B > 0 && H > 0 && W > 0, "refine_small_components: bad argument");
    UCOD_REQUIRE(H * W <= RF_MAX_PIX, "refine_small_components: masks up to %d pixels are supported (got %dx%d)",
                 RF_MAX_PIX, H, W);
    ProfScope ps(KC_PSEUDO, stream, (double)B * H * W * 2);
    refine_small_components_kernel<<<B, 256, 0, stream>>>(mask_in, mask_out, H, W, area_threshold);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
