// Frozen ViT-B last-layer-key extractor (DINOv2-B/14, DINO ViT-B/8), B200-native.
//
// Computes exactly what the reference consumes from the backbone: K_last = LN1_last(x_last) W_k^T + b_k
// for the patch tokens (data/utils/feature_extractor.py:42,49-59), after running layers 0..L-2 in full.
// The last layer's Q/V/attention/MLP and the final LayerNorm are dead work on this path and are skipped.
// The pseudo-label variant additionally returns the last layer's CLS->patch attention row
// (generate_pseudo_label.py:76-81, data/utils/found_bkg_mask.py:24).
//
// Per layer: LN (warp/row, fp32 stats) -> fused QKV GEMM (tcgen05, [M,3D] bf16) -> fused attention reading Q/K/V in place
// (tcgen05) -> out-proj GEMM (+bias, residual reduce-add into the fp32 stream; LayerScale folded into the weights) -> LN -> fc1 GEMM (+bias, erf-GELU)
// -> fc2 GEMM (+bias, LayerScale, residual).
#include "vit.cuh"

#include "attention.cuh"
#include "gemm.cuh"
#include "prof.cuh"

namespace ucod {

// ------------------------------------------------------------------------------------------------
// im2col for the patch-embedding GEMM (+ optional u8 -> normalised conversion).
// out[(b*P + py*gw + px), c*p*p + i*p + j] = norm(img[b, c, py*p+i, px*p+j]) ; columns >= 3*p*p are zero.
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void im2col_patch_kernel(const TIn* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int Hh, int Ww,
                                    int p, int Kpad, float3 mean, float3 inv_std, const int* __restrict__ batch_dev) {
    const int gw = Ww / p, gh = Hh / p;
    const int b = blockIdx.y, py = blockIdx.x;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    const int kvalid = 3 * p * p;
    __nv_bfloat16* orow = out + ((size_t)b * gh * gw + (size_t)py * gw) * Kpad;
    const int wused = gw * p;
    // image rows of this patch-row: 3 channels x p rows x wused pixels, coalesced along x
    for (int idx = threadIdx.x; idx < 3 * p * wused; idx += blockDim.x) {
        const int x = idx % wused;
        const int ci = idx / wused;  // c*p + i
        const int c = ci / p, i = ci - c * p;
        float v = static_cast<float>(img[(((size_t)b * 3 + c) * Hh + (size_t)py * p + i) * Ww + x]);
        if constexpr (sizeof(TIn) == 1) {
            const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
            const float s = c == 0 ? inv_std.x : (c == 1 ? inv_std.y : inv_std.z);
            v = (v / 255.0f - m) * s;
        }
        const int px = x / p, j = x - px * p;
        orow[(size_t)px * Kpad + c * p * p + i * p + j] = __float2bfloat16_rn(v);
    }
    for (int idx = threadIdx.x; idx < gw * (Kpad - kvalid); idx += blockDim.x) {
        const int px = idx / (Kpad - kvalid), k = kvalid + idx % (Kpad - kvalid);
        orow[(size_t)px * Kpad + k] = __float2bfloat16_rn(0.f);
    }
}

// Same mapping, staged through shared memory: one CTA per (image, patch row) gathers the 3 x p image rows with
// coalesced loads into the [gw, Kpad] output layout in smem, then streams the patch rows out with 16-byte stores
// (the direct version writes 2p-byte fragments: 0.36 ms per 64 images @518^2, 7 % of the HBM roofline).
template <typename TIn>
__global__ void im2col_patch_smem_kernel(const TIn* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int Hh,
                                         int Ww, int p, int Kpad, float3 mean, float3 inv_std,
                                         const int* __restrict__ batch_dev) {
    extern __shared__ __align__(16) uint8_t im2col_smem[];
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(im2col_smem);
    const int gw = Ww / p, gh = Hh / p;
    const int b = blockIdx.y, py = blockIdx.x;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    const int kvalid = 3 * p * p;
    const int wused = gw * p;
    const int pitch = Kpad + 8;  // shared-memory row pitch: the patches' rows start 4 banks apart instead of on one bank
    for (int idx = threadIdx.x; idx < gw * (Kpad - kvalid); idx += blockDim.x) {
        const int px = idx / (Kpad - kvalid), k = kvalid + idx % (Kpad - kvalid);
        tile[px * pitch + k] = __float2bfloat16_rn(0.f);
    }
    if constexpr (sizeof(TIn) == 1) {
        // ToTensor + Normalize of a byte has 256 possible results per channel: a 3 x 256 bf16 table built once per CTA
        // (same arithmetic as before, so the patches are bit-identical) replaces a float division, two FP ops and a
        // conversion per pixel by one shared-memory read
        __shared__ __nv_bfloat16 lut[3][256];
        for (int t = threadIdx.x; t < 768; t += blockDim.x) {
            const int c = t >> 8;
            const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
            const float sd = c == 0 ? inv_std.x : (c == 1 ? inv_std.y : inv_std.z);
            lut[c][t & 255] = __float2bfloat16_rn(((float)(t & 255) / 255.0f - m) * sd);
        }
        __syncthreads();
        // one work item = the p bytes of one patch row (channel c, row i, patch px): p table look-ups, written as
        // 32-bit pairs (the destination offset (c*p*p + i*p) is even when p is).  The former 16-byte vector walk with
        // per-byte (row, patch, column) bookkeeping cost ~38 instructions per byte and was issue-bound.
        const int items = 3 * p * gw;
        for (int it = threadIdx.x; it < items; it += blockDim.x) {
            const int px = it % gw;
            const int ci = it / gw;  // c * p + i
            const int c = ci / p, i = ci - c * p;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(img) + (((size_t)b * 3 + c) * Hh + (size_t)py * p + i) * Ww + px * p;
            const __nv_bfloat16* lc = lut[c];
            __nv_bfloat16* dstp = tile + px * pitch + c * p * p + i * p;
            if ((p & 1) == 0) {
                uint32_t* d32 = reinterpret_cast<uint32_t*>(dstp);
                for (int j = 0; j < p; j += 2) {
                    const uint32_t lo = *reinterpret_cast<const uint16_t*>(lc + src[j]);
                    const uint32_t hi = *reinterpret_cast<const uint16_t*>(lc + src[j + 1]);
                    d32[j >> 1] = lo | (hi << 16);
                }
            } else {
                for (int j = 0; j < p; ++j) dstp[j] = lc[src[j]];
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < 3 * p * wused; idx += blockDim.x) {
            const int x = idx % wused;
            const int ci = idx / wused;  // c*p + i
            const int c = ci / p, i = ci - c * p;
            const float v = static_cast<float>(img[(((size_t)b * 3 + c) * Hh + (size_t)py * p + i) * Ww + x]);
            const int px = x / p, j = x - px * p;
            tile[px * pitch + c * p * p + i * p + j] = __float2bfloat16_rn(v);
        }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * gh * gw + (size_t)py * gw) * Kpad);
    const int vec_per_row = Kpad / 8;
    for (int idx = threadIdx.x; idx < gw * vec_per_row; idx += blockDim.x) {
        const int px = idx / vec_per_row, v = idx - px * vec_per_row;
        dst[idx] = *reinterpret_cast<const uint4*>(tile + px * pitch + v * 8);
    }
}

// x[b*T + 0, :] = cls + pos[0, :]
__global__ void cls_init_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                                int T, int D, const int* __restrict__ batch_dev) {
    const int b = blockIdx.x;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    for (int d = threadIdx.x; d < D; d += blockDim.x) x[(size_t)b * T * D + d] = cls[d] + pos[d];
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (D = 768): one warp per row, fp32 two-pass statistics, bf16 output.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ bsh, __nv_bfloat16* __restrict__ y, int rows,
                                      float eps, const int* __restrict__ rows_dev, int rows_per) {
    constexpr int V = D / 128;  // float4 per lane
    if (rows_dev != nullptr) {  // device-side row count (capacity `rows`)
        const int rd = __ldg(rows_dev) * rows_per;
        rows = rd < rows ? rd : rows;
    }
    // rows are walked from the end: the GEMM that produced x finished with its last rows (still in the 126 MB L2),
    // and the GEMM that consumes y starts with the first rows, which are then the ones written last
    const int row = rows - 1 - (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    const int lane = threadIdx.x & 31;
    if (row < 0) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i] = xr[lane + 32 * i];
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mu = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
        q += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    uint2* yr = reinterpret_cast<uint2*>(y + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
        const float4 be = __ldg(reinterpret_cast<const float4*>(bsh) + lane + 32 * i);
        uint2 o;
        o.x = pack_bf16x2((v[i].x - mu) * rstd * g.x + be.x, (v[i].y - mu) * rstd * g.y + be.y);
        o.y = pack_bf16x2((v[i].z - mu) * rstd * g.z + be.z, (v[i].w - mu) * rstd * g.w + be.w);
        yr[lane + 32 * i] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// Last-layer CLS-row attention (pseudo-label variant): attn[b,h,j-1] = softmax_j(q_cls . k_j / sqrt(d)) for
// j = 1..T-1, the softmax running over all T keys (CLS included), found_bkg_mask.py:24.
// One CTA per (b,h): q_cls = xn[b,0,:] W_q[h]^T + b_q[h] (64 dot products of length D), then T logits.
// ------------------------------------------------------------------------------------------------
__global__ void cls_row_attention_kernel(const __nv_bfloat16* __restrict__ xn, const __nv_bfloat16* __restrict__ wq,
                                         const float* __restrict__ bq, const float* __restrict__ keys_all,
                                         float* __restrict__ attn, int T, int D, int H, float scale,
                                         const int* __restrict__ batch_dev) {
    extern __shared__ float sm[];
    float* q = sm;            // 64
    float* logits = sm + 64;  // T
    __shared__ float red[32];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const __nv_bfloat16* xr = xn + (size_t)b * T * D;
    for (int d = warp; d < 64; d += nw) {
        const __nv_bfloat16* wr = wq + (size_t)(h * 64 + d) * D;
        float acc = 0.f;
        for (int k = lane; k < D; k += 32) acc += __bfloat162float(xr[k]) * __bfloat162float(wr[k]);
        acc = warp_sum(acc);
        if (lane == 0) q[d] = acc + bq[h * 64 + d];
    }
    __syncthreads();
    float lmax = -INFINITY;
    for (int j = warp; j < T; j += nw) {
        const float* kr = keys_all + ((size_t)b * T + j) * D + h * 64;
        float acc = q[lane] * kr[lane] + q[lane + 32] * kr[lane + 32];
        acc = warp_sum(acc) * scale;
        if (lane == 0) logits[j] = acc;
        lmax = fmaxf(lmax, acc);
    }
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    float m = -INFINITY;
    for (int i = 0; i < nw; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float lsum = 0.f;
    for (int j = threadIdx.x; j < T; j += blockDim.x) {
        const float e = expf(logits[j] - m);
        logits[j] = e;
        lsum += e;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < nw; ++i) tot += red[i];
    const float inv = 1.0f / tot;
    for (int j = 1 + threadIdx.x; j < T; j += blockDim.x)
        attn[((size_t)b * H + h) * (T - 1) + (j - 1)] = logits[j] * inv;
}

// ------------------------------------------------------------------------------------------------
// Handle + forward
// ------------------------------------------------------------------------------------------------
struct VitModel {
    ucod_vit_cfg cfg;
    const void* patch_w;
    const float* patch_b;
    const float* cls;
    ucod_vit_layer* layers;  // host copy
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct VitWorkspace {
    float* x;
    __nv_bfloat16 *xn, *qkv, *ctx, *h, *patches;
    float* keys_all;
    size_t bytes;
};

static VitWorkspace carve(const ucod_vit_cfg& c, int B, int T, int P, void* base) {
    VitWorkspace w{};
    size_t off = 0;
    auto take = [&](size_t n) {
        void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
        off += align_up(n, 1024);
        return p;
    };
    const size_t M = (size_t)B * T, D = c.hidden;
    w.x = static_cast<float*>(take(M * D * 4));
    w.xn = static_cast<__nv_bfloat16*>(take(M * D * 2));
    w.qkv = static_cast<__nv_bfloat16*>(take(M * 3 * D * 2));  // fused projection output [M, 3D]: Q | K | V
    w.ctx = static_cast<__nv_bfloat16*>(take(M * D * 2));
    const size_t hbytes = M * (size_t)c.mlp_dim * 2;
    const size_t pbytes = (size_t)B * P * c.patch_kpad * 2;
    const size_t kall = M * D * 4;  // fp32 all-token keys of the last layer (pseudo-label variant); aliases h
    size_t big = hbytes > pbytes ? hbytes : pbytes;
    big = big > kall ? big : kall;
    void* hp = take(big);
    w.h = static_cast<__nv_bfloat16*>(hp);
    w.patches = static_cast<__nv_bfloat16*>(hp);
    w.keys_all = static_cast<float*>(hp);
    w.bytes = off;
    return w;
}

int vit_create(void** handle, const ucod_vit_cfg* cfg, const void* patch_w, const float* patch_b, const float* cls,
               const ucod_vit_layer* layers) {
    UCOD_REQUIRE(handle && cfg && patch_w && patch_b && cls && layers, "ucod_vit_create: null argument");
    UCOD_REQUIRE(cfg->hidden == 768 && cfg->heads * 64 == cfg->hidden, "ucod_vit_create: only ViT-B (768 = 12x64)");
    UCOD_REQUIRE(cfg->mlp_dim % 256 == 0 && cfg->layers >= 1 && cfg->patch > 0, "ucod_vit_create: bad config");
    UCOD_REQUIRE(cfg->patch_kpad % 64 == 0 && cfg->patch_kpad >= 3 * cfg->patch * cfg->patch,
                 "ucod_vit_create: patch_kpad must be a multiple of 64 covering 3*p*p");
    VitModel* m = new VitModel();
    m->cfg = *cfg;
    m->patch_w = patch_w;
    m->patch_b = patch_b;
    m->cls = cls;
    m->layers = new ucod_vit_layer[cfg->layers];
    for (int i = 0; i < cfg->layers; ++i) m->layers[i] = layers[i];
    *handle = m;
    return 0;
}

int vit_destroy(void* handle) {
    if (!handle) return 0;
    VitModel* m = static_cast<VitModel*>(handle);
    delete[] m->layers;
    delete m;
    return 0;
}

int vit_workspace_bytes(void* handle, int B, int img_h, int img_w, size_t* out) {
    UCOD_REQUIRE(handle && out, "ucod_vit_workspace_bytes: null argument");
    VitModel* m = static_cast<VitModel*>(handle);
    const int p = m->cfg.patch;
    UCOD_REQUIRE(B > 0 && img_h >= p && img_w >= p, "ucod_vit_workspace_bytes: bad geometry");
    const int P = (img_h / p) * (img_w / p), T = P + 1;
    *out = carve(m->cfg, B, T, P, nullptr).bytes;
    return 0;
}

static int layernorm(const float* x, const float* w, const float* b, __nv_bfloat16* y, int rows, float eps,
                     cudaStream_t s, const int* rows_dev = nullptr, int rows_per = 1) {
    const int wpb = 8;
    ProfScope ps(KC_LAYERNORM, s, rows_dev ? 0.0 : (double)rows * 768 * 6);
    layernorm_bf16_kernel<768><<<ceil_div(rows, wpb), wpb * 32, 0, s>>>(x, w, b, y, rows, eps, rows_dev, rows_per);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int vit_keys(void* handle, const void* images, int image_dtype, int B, int img_h, int img_w, const float* pos_emb,
             void* workspace, size_t ws_bytes, float* keys_f32, void* keys_bf16, float* cls_attn, int keep_cls,
             cudaStream_t stream, const int* batch_dev) {
    UCOD_REQUIRE(handle && images && pos_emb && workspace, "ucod_vit_keys: null argument");
    VitModel* m = static_cast<VitModel*>(handle);
    const ucod_vit_cfg& c = m->cfg;
    const int p = c.patch, D = c.hidden, H = c.heads;
    UCOD_REQUIRE(B > 0 && img_h >= p && img_w >= p, "ucod_vit_keys: bad geometry");
    UCOD_REQUIRE(image_dtype == 0 || image_dtype == 1, "ucod_vit_keys: image_dtype must be 0 (f32) or 1 (u8)");
    UCOD_REQUIRE(keys_f32 || keys_bf16 || cls_attn, "ucod_vit_keys: no output requested");
    const int gh = img_h / p, gw = img_w / p, P = gh * gw, T = P + 1;
    const int M = B * T;
    VitWorkspace w = carve(c, B, T, P, workspace);
    UCOD_REQUIRE(ws_bytes >= w.bytes, "ucod_vit_keys: workspace too small (%zu < %zu bytes)", ws_bytes, w.bytes);
    UCOD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "ucod_vit_keys: workspace must be 1 KiB aligned");

    // ---- embeddings ----
    const float3 mean = make_float3(0.485f, 0.456f, 0.406f);
    const float3 istd = make_float3(1.0f / 0.229f, 1.0f / 0.224f, 1.0f / 0.225f);
    dim3 g_im(gh, B);
    prof_pre(KC_EMBED, stream, batch_dev ? 0.0 : (double)B * 3 * img_h * img_w * (image_dtype ? 1 : 4) + (double)B * P * c.patch_kpad * 2);
    const size_t im_smem = (size_t)gw * (c.patch_kpad + 8) * sizeof(__nv_bfloat16);
    if (im_smem <= 200 * 1024 && c.patch_kpad % 8 == 0) {
        static bool im_configured = false;
        if (!im_configured) {
            UCOD_CHECK_CUDA(cudaFuncSetAttribute(im2col_patch_smem_kernel<float>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            UCOD_CHECK_CUDA(cudaFuncSetAttribute(im2col_patch_smem_kernel<uint8_t>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            im_configured = true;
        }
        if (image_dtype == 0)
            im2col_patch_smem_kernel<float><<<g_im, 512, im_smem, stream>>>(static_cast<const float*>(images), w.patches,
                                                                           B, img_h, img_w, p, c.patch_kpad, mean, istd, batch_dev);
        else
            im2col_patch_smem_kernel<uint8_t><<<g_im, 512, im_smem, stream>>>(
                static_cast<const uint8_t*>(images), w.patches, B, img_h, img_w, p, c.patch_kpad, mean, istd, batch_dev);
    } else if (image_dtype == 0)
        im2col_patch_kernel<float><<<g_im, 256, 0, stream>>>(static_cast<const float*>(images), w.patches, B, img_h,
                                                             img_w, p, c.patch_kpad, mean, istd, batch_dev);
    else
        im2col_patch_kernel<uint8_t><<<g_im, 256, 0, stream>>>(static_cast<const uint8_t*>(images), w.patches, B,
                                                               img_h, img_w, p, c.patch_kpad, mean, istd, batch_dev);
    prof_post(KC_EMBED, stream);
    UCOD_CHECK_CUDA(cudaGetLastError());
    {
        GemmEpi ep;
        ep.mode = EPI_PATCH;
        ep.m_dev = batch_dev, ep.m_per = P;
        ep.bias = m->patch_b;
        ep.pos = pos_emb;
        ep.out = w.x;
        ep.ld_out = D;
        ep.tokens = P;
        if (int rc = launch_gemm_bf16(w.patches, c.patch_kpad, m->patch_w, c.patch_kpad, B * P, D, c.patch_kpad, ep,
                                      stream))
            return rc;
    }
    {
        ProfScope ps(KC_EMBED, stream, batch_dev ? 0.0 : (double)B * D * 4);
        cls_init_kernel<<<B, 256, 0, stream>>>(w.x, m->cls, pos_emb, T, D, batch_dev);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());

    // ---- layers 0 .. L-2 in full ----
    const float scale = 0.125f;  // 1/sqrt(64)
    for (int l = 0; l + 1 < c.layers; ++l) {
        const ucod_vit_layer& L = m->layers[l];
        if (int rc = layernorm(w.x, L.ln1_w, L.ln1_b, w.xn, M, c.ln_eps, stream, batch_dev, T)) return rc;
        {
            GemmEpi ep;
            ep.mode = EPI_BIAS_BF16;
            ep.m_dev = batch_dev, ep.m_per = T;
            ep.bias = L.b_qkv;
            ep.out = w.qkv;
            ep.ld_out = 3 * D;
            if (int rc = launch_gemm_bf16(w.xn, D, L.w_qkv, D, M, 3 * D, D, ep, stream)) return rc;
        }
        {
            AttentionArgs a;
            a.q = w.qkv, a.k = w.qkv + D, a.v = w.qkv + 2 * D, a.ctx = w.ctx;
            a.batch = B, a.heads = H, a.tokens_q = T, a.tokens_kv = T;
            a.ld_q = 3 * D, a.ld_kv = 3 * D, a.ld_ctx = D;
            a.scale = scale;
            a.batch_dev = batch_dev;
            if (int rc = launch_attention(a, stream)) return rc;
        }
        {
            GemmEpi ep;
            ep.mode = EPI_RESID_F32;
            ep.m_dev = batch_dev, ep.m_per = T;
            ep.bias = L.b_o;
            ep.out = w.x;
            ep.ld_out = D;
            if (int rc = launch_gemm_bf16(w.ctx, D, L.w_o, D, M, D, D, ep, stream)) return rc;
        }
        if (int rc = layernorm(w.x, L.ln2_w, L.ln2_b, w.xn, M, c.ln_eps, stream, batch_dev, T)) return rc;
        {
            GemmEpi ep;
            ep.mode = EPI_BIAS_GELU_BF16;
            ep.m_dev = batch_dev, ep.m_per = T;
            ep.bias = L.b_fc1;
            ep.out = w.h;
            ep.ld_out = c.mlp_dim;
            if (int rc = launch_gemm_bf16(w.xn, D, L.w_fc1, D, M, c.mlp_dim, D, ep, stream)) return rc;
        }
        {
            GemmEpi ep;
            ep.mode = EPI_RESID_F32;
            ep.m_dev = batch_dev, ep.m_per = T;
            ep.bias = L.b_fc2;
            ep.out = w.x;
            ep.ld_out = D;
            if (int rc = launch_gemm_bf16(w.h, c.mlp_dim, L.w_fc2, c.mlp_dim, M, D, c.mlp_dim, ep, stream)) return rc;
        }
    }

    // ---- last layer: LN1 + key projection only ----
    const ucod_vit_layer& L = m->layers[c.layers - 1];
    if (int rc = layernorm(w.x, L.ln1_w, L.ln1_b, w.xn, M, c.ln_eps, stream, batch_dev, T)) return rc;
    const __nv_bfloat16* wk = static_cast<const __nv_bfloat16*>(L.w_qkv) + (size_t)D * D;
    const float* bk = L.b_qkv + D;
    if (keys_f32 || keys_bf16) {
        GemmEpi ep;
        ep.mode = EPI_KEYS;
        ep.m_dev = batch_dev, ep.m_per = T;
        ep.bias = bk;
        ep.out = keys_f32;
        ep.out2 = keys_bf16;
        ep.ld_out = D;
        ep.tokens = T;
        ep.skip = keep_cls ? 0 : 1;
        if (int rc = launch_gemm_bf16(w.xn, D, wk, D, M, D, D, ep, stream)) return rc;
    }
    if (cls_attn) {
        // all-token fp32 keys (CLS included) for the CLS-row softmax
        GemmEpi ep;
        ep.mode = EPI_BIAS_F32;
        ep.m_dev = batch_dev, ep.m_per = T;
        ep.bias = bk;
        ep.out = w.keys_all;
        ep.ld_out = D;
        if (int rc = launch_gemm_bf16(w.xn, D, wk, D, M, D, D, ep, stream)) return rc;
        const size_t smem = (64 + (size_t)T) * sizeof(float);
        UCOD_REQUIRE(smem <= 48 * 1024, "ucod_vit_keys: CLS-row attention supports at most %d tokens", 48 * 256 - 64);
        {
            ProfScope ps(KC_EMBED, stream, (double)B * T * D * 4);
            cls_row_attention_kernel<<<B * H, 256, smem, stream>>>(w.xn, static_cast<const __nv_bfloat16*>(L.w_qkv),
                                                                   L.b_qkv, w.keys_all, cls_attn, T, D, H, scale, batch_dev);
        }
        UCOD_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace ucod
