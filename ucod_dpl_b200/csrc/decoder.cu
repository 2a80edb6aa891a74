// Dual-Branch Adversarial decoder forward (RevDecoder, models/modules/DBA.py:31-59) and friends.
//
// The reference upsamples the 768-channel key map 37^2 -> 68^2 (engine/runner/loop_UCOD_DPL.py:153,305) and
// then applies the 1x1 `decoupling` conv.  Both are linear and the bilinear weights sum to one, so the conv is
// done FIRST on the small grid (tcgen05 GEMM, 768 -> 128) and the 128-channel result is upsampled on the fly:
//   1. d_in[B*P_in,128] = keys * W_d^T + b_d                      (GEMM, fp32 out, stays in L2)
//   2. sumsq[b,c] = sum_pix bilinear(d_in)[pix,c]^2              (per-channel spatial L2 norm of DBA.py:40-41)
//   3. per pixel: d = bilinear(d_in); f = d*e/max(|e|*sqrt(sumsq),1e-12); a = sigmoid(f*d)+d; 64->1 heads
//   4. (student only) orthogonality loss via the Gram identity instead of the [B,HW,HW] bmm of DBA.py:25-29:
//        sum_{i!=j}(f1_i.f2_j)^2 = <F1^T F1, F2^T F2>_F - sum_i (f1_i.f2_i)^2
// Memory-bound: the only HBM-sized read is the bf16 key map (2.1 MB / image).
#include "decoder.cuh"

#include "gemm.cuh"
#include "prof.cuh"
#include "wgrad.cuh"

namespace ucod {

// bilinear source index, F.interpolate(align_corners=False) semantics (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void bilinear_tap(int dst, int in_size, int out_size, int& i0, int& i1, float& l1) {
    if (in_size == out_size) {
        i0 = dst, i1 = dst, l1 = 0.f;
        return;
    }
    const float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

// One warp = one output pixel, lane = 4 consecutive channels of the 128.
__device__ __forceinline__ float4 sample_d(const float* __restrict__ d_img, int gin_w, int y0, int y1, float ly, int x0,
                                           int x1, float lx, int lane) {
    const float4 v00 = reinterpret_cast<const float4*>(d_img + ((size_t)y0 * gin_w + x0) * 128)[lane];
    const float4 v01 = reinterpret_cast<const float4*>(d_img + ((size_t)y0 * gin_w + x1) * 128)[lane];
    const float4 v10 = reinterpret_cast<const float4*>(d_img + ((size_t)y1 * gin_w + x0) * 128)[lane];
    const float4 v11 = reinterpret_cast<const float4*>(d_img + ((size_t)y1 * gin_w + x1) * 128)[lane];
    const float hx = 1.f - lx, hy = 1.f - ly;
    float4 r;
    r.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    r.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    r.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    r.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    return r;
}

constexpr int DEC_PIX_PER_BLOCK = 64;  // 8 warps x 8 pixels

// sumsq[b, c] += sum over this block's pixels of d_up[pix, c]^2
__global__ void __launch_bounds__(256)
    decoder_sumsq_kernel(const float* __restrict__ d_in, float* __restrict__ sumsq, int gin_h, int gin_w, int out_h,
                         int out_w, const int* __restrict__ batch_dev) {
    __shared__ float4 part[8][32];
    const int b = blockIdx.y;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* d_img = d_in + (size_t)b * gin_h * gin_w * 128;
    const int npix = out_h * out_w;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < DEC_PIX_PER_BLOCK / 8; ++i) {
        const int pix = blockIdx.x * DEC_PIX_PER_BLOCK + i * 8 + warp;
        if (pix >= npix) break;
        const int oy = pix / out_w, ox = pix - oy * out_w;
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_tap(oy, gin_h, out_h, y0, y1, ly);
        bilinear_tap(ox, gin_w, out_w, x0, x1, lx);
        const float4 d = sample_d(d_img, gin_w, y0, y1, ly, x0, x1, lx, lane);
        acc.x += d.x * d.x, acc.y += d.y * d.y, acc.z += d.z * d.z, acc.w += d.w * d.w;
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        float4 t = part[0][lane];
        for (int w = 1; w < 8; ++w) {
            const float4 u = part[w][lane];
            t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
        }
        float* dst = sumsq + (size_t)b * 128 + lane * 4;
        atomicAdd(dst + 0, t.x);
        atomicAdd(dst + 1, t.y);
        atomicAdd(dst + 2, t.z);
        atomicAdd(dst + 3, t.w);
    }
}

// gate + heads (+ optional Gram accumulation for the orthogonality loss)
__global__ void __launch_bounds__(256)
    decoder_head_kernel(const float* __restrict__ d_in, const float* __restrict__ sumsq, const float* __restrict__ emb,
                        const float* __restrict__ w_fg, const float* __restrict__ b_fg, const float* __restrict__ w_bg,
                        const float* __restrict__ b_bg, float* __restrict__ fg, float* __restrict__ bg,
                        float* __restrict__ fhat_out, int gin_h, int gin_w, int out_h, int out_w,
                        const int* __restrict__ batch_dev) {
    const int b = blockIdx.y;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* d_img = d_in + (size_t)b * gin_h * gin_w * 128;
    const int npix = out_h * out_w;
    // per-channel constants for this lane's 4 channels
    const float4 ss = reinterpret_cast<const float4*>(sumsq + (size_t)b * 128)[lane];
    const float4 e = __ldg(reinterpret_cast<const float4*>(emb) + lane);  // emb[2,64] flat == channel order
    const float4 wh = lane < 16 ? __ldg(reinterpret_cast<const float4*>(w_fg) + lane)
                                : __ldg(reinterpret_cast<const float4*>(w_bg) + (lane - 16));
    float4 g;  // e / max(|e| * sqrt(sumsq), 1e-12)
    g.x = e.x / fmaxf(fabsf(e.x) * sqrtf(ss.x), 1e-12f);
    g.y = e.y / fmaxf(fabsf(e.y) * sqrtf(ss.y), 1e-12f);
    g.z = e.z / fmaxf(fabsf(e.z) * sqrtf(ss.z), 1e-12f);
    g.w = e.w / fmaxf(fabsf(e.w) * sqrtf(ss.w), 1e-12f);
    const float bias_fg = __ldg(b_fg), bias_bg = __ldg(b_bg);

    for (int i = 0; i < DEC_PIX_PER_BLOCK / 8; ++i) {
        const int pix = blockIdx.x * DEC_PIX_PER_BLOCK + i * 8 + warp;
        if (pix >= npix) break;
        const int oy = pix / out_w, ox = pix - oy * out_w;
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_tap(oy, gin_h, out_h, y0, y1, ly);
        bilinear_tap(ox, gin_w, out_w, x0, x1, lx);
        const float4 d = sample_d(d_img, gin_w, y0, y1, ly, x0, x1, lx, lane);
        float4 f;
        f.x = d.x * g.x, f.y = d.y * g.y, f.z = d.z * g.z, f.w = d.w * g.w;
        if (fhat_out != nullptr) reinterpret_cast<float4*>(fhat_out + ((size_t)b * npix + pix) * 128)[lane] = f;
        const float ax = 1.f / (1.f + __expf(-f.x * d.x)) + d.x;
        const float ay = 1.f / (1.f + __expf(-f.y * d.y)) + d.y;
        const float az = 1.f / (1.f + __expf(-f.z * d.z)) + d.z;
        const float aw = 1.f / (1.f + __expf(-f.w * d.w)) + d.w;
        float s = wh.x * ax + wh.y * ay + wh.z * az + wh.w * aw;
        // reduce inside each half-warp (lanes 0-15: fg branch, 16-31: bg branch)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) fg[(size_t)b * npix + pix] = s + bias_fg;
        if (lane == 16 && bg != nullptr) bg[(size_t)b * npix + pix] = s + bias_bg;
    }
}

// Orthogonality loss pieces from the normalised features fhat [B, npix, 128] (first 64 = branch 1):
//   gram[b, 0|1, 64, 64] = F_k^T F_k ; diag[b] = sum_i (f1_i . f2_i)^2
// Stage 1: one CTA per (pixel chunk, image) computes both 64x64 Grams of its chunk with 4x4 register tiles (256 threads
// = 16 x 16 tiles; per pixel a thread reads two float4 per branch from shared memory for 32 FMAs) and writes them to
// its own slot of `part` — no atomics.  Stage 2 sums the slots in a fixed order (bit-reproducible), stage 3 reduces
// <G1, G2>_F - diag in fp64.  (Round 1: 8-pixel tiles, 34 shared loads per 32 FMAs and 8 192 float atomics per CTA:
// 240 us for 16 images; this version: see DESIGN.md.)
constexpr int GRAM_CHUNK = 128;  // pixels per CTA
constexpr int GRAM_TILE = 32;    // pixels staged per synchronisation

__global__ void __launch_bounds__(256)
    decoder_gram_kernel(const float* __restrict__ fhat, float* __restrict__ part, float* __restrict__ dpart, int npix) {
    __shared__ __align__(16) float sf[GRAM_TILE][128];
    __shared__ float sdiag[8];
    const int b = blockIdx.y, nchunks = gridDim.x;
    const int p0 = blockIdx.x * GRAM_CHUNK;
    const int p1 = min(npix, p0 + GRAM_CHUNK);
    const int tr = (threadIdx.x >> 4) * 4, tc = (threadIdx.x & 15) * 4;  // this thread's 4x4 tile: rows tr.., cols tc..
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float g1[4][4], g2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) g1[i][j] = 0.f, g2[i][j] = 0.f;
    float dacc = 0.f;
    for (int p = p0; p < p1; p += GRAM_TILE) {
        const int n = min(GRAM_TILE, p1 - p);
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(fhat + ((size_t)b * npix + p) * 128);
        for (int i = threadIdx.x; i < n * 32; i += 256) reinterpret_cast<float4*>(&sf[0][0])[i] = src[i];
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            const float4 a1 = *reinterpret_cast<const float4*>(&sf[k][tr]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sf[k][tc]);
            const float4 a2 = *reinterpret_cast<const float4*>(&sf[k][64 + tr]);
            const float4 b2 = *reinterpret_cast<const float4*>(&sf[k][64 + tc]);
            const float av1[4] = {a1.x, a1.y, a1.z, a1.w}, bv1[4] = {b1.x, b1.y, b1.z, b1.w};
            const float av2[4] = {a2.x, a2.y, a2.z, a2.w}, bv2[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g1[i][j] = fmaf(av1[i], bv1[j], g1[i][j]);
                    g2[i][j] = fmaf(av2[i], bv2[j], g2[i][j]);
                }
        }
        for (int k = warp; k < n; k += 8) {  // per-pixel dot product f1 . f2
            float d = sf[k][lane] * sf[k][64 + lane] + sf[k][lane + 32] * sf[k][96 + lane];
            d = warp_sum(d);
            dacc += d * d;  // identical in every lane
        }
    }
    float* gp = part + ((size_t)b * nchunks + blockIdx.x) * 8192;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<float4*>(gp + (tr + i) * 64 + tc) = make_float4(g1[i][0], g1[i][1], g1[i][2], g1[i][3]);
        *reinterpret_cast<float4*>(gp + 4096 + (tr + i) * 64 + tc) = make_float4(g2[i][0], g2[i][1], g2[i][2], g2[i][3]);
    }
    if (lane == 0) sdiag[warp] = dacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sdiag[w];
        dpart[(size_t)b * nchunks + blockIdx.x] = t;
    }
}

// gram[b, e] = sum over chunks (fixed order) ; prod[b, blk] = partial <G1, G2>_F of this block's elements (fp64)
__global__ void __launch_bounds__(256)
    decoder_gram_reduce_kernel(const float* __restrict__ part, float* __restrict__ gram, double* __restrict__ prod,
                               int nchunks) {
    __shared__ double red[8];
    const int b = blockIdx.y;
    const int e = blockIdx.x * 256 + threadIdx.x;  // element of one 64x64 matrix (gridDim.x = 16)
    const float* pp = part + (size_t)b * nchunks * 8192;
    float s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < nchunks; ++c) {
        s1 += pp[(size_t)c * 8192 + e];
        s2 += pp[(size_t)c * 8192 + 4096 + e];
    }
    gram[(size_t)b * 8192 + e] = s1;
    gram[(size_t)b * 8192 + 4096 + e] = s2;
    double acc = (double)s1 * (double)s2;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        prod[(size_t)b * gridDim.x + blockIdx.x] = t;
    }
}

// ortho = (sum_b <G1_b, G2_b>_F - sum_b diag_b) / (B * npix^2)
__global__ void decoder_ortho_finish_kernel(const double* __restrict__ prod, int nprod, const float* __restrict__ dpart,
                                            int ndiag, float* __restrict__ out, int B, int npix) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nprod; i += blockDim.x) acc += prod[i];
    for (int i = threadIdx.x; i < ndiag; i += blockDim.x) acc -= (double)dpart[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        out[0] = (float)(t / ((double)B * (double)npix * (double)npix));
    }
}

static size_t gram_scratch_bytes(int B, int npix) {
    const size_t nchunks = (size_t)ceil_div(npix, GRAM_CHUNK);
    return (size_t)B * nchunks * 8192 * 4 + (size_t)B * nchunks * 4 + (size_t)B * 16 * 8 + 1024;
}

size_t decoder_workspace_bytes(int B, int gin_h, int gin_w, int out_h, int out_w, int want_ortho) {
    size_t n = (size_t)B * gin_h * gin_w * 128 * 4;  // d_in
    n += (size_t)B * 128 * 4;                        // sumsq
    if (want_ortho) {
        n += (size_t)B * out_h * out_w * 128 * 4;    // fhat
        n += (size_t)B * 8192 * 4 + (size_t)B * 4;   // gram + diag (layout shared with the backward)
        n += gram_scratch_bytes(B, out_h * out_w);   // per-chunk partial Grams / diag sums / fp64 products
    }
    return n + 4096;
}

int decoder_forward(const void* keys_bf16, int B, int gin_h, int gin_w, int out_h, int out_w, const DecoderWeights& w,
                    float* fg, float* bg, float* ortho, void* workspace, size_t ws_bytes, cudaStream_t stream,
                    const int* batch_dev) {
    UCOD_REQUIRE(keys_bf16 && fg && workspace, "decoder_forward: null argument");
    UCOD_REQUIRE(batch_dev == nullptr || ortho == nullptr, "decoder_forward: device-side batch count is eval-only");
    UCOD_REQUIRE(B > 0 && gin_h > 0 && gin_w > 0 && out_h > 0 && out_w > 0, "decoder_forward: bad geometry");
    const size_t need = decoder_workspace_bytes(B, gin_h, gin_w, out_h, out_w, ortho != nullptr);
    UCOD_REQUIRE(ws_bytes >= need, "decoder_forward: workspace too small (%zu < %zu)", ws_bytes, need);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    float* d_in = reinterpret_cast<float*>(base);
    size_t off = (size_t)B * gin_h * gin_w * 128 * 4;
    float* sumsq = reinterpret_cast<float*>(base + off);
    off += (size_t)B * 128 * 4;
    float *fhat = nullptr, *gram = nullptr, *diag = nullptr;
    const int npix = out_h * out_w;
    if (ortho) {
        fhat = reinterpret_cast<float*>(base + off);
        off += (size_t)B * npix * 128 * 4;
        gram = reinterpret_cast<float*>(base + off);
        off += (size_t)B * 8192 * 4;
        diag = reinterpret_cast<float*>(base + off);
    }
    GemmEpi ep;
    ep.mode = EPI_BIAS_F32;
    ep.bias = w.b_dec;
    ep.out = d_in;
    ep.ld_out = 128;
    ep.m_dev = batch_dev, ep.m_per = gin_h * gin_w;
    if (int rc = launch_gemm_bf16(keys_bf16, w.dim, w.w_dec, w.dim, B * gin_h * gin_w, 128, w.dim, ep, stream))
        return rc;
    UCOD_CHECK_CUDA(cudaMemsetAsync(sumsq, 0, (size_t)B * 128 * 4, stream));
    dim3 grid(ceil_div(npix, DEC_PIX_PER_BLOCK), B);
    const double d_bytes = (double)B * gin_h * gin_w * 128 * 4;
    {
        ProfScope ps(KC_DECODER, stream, batch_dev ? 0.0 : d_bytes);
        decoder_sumsq_kernel<<<grid, 256, 0, stream>>>(d_in, sumsq, gin_h, gin_w, out_h, out_w, batch_dev);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    {
        ProfScope ps(KC_DECODER, stream, batch_dev ? 0.0 : d_bytes + (double)B * npix * 8);
        decoder_head_kernel<<<grid, 256, 0, stream>>>(d_in, sumsq, w.emb, w.w_fg, w.b_fg, w.w_bg, w.b_bg, fg, bg, fhat,
                                                      gin_h, gin_w, out_h, out_w, batch_dev);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    if (ortho) {
        const int nchunks = ceil_div(npix, GRAM_CHUNK);
        float* part = diag + (((size_t)B + 63) / 64) * 64;  // 256-byte aligned past the per-image diag slots
        float* dpart = part + (size_t)B * nchunks * 8192;
        double* prod = reinterpret_cast<double*>(dpart + (((size_t)B * nchunks + 1) / 2) * 2);
        dim3 g2(nchunks, B);
        {
            ProfScope ps(KC_DECODER, stream, (double)B * npix * 128 * 4);
            decoder_gram_kernel<<<g2, 256, 0, stream>>>(fhat, part, dpart, npix);
        }
        UCOD_CHECK_CUDA(cudaGetLastError());
        {
            ProfScope ps(KC_DECODER, stream, (double)B * nchunks * 8192 * 4);
            decoder_gram_reduce_kernel<<<dim3(16, B), 256, 0, stream>>>(part, gram, prod, nchunks);
        }
        ProfScope ps(KC_DECODER, stream, (double)B * 16 * 8);
        decoder_ortho_finish_kernel<<<1, 256, 0, stream>>>(prod, B * 16, dpart, B * nchunks, ortho, B, npix);
        UCOD_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Feature layout conversion for the drop-in API: [B, C, H*W] fp32 with arbitrary (c, p) strides
// -> token-major bf16 [B, H*W, C].  32x32 smem-tiled transpose; coalesced on both sides for the
// contiguous-NCHW case, and for the channels-last case reads are coalesced along c directly.
// ------------------------------------------------------------------------------------------------
// Backward of the student decoder for the first-stage training step (engine/runner/loop_UCOD_DPL.py:148-184):
//   loss = BCEwL(fg, t) + BCEwL(bg, 1 - t) + ortho,  t = APM-merged pseudo label (no gradient flows into t).
// With d = U(D_in) (bilinear upsample of the 1x1-conv output), r_c = sign(e_c) / sqrt(sum_p d_pc^2), fh = d r,
// u = fh d, a = sigmoid(u) + d:
//   dlogit = (sigmoid(logit) - target) / (B npix)          da = dlogit * w_head
//   dfh    = da s(1-s) d + (2 / (B npix^2)) [ (Fh_k G_other)_pc - (fh1_p . fh2_p) fh_other_pc ]      (Gram identity)
//   dd     = da (1 + s(1-s) fh) + r (dfh - fh t_c),  t_c = sum_p dfh_pc fh_pc                         (normalisation)
//          = P1 - d r^2 t_c
//   dD_in  = U^T P1 - (r^2 t_c) U^T d        -> two scatter-add buffers (A = U^T P1, E = U^T d), one elementwise pass
//   dW_dec = dD_in^T X, db_dec = sum dD_in ; the gradient of learnable_embedding is identically zero
//   (F.normalize removes |e|; the reference's autograd value is rounding noise, see tests/test_oracle_train.py).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_add4(float* dst, float4 v) {
    atomicAdd(reinterpret_cast<float4*>(dst), v);
}

__global__ void __launch_bounds__(256)
    decoder_bwd_pixels_kernel(const float* __restrict__ d_in, const float* __restrict__ sumsq,
                              const float* __restrict__ emb, const float* __restrict__ w_fg,
                              const float* __restrict__ w_bg, const float* __restrict__ fhat,
                              const float* __restrict__ gram, const float* __restrict__ fg,
                              const float* __restrict__ bg, const float* __restrict__ target,
                              const float* __restrict__ dfg, const float* __restrict__ dbg,
                              const float* __restrict__ dortho, float* __restrict__ a_buf, float* __restrict__ e_buf, float* __restrict__ tsum,
                              float* __restrict__ g_wfg, float* __restrict__ g_bfg, float* __restrict__ g_wbg,
                              float* __restrict__ g_bbg, float* __restrict__ loss2, int B, int gin_h, int gin_w,
                              int out_h, int out_w) {
    __shared__ __align__(16) float sG[2][64 * 64];
    __shared__ __align__(16) float sf[8][128];
    __shared__ float4 red4[8][32];
    __shared__ float red1[8][32];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int branch = lane >> 4, ch0 = lane * 4, col = ch0 & 63;
    const int npix = out_h * out_w;
    for (int i = threadIdx.x; i < 2 * 4096; i += 256) (&sG[0][0])[i] = gram[(size_t)b * 8192 + i];
    __syncthreads();
    const float* d_img = d_in + (size_t)b * gin_h * gin_w * 128;
    const float4 ss = reinterpret_cast<const float4*>(sumsq + (size_t)b * 128)[lane];
    const float4 e = __ldg(reinterpret_cast<const float4*>(emb) + lane);
    const float4 wh = lane < 16 ? __ldg(reinterpret_cast<const float4*>(w_fg) + lane)
                                : __ldg(reinterpret_cast<const float4*>(w_bg) + (lane - 16));
    float4 r;
    r.x = e.x / fmaxf(fabsf(e.x) * sqrtf(ss.x), 1e-12f);
    r.y = e.y / fmaxf(fabsf(e.y) * sqrtf(ss.y), 1e-12f);
    r.z = e.z / fmaxf(fabsf(e.z) * sqrtf(ss.z), 1e-12f);
    r.w = e.w / fmaxf(fabsf(e.w) * sqrtf(ss.w), 1e-12f);
    const float inv_n = 1.0f / ((float)B * (float)npix);
    const float oc = 2.0f / ((float)B * (float)npix * (float)npix) * (dortho != nullptr ? __ldg(dortho) : 1.f);
    const float* Gother = sG[branch ^ 1];
    float4 gw = make_float4(0.f, 0.f, 0.f, 0.f), tacc = make_float4(0.f, 0.f, 0.f, 0.f);
    float gb = 0.f, lacc = 0.f;

    for (int i = 0; i < DEC_PIX_PER_BLOCK / 8; ++i) {
        const int pix = blockIdx.x * DEC_PIX_PER_BLOCK + i * 8 + warp;
        if (pix >= npix) break;  // warp-uniform
        const int oy = pix / out_w, ox = pix - oy * out_w;
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_tap(oy, gin_h, out_h, y0, y1, ly);
        bilinear_tap(ox, gin_w, out_w, x0, x1, lx);
        const float4 d = sample_d(d_img, gin_w, y0, y1, ly, x0, x1, lx, lane);
        const float4 f = reinterpret_cast<const float4*>(fhat + ((size_t)b * npix + pix) * 128)[lane];
        __syncwarp();
        reinterpret_cast<float4*>(sf[warp])[lane] = f;
        __syncwarp();
        const float4 fo = reinterpret_cast<const float4*>(sf[warp])[lane ^ 16];  // other branch, same column
        const float sp = 0.5f * warp_sum(f.x * fo.x + f.y * fo.y + f.z * fo.z + f.w * fo.w);
        float4 O = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* own = sf[warp] + branch * 64;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) {
            const float fc = own[c];
            const float4 gr = *reinterpret_cast<const float4*>(Gother + c * 64 + col);
            O.x += fc * gr.x, O.y += fc * gr.y, O.z += fc * gr.z, O.w += fc * gr.w;
        }
        O.x -= sp * fo.x, O.y -= sp * fo.y, O.z -= sp * fo.z, O.w -= sp * fo.w;
        const float logit = branch == 0 ? fg[(size_t)b * npix + pix] : bg[(size_t)b * npix + pix];
        float dlog;
        if (dfg != nullptr) {  // upstream gradients given (autograd entry)
            dlog = branch == 0 ? dfg[(size_t)b * npix + pix] : dbg[(size_t)b * npix + pix];
        } else {               // fused BCE-with-logits against the merged pseudo label
            const float t0 = target[(size_t)b * npix + pix];
            const float tt = branch == 0 ? t0 : 1.f - t0;
            dlog = (1.f / (1.f + expf(-logit)) - tt) * inv_n;
            if ((lane & 15) == 0) lacc += fmaxf(logit, 0.f) - logit * tt + log1pf(expf(-fabsf(logit)));
        }
        float4 sg, a, da, dfh, p1;
        sg.x = 1.f / (1.f + expf(-f.x * d.x)), sg.y = 1.f / (1.f + expf(-f.y * d.y));
        sg.z = 1.f / (1.f + expf(-f.z * d.z)), sg.w = 1.f / (1.f + expf(-f.w * d.w));
        a.x = sg.x + d.x, a.y = sg.y + d.y, a.z = sg.z + d.z, a.w = sg.w + d.w;
        da.x = dlog * wh.x, da.y = dlog * wh.y, da.z = dlog * wh.z, da.w = dlog * wh.w;
        gw.x += dlog * a.x, gw.y += dlog * a.y, gw.z += dlog * a.z, gw.w += dlog * a.w;
        if ((lane & 15) == 0) gb += dlog;
        const float4 ds = make_float4(sg.x * (1.f - sg.x), sg.y * (1.f - sg.y), sg.z * (1.f - sg.z), sg.w * (1.f - sg.w));
        dfh.x = da.x * ds.x * d.x + oc * O.x, dfh.y = da.y * ds.y * d.y + oc * O.y;
        dfh.z = da.z * ds.z * d.z + oc * O.z, dfh.w = da.w * ds.w * d.w + oc * O.w;
        tacc.x += dfh.x * f.x, tacc.y += dfh.y * f.y, tacc.z += dfh.z * f.z, tacc.w += dfh.w * f.w;
        p1.x = da.x * (1.f + ds.x * f.x) + r.x * dfh.x, p1.y = da.y * (1.f + ds.y * f.y) + r.y * dfh.y;
        p1.z = da.z * (1.f + ds.z * f.z) + r.z * dfh.z, p1.w = da.w * (1.f + ds.w * f.w) + r.w * dfh.w;
        // adjoint of the bilinear upsample: scatter to the four source pixels
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        const size_t base = (size_t)b * gin_h * gin_w;
        const size_t q00 = (base + (size_t)y0 * gin_w + x0) * 128 + ch0, q01 = (base + (size_t)y0 * gin_w + x1) * 128 + ch0;
        const size_t q10 = (base + (size_t)y1 * gin_w + x0) * 128 + ch0, q11 = (base + (size_t)y1 * gin_w + x1) * 128 + ch0;
        atomic_add4(a_buf + q00, make_float4(p1.x * w00, p1.y * w00, p1.z * w00, p1.w * w00));
        atomic_add4(a_buf + q01, make_float4(p1.x * w01, p1.y * w01, p1.z * w01, p1.w * w01));
        atomic_add4(a_buf + q10, make_float4(p1.x * w10, p1.y * w10, p1.z * w10, p1.w * w10));
        atomic_add4(a_buf + q11, make_float4(p1.x * w11, p1.y * w11, p1.z * w11, p1.w * w11));
        atomic_add4(e_buf + q00, make_float4(d.x * w00, d.y * w00, d.z * w00, d.w * w00));
        atomic_add4(e_buf + q01, make_float4(d.x * w01, d.y * w01, d.z * w01, d.w * w01));
        atomic_add4(e_buf + q10, make_float4(d.x * w10, d.y * w10, d.z * w10, d.w * w10));
        atomic_add4(e_buf + q11, make_float4(d.x * w11, d.y * w11, d.z * w11, d.w * w11));
    }
    // block reduction of the per-lane accumulators, then one atomic per channel / scalar
    red4[warp][lane] = gw;
    __syncthreads();
    if (warp == 0) {
        float4 t = red4[0][lane];
        for (int w = 1; w < 8; ++w) t.x += red4[w][lane].x, t.y += red4[w][lane].y, t.z += red4[w][lane].z, t.w += red4[w][lane].w;
        float* dst = lane < 16 ? g_wfg + ch0 : g_wbg + (ch0 - 64);
        atomicAdd(dst + 0, t.x), atomicAdd(dst + 1, t.y), atomicAdd(dst + 2, t.z), atomicAdd(dst + 3, t.w);
    }
    __syncthreads();
    red4[warp][lane] = tacc;
    red1[warp][lane] = (lane & 15) == 0 ? gb : 0.f;
    __syncthreads();
    if (warp == 0) {
        float4 t = red4[0][lane];
        float g1 = red1[0][lane];
        for (int w = 1; w < 8; ++w) {
            t.x += red4[w][lane].x, t.y += red4[w][lane].y, t.z += red4[w][lane].z, t.w += red4[w][lane].w;
            g1 += red1[w][lane];
        }
        float* dst = tsum + (size_t)b * 128 + ch0;
        atomicAdd(dst + 0, t.x), atomicAdd(dst + 1, t.y), atomicAdd(dst + 2, t.z), atomicAdd(dst + 3, t.w);
        if (lane == 0) atomicAdd(g_bfg, g1);
        if (lane == 16) atomicAdd(g_bbg, g1);
    }
    __syncthreads();
    red1[warp][lane] = lacc;
    __syncthreads();
    if (warp == 0 && (lane & 15) == 0) {
        float l = 0.f;
        for (int w = 0; w < 8; ++w) l += red1[w][lane];
        atomicAdd(loss2 + (lane >> 4), l * inv_n);
    }
}

// five small accumulators of the backward (head-weight / bias gradients, the two BCE sums) zeroed by one launch
__global__ void decoder_bwd_zero_kernel(float* w_fg, float* w_bg, float* b_fg, float* b_bg, float* loss2) {
    const int t = threadIdx.x;
    if (t < 64) w_fg[t] = 0.f, w_bg[t] = 0.f;
    if (t == 0) b_fg[0] = 0.f, b_bg[0] = 0.f, loss2[0] = 0.f, loss2[1] = 0.f;
}

static size_t bwd_scatter_bytes(size_t rows, int B) {  // A, E scatter buffers + per-image channel sums, 256-aligned
    return ((rows * 128 * 4 * 2 + (size_t)B * 128 * 4) + 255) / 256 * 256;
}

size_t decoder_backward_workspace_bytes(int B, int gin_h, int gin_w) {
    const size_t rows = (size_t)B * gin_h * gin_w;
    // A, E, tsum | dDt (bf16, transposed), split-K partial tiles, column-sum partials (wgrad.cu); dim <= 1024
    size_t wg = 0;
    for (int dim = 256; dim <= 1024; dim += 256) {
        const size_t n = wgrad_workspace_bytes((int)rows, dim);
        wg = n > wg ? n : wg;
    }
    return bwd_scatter_bytes(rows, B) + wg + 4096;
}

int decoder_backward(const void* keys_bf16, int B, int gin_h, int gin_w, int out_h, int out_w, const DecoderWeights& w,
                     const float* fg, const float* bg, const float* target, const float* dfg, const float* dbg,
                     const float* dortho, void* fwd_workspace, size_t fwd_ws_bytes, const DecoderGrads& g, float* loss2,
                     void* workspace, size_t ws_bytes, cudaStream_t stream) {
    UCOD_REQUIRE(keys_bf16 && fg && bg && fwd_workspace && workspace && loss2, "decoder_backward: null argument");
    UCOD_REQUIRE(target != nullptr || (dfg != nullptr && dbg != nullptr),
                 "decoder_backward: give either the BCE target or the upstream gradients dfg/dbg");
    UCOD_REQUIRE(g.w_dec && g.b_dec && g.w_fg && g.b_fg && g.w_bg && g.b_bg, "decoder_backward: null gradient pointer");
    UCOD_REQUIRE(w.dim % 64 == 0, "decoder_backward: dim must be a multiple of 64");
    UCOD_REQUIRE(fwd_ws_bytes >= decoder_workspace_bytes(B, gin_h, gin_w, out_h, out_w, 1),
                 "decoder_backward: the forward workspace must come from a forward with the ortho loss enabled");
    UCOD_REQUIRE(ws_bytes >= decoder_backward_workspace_bytes(B, gin_h, gin_w), "decoder_backward: workspace too small");
    const int npix = out_h * out_w;
    const size_t rows = (size_t)B * gin_h * gin_w;
    // forward workspace layout (see decoder_forward)
    uint8_t* fb = static_cast<uint8_t*>(fwd_workspace);
    const float* d_in = reinterpret_cast<const float*>(fb);
    size_t off = rows * 128 * 4;
    const float* sumsq = reinterpret_cast<const float*>(fb + off);
    off += (size_t)B * 128 * 4;
    const float* fhat = reinterpret_cast<const float*>(fb + off);
    off += (size_t)B * npix * 128 * 4;
    const float* gram = reinterpret_cast<const float*>(fb + off);
    UCOD_REQUIRE(w.dim <= 1024 && w.dim % 256 == 0, "decoder_backward: dim must be a multiple of 256 (<= 1024)");
    UCOD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "decoder_backward: workspace must be 256-byte aligned");
    float* a_buf = static_cast<float*>(workspace);
    float* e_buf = a_buf + rows * 128;
    float* tsum = e_buf + rows * 128;
    uint8_t* wg_ws = static_cast<uint8_t*>(workspace) + bwd_scatter_bytes(rows, B);
    const size_t wg_bytes = ws_bytes - bwd_scatter_bytes(rows, B);
    // one memset for the scatter targets (A, E, tsum are contiguous), one launch for the five small accumulators; the
    // weight / bias gradient of the 1x1 conv is written (not accumulated) by the split-K reduction
    UCOD_CHECK_CUDA(cudaMemsetAsync(a_buf, 0, rows * 128 * 4 * 2 + (size_t)B * 128 * 4, stream));
    decoder_bwd_zero_kernel<<<1, 64, 0, stream>>>(g.w_fg, g.w_bg, g.b_fg, g.b_bg, loss2);
    {
        dim3 grid(ceil_div(npix, DEC_PIX_PER_BLOCK), B);
        ProfScope ps(KC_DECODER, stream, (double)B * npix * 128 * 4 * 2 + (double)rows * 128 * 4 * 3);
        decoder_bwd_pixels_kernel<<<grid, 256, 0, stream>>>(d_in, sumsq, w.emb, w.w_fg, w.w_bg, fhat, gram, fg, bg,
                                                            target, dfg, dbg, dortho, a_buf, e_buf, tsum, g.w_fg,
                                                            g.b_fg, g.w_bg, g.b_bg, loss2, B, gin_h, gin_w, out_h, out_w);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return wgrad_tensor_core(a_buf, e_buf, tsum, sumsq, w.emb, keys_bf16, gin_h * gin_w, (int)rows, w.dim, g.w_dec,
                             g.b_dec, wg_ws, wg_bytes, stream);
}

// Fused AdamW (torch semantics, decoupled weight decay) + EMA of the updated parameters
// (engine/runner/runner.py:282-285, loop_UCOD_DPL.py:186-191) over flat fp32 buffers.
__global__ void adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ ema, size_t n, float lr, float b1, float b2,
                                 float eps, float wd, float bc1, float bc2_sqrt, float grad_scale, float ema_alpha) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
        p[i] = pi, m[i] = mi, v[i] = vi;
        if (ema != nullptr) ema[i] = ema_alpha * ema[i] + (1.f - ema_alpha) * pi;
    }
}

int adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, size_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step_t, float grad_scale, float ema_alpha,
                   cudaStream_t stream) {
    UCOD_REQUIRE(p && g && m && v && n > 0 && step_t >= 1, "adamw_ema_step: bad argument");
    const float bc1 = 1.f - powf(beta1, (float)step_t);
    const float bc2s = sqrtf(1.f - powf(beta2, (float)step_t));
    const unsigned grid = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    ProfScope ps(KC_OTHER, stream, (double)n * 4 * 9);
    adamw_ema_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s,
                                               grad_scale, ema_alpha);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void features_to_tokens_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int P,
                                          long long sb, long long sc, long long sp) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const float* src = in + (size_t)b * sb;
    if (sp == 1) {  // pixel index contiguous: read rows of p, transpose through smem
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int c = c0 + i, p = p0 + threadIdx.x;
            tile[i][threadIdx.x] = (c < C && p < P) ? src[(size_t)c * sc + p] : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int p = p0 + i, c = c0 + threadIdx.x;
            if (p < P && c < C) out[((size_t)b * P + p) * C + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
        }
    } else {  // generic / channels-last: read along c
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int p = p0 + i, c = c0 + threadIdx.x;
            if (p < P && c < C)
                out[((size_t)b * P + p) * C + c] = __float2bfloat16_rn(src[(size_t)c * sc + (size_t)p * sp]);
        }
    }
}

int features_to_tokens_bf16(const float* in, void* out, int B, int C, int P, long long sb, long long sc, long long sp,
                            cudaStream_t stream) {
    UCOD_REQUIRE(in && out && B > 0 && C > 0 && P > 0, "features_to_tokens: bad argument");
    dim3 grid(ceil_div(P, 32), ceil_div(C, 32), B), block(32, 8);
    ProfScope ps(KC_DECODER, stream, (double)B * C * P * 6);
    features_to_tokens_kernel<<<grid, block, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), C, P, sb, sc, sp);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Bilinear upsample of [B,in_h,in_w] fp32 (align_corners=False).  MODE 0: write fp32 ; MODE 1: write the
// binarised u8 mask `sigmoid(x) > 0.5` of loop_UCOD_DPL.py:356-361 without materialising the fp32 map
// (fp32 sigmoid(x) > 0.5  <=>  x > 1.5 * 2^-24, probed on torch CPU; pinned in tests/test_oracle_looktwice.py).
// One thread = 4 consecutive output pixels (32-bit store of 4 mask bytes / float4 store).
// ------------------------------------------------------------------------------------------------
#define UCOD_SIGMOID_HALF_THRESHOLD 0x1.8p-24f

template <int MODE>
__global__ void upsample_bilinear_kernel(const float* __restrict__ in, void* __restrict__ out, int in_h, int in_w,
                                         int out_h, int out_w) {
    const int b = blockIdx.z, oy = blockIdx.y;
    const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (ox0 >= out_w) return;
    const float* src = in + (size_t)b * in_h * in_w;
    int y0, y1;
    float ly;
    bilinear_tap(oy, in_h, out_h, y0, y1, ly);
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ox = ox0 + i;
        int x0, x1;
        float lx;
        bilinear_tap(ox < out_w ? ox : out_w - 1, in_w, out_w, x0, x1, lx);
        const float hx = 1.f - lx, hy = 1.f - ly;
        float t00 = __ldg(src + y0 * in_w + x0), t01 = __ldg(src + y0 * in_w + x1);
        float t10 = __ldg(src + y1 * in_w + x0), t11 = __ldg(src + y1 * in_w + x1);
        if constexpr (MODE == 2) {  // probabilities first, then interpolate (loop_CORAL.py:331-338)
            t00 = 1.f / (1.f + expf(-t00)), t01 = 1.f / (1.f + expf(-t01));
            t10 = 1.f / (1.f + expf(-t10)), t11 = 1.f / (1.f + expf(-t11));
        }
        v[i] = hy * (hx * t00 + lx * t01) + ly * (hx * t10 + lx * t11);
    }
    const size_t o = ((size_t)b * out_h + oy) * out_w + ox0;
    if constexpr (MODE == 0) {
        float* dst = static_cast<float*>(out) + o;
        for (int i = 0; i < 4 && ox0 + i < out_w; ++i) dst[i] = v[i];
    } else {
        uint8_t* dst = static_cast<uint8_t*>(out) + o;
        for (int i = 0; i < 4 && ox0 + i < out_w; ++i)
            dst[i] = (MODE == 1 ? v[i] > UCOD_SIGMOID_HALF_THRESHOLD : v[i] > 0.5f) ? 1 : 0;
    }
}

int upsample_bilinear(const float* in, void* out, int B, int in_h, int in_w, int out_h, int out_w, int binarize,
                      cudaStream_t stream) {
    UCOD_REQUIRE(in && out && B > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "upsample: bad argument");
    dim3 block(128), grid(ceil_div(ceil_div(out_w, 4), 128), out_h, B);
    ProfScope ps(KC_RESAMPLE, stream, (double)B * in_h * in_w * 4 + (double)B * out_h * out_w * (binarize ? 1 : 4));
    UCOD_REQUIRE(binarize >= 0 && binarize <= 3, "upsample: binarize mode %d unknown", binarize);
    if (binarize == 1)
        upsample_bilinear_kernel<1><<<grid, block, 0, stream>>>(in, out, in_h, in_w, out_h, out_w);
    else if (binarize == 2)
        upsample_bilinear_kernel<2><<<grid, block, 0, stream>>>(in, out, in_h, in_w, out_h, out_w);
    else if (binarize == 3)
        upsample_bilinear_kernel<3><<<grid, block, 0, stream>>>(in, out, in_h, in_w, out_h, out_w);
    else
        upsample_bilinear_kernel<0><<<grid, block, 0, stream>>>(in, out, in_h, in_w, out_h, out_w);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
