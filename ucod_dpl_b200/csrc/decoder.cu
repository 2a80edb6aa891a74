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
                         int out_w) {
    __shared__ float4 part[8][32];
    const int b = blockIdx.y;
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
                        float* __restrict__ fhat_out, int gin_h, int gin_w, int out_h, int out_w) {
    const int b = blockIdx.y;
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
//   gram[b, 0|1, 64, 64] += F_k^T F_k over a pixel chunk ; diag[b] += sum_i (f1_i . f2_i)^2
__global__ void __launch_bounds__(256)
    decoder_gram_kernel(const float* __restrict__ fhat, float* __restrict__ gram, float* __restrict__ diag, int npix,
                        int chunk) {
    __shared__ float sf[8][128];
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * chunk;
    const int p1 = min(npix, p0 + chunk);
    // thread t owns entries (r, c0..c0+15) of both Grams: r = t / 4, c0 = (t % 4) * 16
    const int r = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16;
    float g1[16], g2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) g1[i] = 0.f, g2[i] = 0.f;
    float dacc = 0.f;
    for (int p = p0; p < p1; p += 8) {
        const int n = min(8, p1 - p);
        __syncthreads();
        for (int i = threadIdx.x; i < n * 128; i += 256) sf[i >> 7][i & 127] = fhat[((size_t)b * npix + p) * 128 + i];
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            const float a1 = sf[k][r], a2 = sf[k][64 + r];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                g1[i] += a1 * sf[k][c0 + i];
                g2[i] += a2 * sf[k][64 + c0 + i];
            }
        }
        if (threadIdx.x < n * 32) {  // warp k handles pixel k's dot product f1.f2
            const int k = threadIdx.x >> 5, l = threadIdx.x & 31;
            float d = sf[k][l] * sf[k][64 + l] + sf[k][l + 32] * sf[k][96 + l];
            d = warp_sum(d);
            if (l == 0) dacc += d * d;
        }
    }
    float* gb = gram + (size_t)b * 2 * 4096;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        atomicAdd(gb + r * 64 + c0 + i, g1[i]);
        atomicAdd(gb + 4096 + r * 64 + c0 + i, g2[i]);
    }
    if ((threadIdx.x & 31) == 0 && dacc != 0.f) atomicAdd(diag + b, dacc);
}

// ortho = (sum_b <G1_b, G2_b>_F - sum_b diag_b) / (B * npix^2)
__global__ void decoder_ortho_finish_kernel(const float* __restrict__ gram, const float* __restrict__ diag,
                                            float* __restrict__ out, int B, int npix) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B * 4096; i += blockDim.x) {
        const int b = i >> 12, e = i & 4095;
        acc += (double)gram[(size_t)b * 8192 + e] * (double)gram[(size_t)b * 8192 + 4096 + e];
    }
    for (int b = threadIdx.x; b < B; b += blockDim.x) acc -= (double)diag[b];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        out[0] = (float)(t / ((double)B * (double)npix * (double)npix));
    }
}

size_t decoder_workspace_bytes(int B, int gin_h, int gin_w, int out_h, int out_w, int want_ortho) {
    size_t n = (size_t)B * gin_h * gin_w * 128 * 4;  // d_in
    n += (size_t)B * 128 * 4;                        // sumsq
    if (want_ortho) {
        n += (size_t)B * out_h * out_w * 128 * 4;    // fhat
        n += (size_t)B * 8192 * 4 + (size_t)B * 4;   // gram + diag
    }
    return n + 4096;
}

int decoder_forward(const void* keys_bf16, int B, int gin_h, int gin_w, int out_h, int out_w, const DecoderWeights& w,
                    float* fg, float* bg, float* ortho, void* workspace, size_t ws_bytes, cudaStream_t stream) {
    UCOD_REQUIRE(keys_bf16 && fg && workspace, "decoder_forward: null argument");
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
    if (int rc = launch_gemm_bf16(keys_bf16, w.dim, w.w_dec, w.dim, B * gin_h * gin_w, 128, w.dim, ep, stream))
        return rc;
    UCOD_CHECK_CUDA(cudaMemsetAsync(sumsq, 0, (size_t)B * 128 * 4, stream));
    dim3 grid(ceil_div(npix, DEC_PIX_PER_BLOCK), B);
    const double d_bytes = (double)B * gin_h * gin_w * 128 * 4;
    {
        ProfScope ps(KC_DECODER, stream, d_bytes);
        decoder_sumsq_kernel<<<grid, 256, 0, stream>>>(d_in, sumsq, gin_h, gin_w, out_h, out_w);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    {
        ProfScope ps(KC_DECODER, stream, d_bytes + (double)B * npix * 8);
        decoder_head_kernel<<<grid, 256, 0, stream>>>(d_in, sumsq, w.emb, w.w_fg, w.b_fg, w.w_bg, w.b_bg, fg, bg, fhat,
                                                      gin_h, gin_w, out_h, out_w);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    if (ortho) {
        UCOD_CHECK_CUDA(cudaMemsetAsync(gram, 0, (size_t)B * 8192 * 4 + (size_t)B * 4, stream));
        const int chunk = 256;
        dim3 g2(ceil_div(npix, chunk), B);
        {
            ProfScope ps(KC_DECODER, stream, (double)B * npix * 128 * 4);
            decoder_gram_kernel<<<g2, 256, 0, stream>>>(fhat, gram, diag, npix, chunk);
        }
        UCOD_CHECK_CUDA(cudaGetLastError());
        ProfScope ps(KC_DECODER, stream, (double)B * 8192 * 4);
        decoder_ortho_finish_kernel<<<1, 256, 0, stream>>>(gram, diag, ortho, B, npix);
        UCOD_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Feature layout conversion for the drop-in API: [B, C, H*W] fp32 with arbitrary (c, p) strides
// -> token-major bf16 [B, H*W, C].  32x32 smem-tiled transpose; coalesced on both sides for the
// contiguous-NCHW case, and for the channels-last case reads are coalesced along c directly.
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
