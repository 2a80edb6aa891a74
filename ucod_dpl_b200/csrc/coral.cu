// CORAL second stage (SparseRefiner, eval): the memory-bound pieces around the tcgen05 GEMM / attention kernels.
//
//   coral_entropy_select   — EntropySelector.forward (models/modules/ASR.py:41-51): p = preds | sigmoid(preds)
//                            (branch on "all values in [0,1]" over the whole call, like the reference),
//                            e = -p log max(p,1e-5), adaptive average pool to w x w, mask = score > threshold
//   coral_window_head      — CSF tail (models/modules/CSF.py:41-42): depthwise 7x7 conv (pad 3) followed by the
//                            1x1 mask_dec.  Both are linear, so they are folded into one 768 -> 49 tap projection
//                            (a tcgen05 GEMM, done by the caller) and this 49-tap gather-sum over the window
//   coral_scatter_windows  — HRE.concate_windows (models/modules/HRE.py:18-39): windows never overlap, so the
//                            canvas is window / (1 + 1e-6) inside selected cells and exactly 0 elsewhere
//   coral_gated_ensemble   — GatedEnsembler.forward (models/modules/GE_pix_level.py:16-26)
//   layernorm_rows_bf16 / cast / features_to_tokens_f32 / resize_tokens_bilinear — layout + normalisation helpers
//   (CrossAttentionBlock's LayerNorms, models/modules/mlp.py:134-148; loop_CORAL.py:224-227 feature resize)
#include "coral.cuh"

#include "prof.cuh"

namespace ucod {

// ------------------------------------------------------------------------------------------------
// flag[stride * blockIdx.y] |= 1 when any of that slice's n values lies outside [0, 1] (stride 0: one flag per call)
__global__ void range_flag_kernel(const float* __restrict__ x, size_t n, int* flag, int stride) {
    const float* xs = x + (size_t)blockIdx.y * n;
    bool bad = false;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = xs[i];
        bad |= !(v >= 0.f && v <= 1.f);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag + (size_t)stride * blockIdx.y, 1);
}

// one CTA per image; bins of adaptive_avg_pool2d: [floor(i*n/w), ceil((i+1)*n/w))
__global__ void entropy_select_kernel(const float* __restrict__ preds, int P, int ws, float threshold,
                                      const int* __restrict__ flag, int flag_stride, float* __restrict__ entropy,
                                      float* __restrict__ scores, uint8_t* __restrict__ mask) {
    extern __shared__ float s_bins[];  // ws*ws partial sums
    const int b = blockIdx.x;
    const bool logits = (flag[(size_t)flag_stride * b] != 0);
    const int nb = ws * ws;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_bins[i] = 0.f;
    __syncthreads();
    const float* src = preds + (size_t)b * P * P;
    float* dst = entropy + (size_t)b * P * P;
    for (int wy = 0; wy < ws; ++wy) {
        const int y0 = (wy * P) / ws, y1 = ((wy + 1) * P + ws - 1) / ws;
        for (int wx = 0; wx < ws; ++wx) {
            const int x0 = (wx * P) / ws, x1 = ((wx + 1) * P + ws - 1) / ws;
            const int bw = x1 - x0, cnt = (y1 - y0) * bw;
            float acc = 0.f;
            for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
                const int y = y0 + i / bw, x = x0 + i % bw;
                const float v = src[y * P + x];
                const float p = logits ? 1.f / (1.f + expf(-v)) : v;
                const float e = -p * logf(fmaxf(p, 1e-5f));
                dst[y * P + x] = e;  // overlapping bins rewrite the same value
                acc += e;
            }
            acc = warp_sum(acc);
            if ((threadIdx.x & 31) == 0) atomicAdd(&s_bins[wy * ws + wx], acc);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const int wy = i / ws, wx = i % ws;
        const int y0 = (wy * P) / ws, y1 = ((wy + 1) * P + ws - 1) / ws;
        const int x0 = (wx * P) / ws, x1 = ((wx + 1) * P + ws - 1) / ws;
        const float sc = s_bins[i] / (float)((y1 - y0) * (x1 - x0));
        scores[b * nb + i] = sc;
        mask[b * nb + i] = sc > threshold ? 1 : 0;
    }
}

int coral_entropy_select(const float* preds, int B, int P, int ws, float threshold, float* entropy, float* scores,
                         uint8_t* mask, int* flag, cudaStream_t stream, int per_image) {
    UCOD_REQUIRE(preds && entropy && scores && mask && flag, "coral_entropy_select: null pointer");
    UCOD_REQUIRE(B > 0 && P > 0 && ws > 0 && ws <= 16 && ws <= P, "coral_entropy_select: bad geometry");
    // `preds if all in [0,1] else sigmoid(preds)` (ASR.py:42): the reference decides per CALL and evaluates at batch 1,
    // i.e. per image; per_image != 0 keeps that when several images share a launch (flag: B ints instead of one)
    UCOD_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * (per_image ? B : 1), stream));
    const size_t n = (size_t)B * P * P;
    {
        ProfScope ps(KC_OTHER, stream, (double)n * 4);
        const size_t per = per_image ? (size_t)P * P : n;
        const unsigned gx = (unsigned)((per + 255) / 256 < 1024 ? (per + 255) / 256 : 1024);
        range_flag_kernel<<<dim3(gx, per_image ? B : 1), 256, 0, stream>>>(preds, per, flag, per_image ? 1 : 0);
    }
    {
        ProfScope ps(KC_OTHER, stream, (double)n * 8);
        entropy_select_kernel<<<B, 256, ws * ws * sizeof(float), stream>>>(preds, P, ws, threshold, flag,
                                                                           per_image ? 1 : 0, entropy, scores, mask);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// out[n,y,x] = bias_const + sum_{ky,kx} taps[n*g*g + (y+ky-3)*g + (x+kx-3)][ky*7+kx]   (zero padding)
__global__ void window_head_kernel(const float* __restrict__ taps, int ld, int g, float bias_const,
                                   float* __restrict__ out) {
    const int n = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g * g) return;
    const int y = idx / g, x = idx - y * g;
    const float* base = taps + (size_t)n * g * g * ld;
    float acc = bias_const;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + ky - 3;
        if (yy < 0 || yy >= g) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
            const int xx = x + kx - 3;
            if (xx < 0 || xx >= g) continue;
            acc += __ldg(base + (size_t)(yy * g + xx) * ld + ky * 7 + kx);
        }
    }
    out[(size_t)n * g * g + idx] = acc;
}

int coral_window_head(const float* taps, int ld_taps, int n_windows, int g, float bias_const, float* out,
                      cudaStream_t stream) {
    UCOD_REQUIRE(taps && out && n_windows > 0 && g > 0 && ld_taps >= 49, "coral_window_head: bad argument");
    dim3 grid(ceil_div(g * g, 128), n_windows);
    ProfScope ps(KC_OTHER, stream, (double)n_windows * g * g * (49 + 1) * 4);
    window_head_kernel<<<grid, 128, 0, stream>>>(taps, ld_taps, g, bias_const, out);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void scatter_windows_kernel(const float* __restrict__ win, const int* __restrict__ slot, int ws, int g,
                                       float* __restrict__ out) {
    const int b = blockIdx.y, S = ws * g;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * S) return;
    const int y = idx / S, x = idx - y * S;
    const int cell = (y / g) * ws + (x / g);
    const int s = slot[b * ws * ws + cell];
    float v = 0.f;
    if (s >= 0) v = win[(size_t)s * g * g + (y % g) * g + (x % g)] / (1.0f + 1e-6f);
    out[(size_t)b * S * S + idx] = v;
}

int coral_scatter_windows(const float* window_preds, const int* slot_of_cell, int B, int ws, int g, float* out,
                          cudaStream_t stream) {
    UCOD_REQUIRE(slot_of_cell && out && B > 0 && ws > 0 && g > 0, "coral_scatter_windows: bad argument");
    const int S = ws * g;
    dim3 grid(ceil_div(S * S, 256), B);
    ProfScope ps(KC_OTHER, stream, (double)B * S * S * 8);
    scatter_windows_kernel<<<grid, 256, 0, stream>>>(window_preds, slot_of_cell, ws, g, out);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Gated ensemble.  ws layout: l1 [B,S,S] | prob [B,S,S] | en [B,S,S] | sums [B] | enmax (uint bits) [B]
// (enmax slot = image index when the maximum is per image, slot 0 when it is over the whole call)
__device__ __forceinline__ void bilin_tap(int dst, int in, int out, int& i0, int& i1, float& l) {
    const float scale = (float)in / (float)out;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l = src - (float)i0;
}
__global__ void ge_upsample_kernel(const float* __restrict__ preds, int P, int S, float* __restrict__ l1,
                                   float* __restrict__ prob, float* __restrict__ sums) {
    const int b = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float p = 0.f;
    if (idx < S * S) {
        const int y = idx / S, x = idx - y * S;
        int y0, y1, x0, x1;
        float ly, lx;
        bilin_tap(y, P, S, y0, y1, ly);
        bilin_tap(x, P, S, x0, x1, lx);
        const float* src = preds + (size_t)b * P * P;
        const float hx = 1.f - lx, hy = 1.f - ly;
        const float v = hy * (hx * src[y0 * P + x0] + lx * src[y0 * P + x1]) +
                        ly * (hx * src[y1 * P + x0] + lx * src[y1 * P + x1]);
        p = 1.f / (1.f + expf(-v));
        l1[(size_t)b * S * S + idx] = v;
        prob[(size_t)b * S * S + idx] = p;
    }
    p = warp_sum(p);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sums[b], p);
}
// 19x19 zero-padded box mean (divide by 361 always), local entropy, global max
__global__ void ge_box_entropy_kernel(const float* __restrict__ prob, int S, float* __restrict__ en,
                                      unsigned int* __restrict__ enmax, int max_per_image) {
    const int b = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float e = 0.f;
    if (idx < S * S) {
        const int y = idx / S, x = idx - y * S;
        const float* src = prob + (size_t)b * S * S;
        float acc = 0.f;
        const int ya = y - 9 < 0 ? 0 : y - 9, yb = y + 9 >= S ? S - 1 : y + 9;
        const int xa = x - 9 < 0 ? 0 : x - 9, xb = x + 9 >= S ? S - 1 : x + 9;
        for (int yy = ya; yy <= yb; ++yy) {
            float row = 0.f;
            for (int xx = xa; xx <= xb; ++xx) row += src[yy * S + xx];
            acc += row;
        }
        const float f = acc / 361.0f;
        e = -f * logf(fmaxf(f, 1e-5f));
        en[(size_t)b * S * S + idx] = e;
    }
    e = warp_max(e);
    // e >= 0: uint order == float order
    if ((threadIdx.x & 31) == 0 && e > 0.f) atomicMax(enmax + (max_per_image ? b : 0), __float_as_uint(e));
}
__global__ void ge_fuse_kernel(const float* __restrict__ l1, const float* __restrict__ en,
                               const float* __restrict__ l2, const float* __restrict__ sums,
                               const unsigned int* __restrict__ enmax, int max_per_image, int S,
                               const float* __restrict__ w0, const float* __restrict__ b0,
                               const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ out,
                               float* __restrict__ weight) {
    __shared__ float sw0[64], sb0[64], sw2[64];
    if (threadIdx.x < 64) sw0[threadIdx.x] = w0[threadIdx.x], sb0[threadIdx.x] = b0[threadIdx.x], sw2[threadIdx.x] = w2[threadIdx.x];
    __syncthreads();
    const int b = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * S) return;
    const size_t o = (size_t)b * S * S + idx;
    const float emax = __uint_as_float(enmax[max_per_image ? b : 0]);
    const float fg_g = sums[b] / (float)(S * S);
    const float wl = ((1.f - en[o] / emax) + fg_g) * 0.5f;
    const float y = l1[o] * wl + l2[o] * (1.f - wl);
    float acc = b2[0];
#pragma unroll 8
    for (int c = 0; c < 64; ++c) acc += sw2[c] * fmaxf(sw0[c] * y + sb0[c], 0.f);
    out[o] = acc;
    weight[o] = wl;
}

size_t coral_gated_ensemble_workspace_bytes(int B, int S) {
    return ((size_t)3 * B * S * S + 2 * (size_t)B + 4) * sizeof(float);
}
int coral_gated_ensemble(const float* preds, int P, const float* h_preds, int B, int S, int max_per_image,
                         const float* w0, const float* b0, const float* w2, const float* b2, float* out, float* weight,
                         void* workspace, size_t ws_bytes, cudaStream_t stream) {
    UCOD_REQUIRE(preds && h_preds && w0 && b0 && w2 && b2 && out && weight && workspace,
                 "coral_gated_ensemble: null pointer");
    UCOD_REQUIRE(ws_bytes >= coral_gated_ensemble_workspace_bytes(B, S), "coral_gated_ensemble: workspace too small");
    float* l1 = static_cast<float*>(workspace);
    float* prob = l1 + (size_t)B * S * S;
    float* en = prob + (size_t)B * S * S;
    float* sums = en + (size_t)B * S * S;
    unsigned int* enmax = reinterpret_cast<unsigned int*>(sums + B);
    UCOD_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)B * sizeof(float), stream));
    dim3 grid(ceil_div(S * S, 256), B);
    ProfScope ps(KC_OTHER, stream, (double)B * S * S * 4 * 9);
    ge_upsample_kernel<<<grid, 256, 0, stream>>>(preds, P, S, l1, prob, sums);
    ge_box_entropy_kernel<<<grid, 256, 0, stream>>>(prob, S, en, enmax, max_per_image);
    ge_fuse_kernel<<<grid, 256, 0, stream>>>(l1, en, h_preds, sums, enmax, max_per_image, S, w0, b0, w2, b2, out,
                                             weight);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row, fp32 two-pass statistics, bf16 output (dim % 128 == 0, <= 1024)
template <int V>
__global__ void layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ bsh, __nv_bfloat16* __restrict__ y, int rows, int D,
                                      float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i] = xr[lane + 32 * i];
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mu = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
        q += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
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

int layernorm_rows_bf16(const float* x, const float* w, const float* b, void* y, int rows, int dim, float eps,
                        cudaStream_t stream) {
    UCOD_REQUIRE(x && w && b && y && rows > 0, "layernorm: bad argument");
    UCOD_REQUIRE(dim % 128 == 0 && dim >= 128 && dim <= 1024, "layernorm: dim %d must be a multiple of 128 (<= 1024)", dim);
    const int wpb = 8;
    ProfScope ps(KC_LAYERNORM, stream, (double)rows * dim * 6);
    auto* yo = static_cast<__nv_bfloat16*>(y);
    const int grid = ceil_div(rows, wpb);
    switch (dim / 128) {
        case 1: layernorm_rows_kernel<1><<<grid, wpb * 32, 0, stream>>>(x, w, b, yo, rows, dim, eps); break;
        case 2: layernorm_rows_kernel<2><<<grid, wpb * 32, 0, stream>>>(x, w, b, yo, rows, dim, eps); break;
        case 3: layernorm_rows_kernel<3><<<grid, wpb * 32, 0, stream>>>(x, w, b, yo, rows, dim, eps); break;
        case 4: layernorm_rows_kernel<4><<<grid, wpb * 32, 0, stream>>>(x, w, b, yo, rows, dim, eps); break;
        case 5: layernorm_rows_kernel<5><<<grid, wpb * 32, 0, stream>>>(x, w, b, yo, rows, dim, eps); break;
        case 6: layernorm_rows_kernel<6><<<grid, wpb * 32, 0, stream>>>(x, w, b, yo, rows, dim, eps); break;
        case 7: layernorm_rows_kernel<7><<<grid, wpb * 32, 0, stream>>>(x, w, b, yo, rows, dim, eps); break;
        default: layernorm_rows_kernel<8><<<grid, wpb * 32, 0, stream>>>(x, w, b, yo, rows, dim, eps); break;
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void cast_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        out[i] = o;
    }
}
int cast_f32_to_bf16(const float* in, void* out, size_t n, cudaStream_t stream) {
    UCOD_REQUIRE(in && out && n % 4 == 0, "cast_f32_to_bf16: n must be a multiple of 4");
    const size_t n4 = n / 4;
    const unsigned grid = (unsigned)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    ProfScope ps(KC_OTHER, stream, (double)n * 6);
    cast_bf16_kernel<<<grid ? grid : 1, 256, 0, stream>>>(reinterpret_cast<const float4*>(in),
                                                         static_cast<uint2*>(out), n4);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// [B,C,P] (strides sb,sc,sp) fp32 -> token-major fp32 [B,P,C]; 32x32 smem-tiled transpose
__global__ void features_to_tokens_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int P,
                                              long long sb, long long sc, long long sp) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const float* src = in + (size_t)b * sb;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < P) ? src[(size_t)c * sc + (size_t)p * sp] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < P && c < C) out[((size_t)b * P + p) * C + c] = tile[threadIdx.x][i];
    }
}
int features_to_tokens_f32(const float* in, float* out, int B, int C, int P, long long sb, long long sc, long long sp,
                           cudaStream_t stream) {
    UCOD_REQUIRE(in && out && B > 0 && C > 0 && P > 0, "features_to_tokens_f32: bad argument");
    dim3 grid(ceil_div(P, 32), ceil_div(C, 32), B), block(32, 8);
    ProfScope ps(KC_OTHER, stream, (double)B * C * P * 8);
    features_to_tokens_f32_kernel<<<grid, block, 0, stream>>>(in, out, C, P, sb, sc, sp);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// token-major bilinear resize (F.interpolate(mode='bilinear', align_corners=False) per channel):
// in [n, gin_h*gin_w, C] fp32 -> out [n, gout_h*gout_w, C] fp32 and/or bf16.  One CTA per output token.
__global__ void resize_tokens_kernel(const float* __restrict__ in, float* __restrict__ out_f32,
                                     __nv_bfloat16* __restrict__ out_bf16, int gih, int giw, int goh, int gow, int C) {
    const int n = blockIdx.y, t = blockIdx.x;
    const int y = t / gow, x = t - y * gow;
    int y0, y1, x0, x1;
    float ly, lx;
    bilin_tap(y, gih, goh, y0, y1, ly);
    bilin_tap(x, giw, gow, x0, x1, lx);
    const float hx = 1.f - lx, hy = 1.f - ly;
    const float* base = in + (size_t)n * gih * giw * C;
    const float4* r00 = reinterpret_cast<const float4*>(base + (size_t)(y0 * giw + x0) * C);
    const float4* r01 = reinterpret_cast<const float4*>(base + (size_t)(y0 * giw + x1) * C);
    const float4* r10 = reinterpret_cast<const float4*>(base + (size_t)(y1 * giw + x0) * C);
    const float4* r11 = reinterpret_cast<const float4*>(base + (size_t)(y1 * giw + x1) * C);
    const size_t o = ((size_t)n * goh * gow + t) * C;
    for (int c = threadIdx.x; c < C / 4; c += blockDim.x) {
        const float4 a = r00[c], b = r01[c], d = r10[c], e = r11[c];
        float4 v;
        v.x = hy * (hx * a.x + lx * b.x) + ly * (hx * d.x + lx * e.x);
        v.y = hy * (hx * a.y + lx * b.y) + ly * (hx * d.y + lx * e.y);
        v.z = hy * (hx * a.z + lx * b.z) + ly * (hx * d.z + lx * e.z);
        v.w = hy * (hx * a.w + lx * b.w) + ly * (hx * d.w + lx * e.w);
        if (out_f32) reinterpret_cast<float4*>(out_f32 + o)[c] = v;
        if (out_bf16) {
            uint2 p;
            p.x = pack_bf16x2(v.x, v.y);
            p.y = pack_bf16x2(v.z, v.w);
            reinterpret_cast<uint2*>(out_bf16 + o)[c] = p;
        }
    }
}
int resize_tokens_bilinear(const float* in, float* out_f32, void* out_bf16, int n, int gih, int giw, int goh, int gow,
                           int C, cudaStream_t stream) {
    UCOD_REQUIRE(in && (out_f32 || out_bf16) && n > 0 && C % 4 == 0, "resize_tokens_bilinear: bad argument");
    dim3 grid(goh * gow, n);
    ProfScope ps(KC_RESAMPLE, stream, (double)n * goh * gow * C * 8);
    resize_tokens_kernel<<<grid, 192, 0, stream>>>(in, out_f32, static_cast<__nv_bfloat16*>(out_bf16), gih, giw, goh,
                                                   gow, C);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
