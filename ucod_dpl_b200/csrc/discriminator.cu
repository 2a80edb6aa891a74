// APM scorer (`Discriminator`, models/discriminator.py:73-95, dis_use_features=False) and the Adaptive
// Pseudo-label Module fusion (`TrainLoop.merge_pseudo_label`, engine/runner/loop_UCOD_DPL.py:257-272).
//
// Discriminator: mask [B,1,fs,fs] -> conv3x3(1->32)+BN+LReLU(0.1) -> conv3x3 s2 (32->16)+BN+LReLU ->
// conv3x3 s2 (16->8)+BN+LReLU -> Linear(8*((fs+3)/4)^2 -> 1) -> sigmoid.  The reference never puts the
// discriminator in eval mode, so BatchNorm normally runs with BATCH statistics (and updates its running
// buffers): each conv kernel writes raw outputs + per-channel sum / sum-of-squares, and the NEXT kernel
// normalises on load.  8 473 parameters: everything but the activations lives in shared memory / L2.
#include "discriminator.cuh"

#include "prof.cuh"

namespace ucod {

namespace {

constexpr float LRELU = 0.1f;

struct BnParams {
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
};

// The BatchNorm sums of a forward call live in DISC_SLOTS copies of a [32+16+8][2] fp64 table: a producer CTA adds into
// slot (blockIdx.x % DISC_SLOTS) so that thousands of small CTAs do not serialise on 112 addresses; readers add the slots.
constexpr int DISC_SLOTS = 8;
constexpr int DISC_SUMS = (32 + 16 + 8) * 2;                       // doubles per slot
constexpr size_t DISC_SUMS_BYTES = (size_t)DISC_SLOTS * DISC_SUMS * 8;
__device__ __forceinline__ double slot_sum(const double* sums, int i) {
    double t = 0.0;
#pragma unroll
    for (int s = 0; s < DISC_SLOTS; ++s) t += sums[s * DISC_SUMS + i];
    return t;
}

// scale/shift for channel c from either batch sums (train) or running buffers (eval)
__device__ __forceinline__ void bn_affine(int c, const double* sums, double count, const BnParams& bn, int train,
                                          float eps, float& scale, float& shift) {
    float mean, var;
    if (train) {
        const double m = slot_sum(sums, 2 * c) / count;
        double v = slot_sum(sums, 2 * c + 1) / count - m * m;
        if (v < 0) v = 0;
        mean = (float)m, var = (float)v;
    } else {
        mean = bn.running_mean[c], var = bn.running_var[c];
    }
    const float inv = rsqrtf(var + eps);
    scale = bn.gamma[c] * inv;
    shift = bn.beta[c] - mean * scale;
}

__device__ __forceinline__ void bn_update_running(int c, const double* sums, double count, const BnParams& bn,
                                                  float momentum) {
    const double m = slot_sum(sums, 2 * c) / count;
    double v = slot_sum(sums, 2 * c + 1) / count - m * m;
    if (v < 0) v = 0;
    const double unbiased = count > 1 ? v * count / (count - 1) : v;
    bn.running_mean[c] = (1.f - momentum) * bn.running_mean[c] + momentum * (float)m;
    bn.running_var[c] = (1.f - momentum) * bn.running_var[c] + momentum * (float)unbiased;
}

// Direct 3x3 convolution, pad 1, stride STRIDE, no bias.  One thread = one output pixel, all COUT channels.
// IN_BN: input is the previous layer's RAW conv output; BatchNorm + LeakyReLU are applied on load.
template <int CIN, int COUT, int STRIDE, bool IN_BN>
__global__ void __launch_bounds__(128)
    disc_conv_kernel(const float* __restrict__ in, const float* __restrict__ weight /*[COUT,CIN,3,3]*/,
                     float* __restrict__ out, double* __restrict__ out_sums /*[COUT,2]*/,
                     const double* __restrict__ in_sums, BnParams in_bn, int bn_train, int update_running,
                     float eps, float momentum, int B, int Hin, int Win, int Hout, int Wout, size_t in_gstride,
                     size_t ws_gstride /*bytes between the groups' workspaces*/, int tiles_per_block) {
    // blockIdx.z = group: an independent forward call (own batch statistics, own workspace) sharing the launch
    const int grp = blockIdx.z;
    in += (size_t)grp * in_gstride;
    out = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(out) + (size_t)grp * ws_gstride);
    out_sums = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(out_sums) + (size_t)grp * ws_gstride);
    const double* in_sums0 = in_sums;
    if (IN_BN) in_sums = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(in_sums) + (size_t)grp * ws_gstride);
    __shared__ __align__(16) float s_w[CIN * 9 * COUT];  // [c][k][o]: the COUT weights of one tap are contiguous (LDS.128)
    __shared__ float s_scale[CIN], s_shift[CIN];
    for (int j = threadIdx.x; j < COUT * CIN * 9; j += blockDim.x) s_w[j] = weight[(j % COUT) * (CIN * 9) + j / COUT];
    if (IN_BN) {
        const double count = (double)B * Hin * Win;
        for (int c = threadIdx.x; c < CIN; c += blockDim.x) {
            bn_affine(c, in_sums, count, in_bn, bn_train, eps, s_scale[c], s_shift[c]);
            // running buffers: the groups are the reference's consecutive calls, applied in call order by one thread
            if (bn_train && update_running && blockIdx.x == 0 && blockIdx.y == 0 && grp == 0)
                for (int g = 0; g < (int)gridDim.z; ++g)
                    bn_update_running(c, reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(in_sums0) +
                                                                         (size_t)g * ws_gstride),
                                      count, in_bn, momentum);
        }
    }
    __syncthreads();

    // CTA = 32 output pixels x 4 warps (four times the threads of a pixel-per-thread mapping: the layers are tiny and
    // the launch is latency-bound).  CIN >= 4: the warps split the INPUT channels (every input value is loaded and
    // normalised once per CTA) and their partial sums are added in warp order through shared memory; CIN == 1: the warps
    // split the output channels.  Either way a warp's lanes end up holding the same OG output channels, so the
    // BatchNorm sums are plain warp reductions.
    constexpr int OG = COUT / 4;
    constexpr bool KSPLIT = CIN >= 4;
    static_assert(COUT % 4 == 0 && (!KSPLIT || CIN % 4 == 0), "channel split");
    constexpr int NACC = KSPLIT ? COUT : OG;
    __shared__ float s_part[KSPLIT ? 4 * COUT * 32 : 1];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c_begin = KSPLIT ? warp * (CIN / 4) : 0, c_end = KSPLIT ? c_begin + CIN / 4 : CIN;
    const int o_first = KSPLIT ? 0 : warp * OG;  // first output channel of acc[]
    const float* img = in + (size_t)b * CIN * Hin * Win;
    const int n_tiles = (Hout * Wout + 31) / 32;
    const int tile_end = min(n_tiles, ((int)blockIdx.x + 1) * tiles_per_block);
    float ssum[OG], qsum[OG];  // BatchNorm sums of this warp's channels over the CTA's tiles
#pragma unroll
    for (int j = 0; j < OG; ++j) ssum[j] = 0.f, qsum[j] = 0.f;
    // several 32-pixel tiles per CTA: the weight / BatchNorm prologue above costs more than one tile's arithmetic
    for (int tile = blockIdx.x * tiles_per_block; tile < tile_end; ++tile) {
        const int pix = tile * 32 + lane;
        const bool active = pix < Hout * Wout;
        float acc[NACC];
#pragma unroll
        for (int o = 0; o < NACC; ++o) acc[o] = 0.f;
        if (active) {
            const int oy = pix / Wout, ox = pix - oy * Wout;
            for (int c = c_begin; c < c_end; ++c) {
                float v[9];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int iy = oy * STRIDE + ky - 1, ix = ox * STRIDE + kx - 1;
                        float t = 0.f;  // zero padding applies AFTER BN + LeakyReLU of the previous block
                        if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
                            t = img[((size_t)c * Hin + iy) * Win + ix];
                            if (IN_BN) {
                                t = t * s_scale[c] + s_shift[c];
                                t = t > 0.f ? t : LRELU * t;
                            }
                        }
                        v[ky * 3 + kx] = t;
                    }
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float4* w4 = reinterpret_cast<const float4*>(s_w + (c * 9 + k) * COUT + o_first);
#pragma unroll
                    for (int o4 = 0; o4 < NACC / 4; ++o4) {
                        const float4 w = w4[o4];
                        acc[4 * o4 + 0] += v[k] * w.x, acc[4 * o4 + 1] += v[k] * w.y;
                        acc[4 * o4 + 2] += v[k] * w.z, acc[4 * o4 + 3] += v[k] * w.w;
                    }
                }
            }
        }
        float fin[OG];  // this warp's OG output channels [warp*OG, ...) of pixel `pix`
        if constexpr (KSPLIT) {
            if (tile != (int)blockIdx.x * tiles_per_block) __syncthreads();  // previous tile's partials consumed
#pragma unroll
            for (int o = 0; o < COUT; ++o) s_part[(warp * COUT + o) * 32 + lane] = acc[o];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < OG; ++j) {
                const int o = warp * OG + j;
                fin[j] = (s_part[(0 * COUT + o) * 32 + lane] + s_part[(1 * COUT + o) * 32 + lane]) +
                         (s_part[(2 * COUT + o) * 32 + lane] + s_part[(3 * COUT + o) * 32 + lane]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < OG; ++j) fin[j] = acc[j];
        }
        if (active) {
            float* dst = out + ((size_t)b * COUT + warp * OG) * Hout * Wout + pix;
#pragma unroll
            for (int j = 0; j < OG; ++j) {
                dst[(size_t)j * Hout * Wout] = fin[j];
                ssum[j] += fin[j], qsum[j] += fin[j] * fin[j];
            }
        }
    }
    // per-channel sum / sum of squares of the raw outputs (for the next BatchNorm): a warp reduction (fixed order), then
    // one fp64 atomic per channel and CTA into this CTA's slot, whose order only matters below fp32 resolution
#pragma unroll
    for (int j = 0; j < OG; ++j) {
        const float sv = warp_sum(ssum[j]), q = warp_sum(qsum[j]);
        if (lane == 0) {
            double* slot = out_sums + (blockIdx.x % DISC_SLOTS) * DISC_SUMS;
            atomicAdd(&slot[2 * (warp * OG + j)], (double)sv);
            atomicAdd(&slot[2 * (warp * OG + j) + 1], (double)q);
        }
    }
}

// BN3 + LeakyReLU + flatten + Linear + sigmoid: one CTA per image
template <int C>
__global__ void __launch_bounds__(256)
    disc_head_kernel(const float* __restrict__ in /*[B,C,H,W] raw*/, const double* __restrict__ in_sums, BnParams bn,
                     int bn_train, int update_running, float eps, float momentum,
                     const float* __restrict__ lin_w /*[C*H*W]*/, const float* __restrict__ lin_b,
                     float* __restrict__ prob /*[groups*B]*/, int B, int H, int W, size_t ws_gstride) {
    __shared__ float s_scale[C], s_shift[C];
    __shared__ float red[8];
    const int grp = blockIdx.y;
    const double* in_sums0 = in_sums;
    in = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(in) + (size_t)grp * ws_gstride);
    in_sums = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(in_sums) + (size_t)grp * ws_gstride);
    prob += (size_t)grp * B;
    const double count = (double)B * H * W;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        bn_affine(c, in_sums, count, bn, bn_train, eps, s_scale[c], s_shift[c]);
        if (bn_train && update_running && blockIdx.x == 0 && grp == 0)
            for (int g = 0; g < (int)gridDim.y; ++g)
                bn_update_running(c, reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(in_sums0) +
                                                                     (size_t)g * ws_gstride),
                                  count, bn, momentum);
    }
    __syncthreads();
    const int b = blockIdx.x;
    const int n = C * H * W, hw = H * W;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = i / hw;
        float t = in[(size_t)b * n + i] * s_scale[c] + s_shift[c];
        t = t > 0.f ? t : LRELU * t;
        acc += t * lin_w[i];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = lin_b[0];
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        prob[b] = 1.f / (1.f + expf(-t));
    }
}

// APM step 1: hard masks for the discriminator.  student/teacher: sigmoid(x) > 0.5 ; pseudo-label: pl > 0.5
__global__ void apm_binarize_kernel(const float* __restrict__ student, const float* __restrict__ teacher,
                                    const float* __restrict__ pl, float* __restrict__ s_mask,
                                    float* __restrict__ t_mask, float* __restrict__ p_mask, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    s_mask[i] = student[i] > 0x1.8p-24f ? 1.f : 0.f;
    t_mask[i] = teacher[i] > 0x1.8p-24f ? 1.f : 0.f;
    p_mask[i] = pl[i] > 0.5f ? 1.f : 0.f;
}

// APM step 2: w_b = clamp(0.5*(1+cos(pi*|p_s-p_p|)) + epoch_term, 0, 1); merged = pl*(1-w) + T*w;
// dis_loss = mean_b BCE(p_s, 0) = mean_b -max(log(1-p_s), -100)   (torch BCELoss clamps log at -100)
__global__ void apm_merge_kernel(const float* __restrict__ pl, const float* __restrict__ t_mask,
                                 const float* __restrict__ p_s, const float* __restrict__ p_p, float epoch_term,
                                 float* __restrict__ merged, float* __restrict__ weight, float* __restrict__ dis_loss,
                                 int B, int npix) {
    const int b = blockIdx.y;
    const float d = fabsf(p_s[b] - p_p[b]);
    float w = 0.5f * (1.f + cosf(d * 3.14159265358979323846f)) + epoch_term;
    w = fminf(fmaxf(w, 0.f), 1.f);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) {
        const size_t o = (size_t)b * npix + i;
        merged[o] = pl[o] * (1.f - w) + t_mask[o] * w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        weight[b] = w;
        if (b == 0 && dis_loss != nullptr) {
            float acc = 0.f;
            for (int k = 0; k < B; ++k) acc += -fmaxf(logf(1.f - p_s[k]), -100.f);
            dis_loss[0] = acc / (float)B;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward of the discriminator for its own training epochs (`TrainLoop.Discriminator_epoch`,
// engine/runner/loop_UCOD_DPL.py:230-255): BatchNorm in train mode (batch statistics), LeakyReLU(0.1), BCE.
// With z = raw conv output, xh = (z - mu) / sigma, y = gamma xh + beta, h = lrelu(y):
//   dy = dh * lrelu'(y);  dgamma = sum dy xh;  dbeta = sum dy;
//   dz = gamma / sigma * (dy - mean(dy) - xh * mean(dy xh))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bn_stats(int c, const double* sums, double count, float eps, float& mean, float& inv) {
    const double m = slot_sum(sums, 2 * c) / count;
    double v = slot_sum(sums, 2 * c + 1) / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    inv = rsqrtf((float)v + eps);
}

// head: dt[b] = dprob-side gradient wrt the pre-sigmoid logit; dh3 = dt * lin_w; dlin_w += dt * h3; dlin_b += dt
template <int C>
__global__ void __launch_bounds__(256)
    disc_head_bwd_kernel(const float* __restrict__ z3, const double* __restrict__ sums3, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, const float* __restrict__ lin_w,
                         const float* __restrict__ dt, float* __restrict__ dh3, float* __restrict__ g_lin_w,
                         float* __restrict__ g_lin_b, int B, int hw) {
    __shared__ float s_scale[C], s_shift[C];
    const double count = (double)B * hw;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, inv;
        bn_stats(c, sums3, count, eps, mean, inv);
        s_scale[c] = gamma[c] * inv;
        s_shift[c] = beta[c] - mean * gamma[c] * inv;
    }
    __syncthreads();
    const int b = blockIdx.x, n = C * hw;
    const float d = dt[b];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = i / hw;
        float t = z3[(size_t)b * n + i] * s_scale[c] + s_shift[c];
        t = t > 0.f ? t : LRELU * t;
        dh3[(size_t)b * n + i] = d * lin_w[i];
        atomicAdd(g_lin_w + i, d * t);
    }
    if (threadIdx.x == 0) atomicAdd(g_lin_b, d);
}

// pass 1 over a layer: dy = dh * lrelu'(y) (written in place over dh), per-channel sums of dy and dy*xh
template <int C>
__global__ void __launch_bounds__(256)
    disc_bn_bwd_reduce_kernel(const float* __restrict__ z, float* __restrict__ dh, const double* __restrict__ sums,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                              double* __restrict__ dsums /*[C][2]*/, int B, int hw) {
    const int c = blockIdx.y, b = blockIdx.z;
    float mean, inv;
    bn_stats(c, sums, (double)B * hw, eps, mean, inv);
    const float g = gamma[c], be = beta[c];
    const size_t base = ((size_t)b * C + c) * hw;
    float s0 = 0.f, s1 = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const float xh = (z[base + i] - mean) * inv;
        const float y = g * xh + be;
        const float dy = dh[base + i] * (y > 0.f ? 1.f : LRELU);
        dh[base + i] = dy;
        s0 += dy, s1 += dy * xh;
    }
    s0 = warp_sum(s0), s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&dsums[2 * c], (double)s0);
        atomicAdd(&dsums[2 * c + 1], (double)s1);
    }
}
// pass 2: dz = gamma*inv * (dy - mean(dy) - xh*mean(dy*xh)) in place; dgamma / dbeta written by block (0,c,0)
template <int C>
__global__ void __launch_bounds__(256)
    disc_bn_bwd_apply_kernel(const float* __restrict__ z, float* __restrict__ dy, const double* __restrict__ sums,
                             const float* __restrict__ gamma, float eps, const double* __restrict__ dsums,
                             float* __restrict__ g_gamma, float* __restrict__ g_beta, int B, int hw) {
    const int c = blockIdx.y, b = blockIdx.z;
    const double count = (double)B * hw;
    float mean, inv;
    bn_stats(c, sums, count, eps, mean, inv);
    const float m0 = (float)(dsums[2 * c] / count), m1 = (float)(dsums[2 * c + 1] / count);
    const float k = gamma[c] * inv;
    const size_t base = ((size_t)b * C + c) * hw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const float xh = (z[base + i] - mean) * inv;
        dy[base + i] = k * (dy[base + i] - m0 - xh * m1);
    }
    if (blockIdx.x == 0 && b == 0 && threadIdx.x == 0) {
        g_gamma[c] += (float)dsums[2 * c + 1];
        g_beta[c] += (float)dsums[2 * c];
    }
}
// weight gradient: one thread per weight element and image; dW[co,ci,k] += sum_pix dz[b,co,pix] * hin[b,ci,pix*s+k-1]
template <int CIN, int COUT, int STRIDE, bool IN_BN>
__global__ void __launch_bounds__(128)
    disc_conv_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dz, const double* __restrict__ in_sums,
                           const float* __restrict__ in_gamma, const float* __restrict__ in_beta, float eps,
                           float* __restrict__ g_w, int B, int Hin, int Win, int Hout, int Wout) {
    const int widx = blockIdx.x * blockDim.x + threadIdx.x;
    if (widx >= COUT * CIN * 9) return;
    const int b = blockIdx.y;
    const int co = widx / (CIN * 9), ci = (widx / 9) % CIN, k = widx % 9, ky = k / 3, kx = k % 3;
    float scale = 1.f, shift = 0.f;
    if (IN_BN) {
        float mean, inv;
        bn_stats(ci, in_sums, (double)B * Hin * Win, eps, mean, inv);
        scale = in_gamma[ci] * inv;
        shift = in_beta[ci] - mean * scale;
    }
    const float* img = in + ((size_t)b * CIN + ci) * Hin * Win;
    const float* d = dz + ((size_t)b * COUT + co) * Hout * Wout;
    float acc = 0.f;
    for (int oy = 0; oy < Hout; ++oy) {
        const int iy = oy * STRIDE + ky - 1;
        if (iy < 0 || iy >= Hin) continue;
        for (int ox = 0; ox < Wout; ++ox) {
            const int ix = ox * STRIDE + kx - 1;
            if (ix < 0 || ix >= Win) continue;
            float t = img[iy * Win + ix];
            if (IN_BN) {
                t = t * scale + shift;
                t = t > 0.f ? t : LRELU * t;
            }
            acc += d[oy * Wout + ox] * t;
        }
    }
    atomicAdd(g_w + widx, acc);
}
// data gradient: dh_in[b,ci,iy,ix] = sum_{co,ky,kx} dz[b,co,oy,ox] * W[co,ci,ky,kx], iy = oy*s + ky - 1
template <int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(128)
    disc_conv_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ weight, float* __restrict__ dh_in,
                           int Hin, int Win, int Hout, int Wout) {
    __shared__ float s_w[COUT * CIN * 9];
    for (int i = threadIdx.x; i < COUT * CIN * 9; i += blockDim.x) s_w[i] = weight[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= Hin * Win) return;
    const int iy = pix / Win, ix = pix - iy * Win;
    float acc[CIN];
#pragma unroll
    for (int c = 0; c < CIN; ++c) acc[c] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
        const int ty = iy + 1 - ky;
        if (ty < 0 || ty % STRIDE) continue;
        const int oy = ty / STRIDE;
        if (oy >= Hout) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int tx = ix + 1 - kx;
            if (tx < 0 || tx % STRIDE) continue;
            const int ox = tx / STRIDE;
            if (ox >= Wout) continue;
            for (int co = 0; co < COUT; ++co) {
                const float d = dz[((size_t)b * COUT + co) * Hout * Wout + oy * Wout + ox];
                const float* w = s_w + (size_t)co * CIN * 9 + ky * 3 + kx;
#pragma unroll
                for (int c = 0; c < CIN; ++c) acc[c] += d * w[c * 9];
            }
        }
    }
    float* dst = dh_in + (size_t)b * CIN * Hin * Win + pix;
#pragma unroll
    for (int c = 0; c < CIN; ++c) dst[(size_t)c * Hin * Win] = acc[c];
}
// dt[b] = (prob[b] - label) / n_total  (BCE mean over n_total samples, through the sigmoid); loss += BCE terms
__global__ void disc_bce_grad_kernel(const float* __restrict__ prob, float label, float inv_n, float* __restrict__ dt,
                                     float* __restrict__ loss, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float p = prob[b];
    dt[b] = (p - label) * inv_n;
    const float l = -(label * fmaxf(logf(p), -100.f) + (1.f - label) * fmaxf(logf(1.f - p), -100.f));
    atomicAdd(loss, l * inv_n);
}

}  // namespace

size_t discriminator_workspace_bytes(int B, int fs) {
    const int h2 = (fs + 1) / 2, h3 = (h2 + 1) / 2;
    return ((size_t)B * 32 * fs * fs + (size_t)B * 16 * h2 * h2 + (size_t)B * 8 * h3 * h3) * 4 + DISC_SUMS_BYTES +
           1024;
}
static size_t disc_group_stride(int B, int fs) { return (discriminator_workspace_bytes(B, fs) + 255) / 256 * 256; }
size_t discriminator_workspace_bytes_groups(int B, int fs, int groups) { return disc_group_stride(B, fs) * (size_t)groups; }

int discriminator_forward(const float* mask, int B, int fs, const DiscWeights& w, int bn_train, int update_running,
                          float* prob, void* workspace, size_t ws_bytes, cudaStream_t stream, int groups) {
    UCOD_REQUIRE(mask && prob && workspace, "discriminator_forward: null argument");
    UCOD_REQUIRE(B > 0 && fs > 0 && groups >= 1 && groups <= 8, "discriminator_forward: bad geometry");
    const size_t gstride = groups > 1 ? disc_group_stride(B, fs) : 0;
    UCOD_REQUIRE(ws_bytes >= (groups > 1 ? gstride * groups : discriminator_workspace_bytes(B, fs)),
                 "discriminator_forward: workspace too small");
    UCOD_REQUIRE(groups == 1 || (reinterpret_cast<uintptr_t>(workspace) & 7) == 0,
                 "discriminator_forward: workspace must be 8-byte aligned");
    const int h1 = fs, h2 = (fs + 2 - 3) / 2 + 1, h3 = (h2 + 2 - 3) / 2 + 1;
    uint8_t* p = static_cast<uint8_t*>(workspace);
    double* sums = reinterpret_cast<double*>(p);  // [DISC_SLOTS][32+16+8][2]
    p += DISC_SUMS_BYTES;
    float* a1 = reinterpret_cast<float*>(p);
    p += (size_t)B * 32 * h1 * h1 * 4;
    float* a2 = reinterpret_cast<float*>(p);
    p += (size_t)B * 16 * h2 * h2 * 4;
    float* a3 = reinterpret_cast<float*>(p);
    double *s1 = sums, *s2 = sums + 64, *s3 = sums + 96;
    for (int g = 0; g < groups; ++g)
        UCOD_CHECK_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(sums) + g * gstride, 0, DISC_SUMS_BYTES, stream));
    const float eps = 1e-5f, mom = 0.1f;
    BnParams bn1{w.bn1_w, w.bn1_b, w.bn1_mean, w.bn1_var}, bn2{w.bn2_w, w.bn2_b, w.bn2_mean, w.bn2_var},
        bn3{w.bn3_w, w.bn3_b, w.bn3_mean, w.bn3_var}, none{nullptr, nullptr, nullptr, nullptr};
    ProfScope ps(KC_OTHER, stream, (double)groups * B * (32 * h1 * h1 + 16 * h2 * h2 + 8 * h3 * h3) * 8);
    const size_t gf = gstride / 4;  // group stride of an activation pointer, in floats
    // tiles of 32 pixels per CTA: enough CTAs for ~4 per SM, not more (each CTA re-stages the weights / BN constants)
    auto tiles_per_block = [&](int npix) {
        const int n_tiles = ceil_div(npix, 32);
        const int want = ceil_div(4 * device_sm_count(), B * groups);  // CTAs per (image, call)
        return ceil_div(n_tiles, want < 1 ? 1 : (want > n_tiles ? n_tiles : want));
    };
    const int t1 = tiles_per_block(h1 * h1), t2 = tiles_per_block(h2 * h2), t3 = tiles_per_block(h3 * h3);
    disc_conv_kernel<1, 32, 1, false><<<dim3(ceil_div(ceil_div(h1 * h1, 32), t1), B, groups), 128, 0, stream>>>(
        mask, w.conv1, a1, s1, nullptr, none, 0, 0, eps, mom, B, fs, fs, h1, h1, (size_t)B * fs * fs, gstride, t1);
    disc_conv_kernel<32, 16, 2, true><<<dim3(ceil_div(ceil_div(h2 * h2, 32), t2), B, groups), 128, 0, stream>>>(
        a1, w.conv2, a2, s2, s1, bn1, bn_train, update_running, eps, mom, B, h1, h1, h2, h2, gf, gstride, t2);
    disc_conv_kernel<16, 8, 2, true><<<dim3(ceil_div(ceil_div(h3 * h3, 32), t3), B, groups), 128, 0, stream>>>(
        a2, w.conv3, a3, s3, s2, bn2, bn_train, update_running, eps, mom, B, h2, h2, h3, h3, gf, gstride, t3);
    disc_head_kernel<8><<<dim3(B, groups), 256, 0, stream>>>(a3, s3, bn3, bn_train, update_running, eps, mom, w.lin_w,
                                                              w.lin_b, prob, B, h3, h3, gstride);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

size_t discriminator_backward_workspace_bytes(int B, int fs) {
    const int h2 = (fs + 1) / 2, h3 = (h2 + 1) / 2;
    return ((size_t)B * 32 * fs * fs + (size_t)B * 16 * h2 * h2 + (size_t)B * 8 * h3 * h3 + B) * 4 + 56 * 2 * 8 + 1024;
}

// grads are ACCUMULATED (the epoch adds the pseudo-label call and the student call); the caller zeroes them.
int discriminator_backward(const float* mask, int B, int fs, const DiscWeights& w, const float* prob, float label,
                           int n_total, const DiscGrads& g, float* loss, void* fwd_workspace, void* workspace,
                           size_t ws_bytes, cudaStream_t stream) {
    UCOD_REQUIRE(mask && prob && loss && fwd_workspace && workspace, "discriminator_backward: null argument");
    UCOD_REQUIRE(ws_bytes >= discriminator_backward_workspace_bytes(B, fs), "discriminator_backward: workspace too small");
    const int h1 = fs, h2 = (fs + 2 - 3) / 2 + 1, h3 = (h2 + 2 - 3) / 2 + 1;
    uint8_t* p = static_cast<uint8_t*>(fwd_workspace);
    const double* sums = reinterpret_cast<const double*>(p);
    p += DISC_SUMS_BYTES;
    const float* z1 = reinterpret_cast<const float*>(p);
    p += (size_t)B * 32 * h1 * h1 * 4;
    const float* z2 = reinterpret_cast<const float*>(p);
    p += (size_t)B * 16 * h2 * h2 * 4;
    const float* z3 = reinterpret_cast<const float*>(p);
    const double *s1 = sums, *s2 = sums + 64, *s3 = sums + 96;
    uint8_t* q = static_cast<uint8_t*>(workspace);
    double* dsums = reinterpret_cast<double*>(q);
    q += 56 * 2 * 8;
    float* d1 = reinterpret_cast<float*>(q);
    q += (size_t)B * 32 * h1 * h1 * 4;
    float* d2 = reinterpret_cast<float*>(q);
    q += (size_t)B * 16 * h2 * h2 * 4;
    float* d3 = reinterpret_cast<float*>(q);
    q += (size_t)B * 8 * h3 * h3 * 4;
    float* dt = reinterpret_cast<float*>(q);
    double *ds1 = dsums, *ds2 = dsums + 64, *ds3 = dsums + 96;
    const float eps = 1e-5f;
    UCOD_CHECK_CUDA(cudaMemsetAsync(dsums, 0, 56 * 2 * 8, stream));
    ProfScope ps(KC_OTHER, stream, (double)B * (32 * h1 * h1 + 16 * h2 * h2 + 8 * h3 * h3) * 24);
    disc_bce_grad_kernel<<<ceil_div(B, 128), 128, 0, stream>>>(prob, label, 1.0f / (float)n_total, dt, loss, B);
    disc_head_bwd_kernel<8><<<B, 256, 0, stream>>>(z3, s3, w.bn3_w, w.bn3_b, eps, w.lin_w, dt, d3, g.lin_w, g.lin_b, B,
                                                   h3 * h3);
    // layer 3
    disc_bn_bwd_reduce_kernel<8><<<dim3(2, 8, B), 256, 0, stream>>>(z3, d3, s3, w.bn3_w, w.bn3_b, eps, ds3, B, h3 * h3);
    disc_bn_bwd_apply_kernel<8><<<dim3(2, 8, B), 256, 0, stream>>>(z3, d3, s3, w.bn3_w, eps, ds3, g.bn3_w, g.bn3_b, B,
                                                                   h3 * h3);
    disc_conv_wgrad_kernel<16, 8, 2, true><<<dim3(ceil_div(8 * 16 * 9, 128), B), 128, 0, stream>>>(
        z2, d3, s2, w.bn2_w, w.bn2_b, eps, g.conv3, B, h2, h2, h3, h3);
    disc_conv_dgrad_kernel<16, 8, 2><<<dim3(ceil_div(h2 * h2, 128), B), 128, 0, stream>>>(d3, w.conv3, d2, h2, h2, h3, h3);
    // layer 2
    disc_bn_bwd_reduce_kernel<16><<<dim3(4, 16, B), 256, 0, stream>>>(z2, d2, s2, w.bn2_w, w.bn2_b, eps, ds2, B, h2 * h2);
    disc_bn_bwd_apply_kernel<16><<<dim3(4, 16, B), 256, 0, stream>>>(z2, d2, s2, w.bn2_w, eps, ds2, g.bn2_w, g.bn2_b, B,
                                                                     h2 * h2);
    disc_conv_wgrad_kernel<32, 16, 2, true><<<dim3(ceil_div(16 * 32 * 9, 128), B), 128, 0, stream>>>(
        z1, d2, s1, w.bn1_w, w.bn1_b, eps, g.conv2, B, h1, h1, h2, h2);
    disc_conv_dgrad_kernel<32, 16, 2><<<dim3(ceil_div(h1 * h1, 128), B), 128, 0, stream>>>(d2, w.conv2, d1, h1, h1, h2, h2);
    // layer 1
    disc_bn_bwd_reduce_kernel<32><<<dim3(8, 32, B), 256, 0, stream>>>(z1, d1, s1, w.bn1_w, w.bn1_b, eps, ds1, B, h1 * h1);
    disc_bn_bwd_apply_kernel<32><<<dim3(8, 32, B), 256, 0, stream>>>(z1, d1, s1, w.bn1_w, eps, ds1, g.bn1_w, g.bn1_b, B,
                                                                     h1 * h1);
    disc_conv_wgrad_kernel<1, 32, 1, false><<<dim3(ceil_div(32 * 9, 128), B), 128, 0, stream>>>(
        mask, d1, nullptr, nullptr, nullptr, eps, g.conv1, B, h1, h1, h1, h1);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int apm_binarize(const float* student, const float* teacher, const float* pl, float* s_mask, float* t_mask,
                 float* p_mask, size_t n, cudaStream_t stream) {
    UCOD_REQUIRE(student && teacher && pl && s_mask && t_mask && p_mask, "apm_binarize: null argument");
    ProfScope ps(KC_OTHER, stream, (double)n * 24);
    apm_binarize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(student, teacher, pl, s_mask, t_mask, p_mask,
                                                                        n);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int apm_merge(const float* pl, const float* t_mask, const float* p_s, const float* p_p, float epoch_term,
              float* merged, float* weight, float* dis_loss, int B, int npix, cudaStream_t stream) {
    UCOD_REQUIRE(pl && t_mask && p_s && p_p && merged && weight, "apm_merge: null argument");
    ProfScope ps(KC_OTHER, stream, (double)B * npix * 12);
    apm_merge_kernel<<<dim3(ceil_div(npix, 256), B), 256, 0, stream>>>(pl, t_mask, p_s, p_p, epoch_term, merged, weight,
                                                                      dis_loss, B, npix);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
