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

// weight of output index `o` on input index `q` of the 1-D bilinear map
__device__ __forceinline__ float tap_weight(int o, int q, int in_size, int out_size) {
    int i0, i1;
    float l1;
    bilinear_tap(o, in_size, out_size, i0, i1, l1);
    return (i0 == q ? 1.f - l1 : 0.f) + (i1 == q ? l1 : 0.f);
}
// The outputs that have a tap on input `q` form a contiguous range [lo, lo + cnt): src(o) in (q-1, q+1).  The window
// is bracketed analytically (with slack for rounding) and then tested exactly with bilinear_tap.
__device__ __forceinline__ void tap_range(int q, int in_size, int out_size, int& lo, int& cnt) {
    const float inv = (float)out_size / (float)in_size;
    int a = (int)floorf(((float)q - 0.5f) * inv - 0.5f) - 2;
    int b = (int)ceilf(((float)q + 1.5f) * inv - 0.5f) + 2;
    a = a < 0 ? 0 : a;
    b = b > out_size - 1 ? out_size - 1 : b;
    lo = out_size, cnt = 0;
    for (int o = a; o <= b; ++o) {
        int i0, i1;
        float l1;
        bilinear_tap(o, in_size, out_size, i0, i1, l1);
        if (i0 == q || i1 == q) {
            lo = o < lo ? o : lo;
            ++cnt;
        }
    }
}
// Row q of M = U^T U for the 1-D bilinear map U (out x in): tridiagonal, m[d+1] = sum_o U[o,q] U[o,q+d].
__device__ __forceinline__ void gram_row_1d(int q, int in_size, int out_size, float* m) {
    int lo, cnt;
    tap_range(q, in_size, out_size, lo, cnt);
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    for (int o = lo; o < lo + cnt; ++o) {
        const float u = tap_weight(o, q, in_size, out_size);
        m0 += u * tap_weight(o, q - 1, in_size, out_size);
        m1 += u * u;
        m2 += u * tap_weight(o, q + 1, in_size, out_size);
    }
    m[0] = m0, m[1] = m1, m[2] = m2;
}

constexpr int DEC_MAXW = 256;  // largest supported grid side (input or output)

// sumsq[b, c] = sum over the OUTPUT pixels of d_up[pix, c]^2 with d_up = U d_in, evaluated on the input grid:
//   sum_p (U d)_p^2 = d^T (U^T U) d,  U^T U = (Uy^T Uy) x (Ux^T Ux), each factor tridiagonal
// i.e. a 3x3 stencil per input pixel instead of a 4-tap gather per output pixel (3.4x fewer pixels at 37 -> 68), and
// the per-row partial sums are written, not accumulated: a fixed summation order, no memset.
// One CTA per (input row, image); the head kernel adds the rows in order.
__global__ void __launch_bounds__(256)
    decoder_sumsq_rows_kernel(const float* __restrict__ d_in, float* __restrict__ part, int gin_h, int gin_w, int out_h,
                              int out_w, const int* __restrict__ batch_dev) {
    __shared__ float s_mx[DEC_MAXW][3];
    __shared__ float s_my[3];
    __shared__ __align__(16) float s_part[8][128];
    const int b = blockIdx.y, qy = blockIdx.x;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    for (int q = threadIdx.x; q < gin_w; q += 256) gram_row_1d(q, gin_w, out_w, s_mx[q]);
    if (threadIdx.x == 255) gram_row_1d(qy, gin_h, out_h, s_my);
    __syncthreads();
    // warp = a contiguous strip of the row, lane = 4 channels; sliding window over the column sums
    //   v(x) = sum_dy My[dy] d[qy+dy, x]  (3 row loads per pixel instead of 9),  E(x) = sum_dx Mx[x][dx] v(x+dx)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* d_img = d_in + (size_t)b * gin_h * gin_w * 128;
    const float my0 = qy > 0 ? s_my[0] : 0.f, my1 = s_my[1], my2 = qy < gin_h - 1 ? s_my[2] : 0.f;
    const float4* r0 = reinterpret_cast<const float4*>(d_img + (size_t)(qy > 0 ? qy - 1 : qy) * gin_w * 128) + lane;
    const float4* r1 = reinterpret_cast<const float4*>(d_img + (size_t)qy * gin_w * 128) + lane;
    const float4* r2 = reinterpret_cast<const float4*>(d_img + (size_t)(qy < gin_h - 1 ? qy + 1 : qy) * gin_w * 128) + lane;
    const int per = (gin_w + 7) / 8;
    const int x_begin = warp * per, x_end = x_begin + per < gin_w ? x_begin + per : gin_w;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), centre = acc, centre_next = acc;
    auto column = [&](int x, float4& mid) -> float4 {
        if (x < 0 || x >= gin_w) {
            mid = make_float4(0.f, 0.f, 0.f, 0.f);
            return mid;
        }
        const float4 a = r0[(size_t)x * 32], c2 = r2[(size_t)x * 32];
        mid = r1[(size_t)x * 32];
        return make_float4(my0 * a.x + my1 * mid.x + my2 * c2.x, my0 * a.y + my1 * mid.y + my2 * c2.y,
                           my0 * a.z + my1 * mid.z + my2 * c2.z, my0 * a.w + my1 * mid.w + my2 * c2.w);
    };
    if (x_begin < x_end) {
        float4 dummy;
        float4 vm = column(x_begin - 1, dummy), v0 = column(x_begin, centre);
        for (int qx = x_begin; qx < x_end; ++qx) {
            const float4 vp = column(qx + 1, centre_next);
            const float m0 = s_mx[qx][0], m1 = s_mx[qx][1], m2 = s_mx[qx][2];
            acc.x += centre.x * (m0 * vm.x + m1 * v0.x + m2 * vp.x);
            acc.y += centre.y * (m0 * vm.y + m1 * v0.y + m2 * vp.y);
            acc.z += centre.z * (m0 * vm.z + m1 * v0.z + m2 * vp.z);
            acc.w += centre.w * (m0 * vm.w + m1 * v0.w + m2 * vp.w);
            vm = v0, v0 = vp, centre = centre_next;
        }
    }
    reinterpret_cast<float4*>(s_part[warp])[lane] = acc;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int c = threadIdx.x;
        float t = s_part[0][c];
        for (int w = 1; w < 8; ++w) t += s_part[w][c];
        part[((size_t)b * gin_h + qy) * 128 + c] = t;
    }
}

// gate + heads (+ the normalised features for the orthogonality loss / backward).  CTA = one output row of one image:
// the two input rows it interpolates between are blended in y ONCE per input column into shared memory (gin_w x 128
// floats), so an output pixel costs two shared-memory reads instead of four 512-byte gathers from L2.
__global__ void __launch_bounds__(256)
    decoder_head_kernel(const float* __restrict__ d_in, const float* __restrict__ sumsq, const float* __restrict__ emb,
                        const float* __restrict__ w_fg, const float* __restrict__ b_fg, const float* __restrict__ w_bg,
                        const float* __restrict__ b_bg, float* __restrict__ fg, float* __restrict__ bg,
                        float* __restrict__ fhat_out, int gin_h, int gin_w, int out_h, int out_w,
                        const int* __restrict__ batch_dev, const float* __restrict__ sq_part,
                        float* __restrict__ sumsq_out) {
    extern __shared__ __align__(16) float rowbuf[];  // [gin_w][128]
    __shared__ int s_x0[DEC_MAXW], s_x1[DEC_MAXW];
    __shared__ float s_lx[DEC_MAXW];
    __shared__ __align__(16) float s_ss[128];
    const int b = blockIdx.y, oy = blockIdx.x;
    if (batch_dev != nullptr && b >= __ldg(batch_dev)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* d_img = d_in + (size_t)b * gin_h * gin_w * 128;
    const int npix = out_h * out_w;
    int y0, y1;
    float ly;
    bilinear_tap(oy, gin_h, out_h, y0, y1, ly);
    {
        const float hy = 1.f - ly;
        const float4* ra = reinterpret_cast<const float4*>(d_img + (size_t)y0 * gin_w * 128);
        const float4* rb = reinterpret_cast<const float4*>(d_img + (size_t)y1 * gin_w * 128);
        for (int i = threadIdx.x; i < gin_w * 32; i += 256) {
            const float4 u = ra[i], v = rb[i];
            reinterpret_cast<float4*>(rowbuf)[i] =
                    make_float4(hy * u.x + ly * v.x, hy * u.y + ly * v.y, hy * u.z + ly * v.z, hy * u.w + ly * v.w);
        }
    }
    for (int ox = threadIdx.x; ox < out_w; ox += 256) bilinear_tap(ox, gin_w, out_w, s_x0[ox], s_x1[ox], s_lx[ox]);
    if (sq_part != nullptr) {
        // sum of squares = the input-row partials added in row order.  Every CTA of the image forms the same sum (a
        // separate B-CTA reduction launch cost more than these 37 L2 reads per thread); row 0's CTA publishes it for
        // the backward
        if (threadIdx.x < 128) {
            float t = 0.f;
            const float* sp = sq_part + (size_t)b * gin_h * 128 + threadIdx.x;
            int qy = 0;
            for (; qy + 8 <= gin_h; qy += 8) {  // 8 loads in flight, added in row order
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = sp[(size_t)(qy + u) * 128];
#pragma unroll
                for (int u = 0; u < 8; ++u) t += v[u];
            }
            for (; qy < gin_h; ++qy) t += sp[(size_t)qy * 128];
            s_ss[threadIdx.x] = t;
            if (oy == 0) sumsq_out[(size_t)b * 128 + threadIdx.x] = t;
        }
    } else if (threadIdx.x < 128) {
        s_ss[threadIdx.x] = sumsq[(size_t)b * 128 + threadIdx.x];
    }
    __syncthreads();
    // per-channel constants for this lane's 4 channels
    const float4 ss = reinterpret_cast<const float4*>(s_ss)[lane];
    const float4 e = __ldg(reinterpret_cast<const float4*>(emb) + lane);  // emb[2,64] flat == channel order
    const float4 wh = lane < 16 ? __ldg(reinterpret_cast<const float4*>(w_fg) + lane)
                                : __ldg(reinterpret_cast<const float4*>(w_bg) + (lane - 16));
    float4 g;  // e / max(|e| * sqrt(sumsq), 1e-12)
    g.x = e.x / fmaxf(fabsf(e.x) * sqrtf(ss.x), 1e-12f);
    g.y = e.y / fmaxf(fabsf(e.y) * sqrtf(ss.y), 1e-12f);
    g.z = e.z / fmaxf(fabsf(e.z) * sqrtf(ss.z), 1e-12f);
    g.w = e.w / fmaxf(fabsf(e.w) * sqrtf(ss.w), 1e-12f);
    const float bias_fg = __ldg(b_fg), bias_bg = __ldg(b_bg);
    __syncthreads();

    for (int ox = warp; ox < out_w; ox += 8) {
        const int pix = oy * out_w + ox;
        const float lx = s_lx[ox], hx = 1.f - lx;
        const float4 u = reinterpret_cast<const float4*>(rowbuf + (size_t)s_x0[ox] * 128)[lane];
        const float4 v = reinterpret_cast<const float4*>(rowbuf + (size_t)s_x1[ox] * 128)[lane];
        const float4 d = make_float4(hx * u.x + lx * v.x, hx * u.y + lx * v.y, hx * u.z + lx * v.z, hx * u.w + lx * v.w);
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
// Stage 1: one CTA per (pixel chunk, image) computes both 64x64 Grams of its chunk (below) and writes them to its own
// slot of `part` — no atomics.  Stage 2 sums the slots in a fixed order (bit-reproducible), stage 3 reduces
// <G1, G2>_F - diag in fp64.  (Round 1: 8-pixel tiles, 34 shared loads per 32 FMAs and 8 192 float atomics per CTA:
// 240 us for 16 images.)
constexpr int GRAM_CHUNK = 256;  // pixels per CTA (two staged passes): 19 partial Grams per 68x68 image (128: no faster, twice the partials)
constexpr int GRAM_TILE = 128;   // pixels staged per pass
constexpr int GRAM_LD = 136;     // shared-memory row pitch in words: fragment loads (4 rows x 8 columns) hit 32 banks

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// D (16x8, fp32) += A (16x8, row) * B (8x8, col), TF32 inputs.  g = lane >> 2, t = lane & 3:
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4) | b0 (k=t, n=g)  b1 (k=t+4, n=g) | c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Both 64x64 Grams of a pixel chunk on the tensor cores (mma.sync TF32, fp32 accumulate; round 2a: 4x4 register tiles
// of FFMAs, 47 us for 16 images).  G_k = F_k^T F_k is a [64 x npix] x [npix x 64] product whose two operands are the
// SAME shared-memory tile F[pixel][channel]: the A fragment reads it transposed (m = channel, k = pixel), the B fragment
// directly (k = pixel, n = channel); with a pitch of 136 words both are conflict-free.  The features are rounded to
// TF32 once when staged (round-to-nearest: no bias; the Gram entries are sums over thousands of pixels).  The diagonal
// term sum_p (f1_p . f2_p)^2 is taken from the unrounded values while staging.  Warp w: branch w>>2, channel rows
// 16*(w&3).., all 8 column tiles: 32 accumulator registers.  Partial Grams are written per chunk (fixed-order reduce).
__global__ void __launch_bounds__(256)
    decoder_gram_kernel(const float* __restrict__ fhat, float* __restrict__ part, float* __restrict__ dpart, int npix) {
    extern __shared__ __align__(16) uint32_t sF[];  // [GRAM_TILE][GRAM_LD] TF32 bit patterns
    __shared__ float sdiag[8];
    const int b = blockIdx.y, nchunks = gridDim.x;
    const int p0 = blockIdx.x * GRAM_CHUNK;
    const int p1 = min(npix, p0 + GRAM_CHUNK);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int br = warp >> 2, m0 = br * 64 + (warp & 3) * 16;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float dacc = 0.f;
    for (int p = p0; p < p1; p += GRAM_TILE) {
        const int n = min(GRAM_TILE, p1 - p);
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(fhat + ((size_t)b * npix + p) * 128);
        for (int k = warp; k < GRAM_TILE; k += 8) {  // warp = pixel row, lane = 4 channels (lanes 0-15: branch 1)
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < n) v = src[k * 32 + lane];
            *reinterpret_cast<uint4*>(sF + k * GRAM_LD + 4 * lane) =
                    make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            const float ox = __shfl_xor_sync(0xffffffffu, v.x, 16), oy = __shfl_xor_sync(0xffffffffu, v.y, 16);
            const float oz = __shfl_xor_sync(0xffffffffu, v.z, 16), ow = __shfl_xor_sync(0xffffffffu, v.w, 16);
            float d = lane < 16 ? v.x * ox + v.y * oy + v.z * oz + v.w * ow : 0.f;
            d = warp_sum(d);
            dacc += d * d;  // identical in every lane; zero rows add nothing
        }
        __syncthreads();
#pragma unroll 4
        for (int k0 = 0; k0 < GRAM_TILE; k0 += 8) {
            const uint32_t* r0 = sF + (k0 + t) * GRAM_LD;
            const uint32_t* r1 = sF + (k0 + t + 4) * GRAM_LD;
            const uint32_t a[4] = {r0[m0 + g], r0[m0 + g + 8], r1[m0 + g], r1[m0 + g + 8]};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(acc[nt], a, r0[br * 64 + nt * 8 + g], r1[br * 64 + nt * 8 + g]);
        }
    }
    float* gp = part + ((size_t)b * nchunks + blockIdx.x) * 8192 + br * 4096;
    const int row = (warp & 3) * 16 + g;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<float2*>(gp + row * 64 + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(gp + (row + 8) * 64 + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
    }
    if (lane == 0) sdiag[warp] = dacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tt = 0.f;
        for (int w = 0; w < 8; ++w) tt += sdiag[w];
        dpart[(size_t)b * nchunks + blockIdx.x] = tt;
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
    n = (n + 255) / 256 * 256;
    n += (size_t)B * gin_h * 128 * 4;                // per-row partials of sumsq (last, so the layout above is unchanged)
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
    UCOD_REQUIRE(gin_h <= DEC_MAXW && gin_w <= DEC_MAXW && out_h <= DEC_MAXW && out_w <= DEC_MAXW,
                 "decoder_forward: grids up to %d x %d", DEC_MAXW, DEC_MAXW);
    const double d_bytes = (double)B * gin_h * gin_w * 128 * 4;
    {
        float* sq_part = reinterpret_cast<float*>(base + (need - 4096 - (size_t)B * gin_h * 128 * 4));
        ProfScope ps(KC_DECODER, stream, batch_dev ? 0.0 : d_bytes);
        decoder_sumsq_rows_kernel<<<dim3(gin_h, B), 256, 0, stream>>>(d_in, sq_part, gin_h, gin_w, out_h, out_w, batch_dev);
    }
    const float* sq_part = reinterpret_cast<const float*>(base + (need - 4096 - (size_t)B * gin_h * 128 * 4));
    UCOD_CHECK_CUDA(cudaGetLastError());
    {
        const size_t row_smem = (size_t)gin_w * 128 * sizeof(float);
        static size_t configured = 48 * 1024;
        if (row_smem > configured) {
            UCOD_CHECK_CUDA(cudaFuncSetAttribute(decoder_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)row_smem));
            configured = row_smem;
        }
        ProfScope ps(KC_DECODER, stream, batch_dev ? 0.0 : d_bytes + (double)B * npix * 8);
        decoder_head_kernel<<<dim3(out_h, B), 256, row_smem, stream>>>(d_in, sumsq, w.emb, w.w_fg, w.b_fg, w.w_bg, w.b_bg,
                                                                       fg, bg, fhat, gin_h, gin_w, out_h, out_w,
                                                                       batch_dev, sq_part, sumsq);
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
            constexpr int gram_smem = GRAM_TILE * GRAM_LD * 4;
            static bool configured = false;
            if (!configured) {
                UCOD_CHECK_CUDA(cudaFuncSetAttribute(decoder_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     gram_smem));
                configured = true;
            }
            decoder_gram_kernel<<<g2, 256, gram_smem, stream>>>(fhat, part, dpart, npix);
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
// Data flow (no atomics anywhere, every sum in a fixed order => bit-reproducible gradients):
//   rows kernel   : CTA = one output row of one image.  P1 for the row's pixels goes to shared memory, then the
//                   x-adjoint of the upsample is applied there: T1[b, oy, qx, c] = sum_ox ux(ox, qx) P1[ox, c].
//                   Per-CTA partial sums (t_c, head-weight / bias gradients, BCE sums) are written, not accumulated.
//   reduce kernel : per image, sums the row partials in row order -> tsum[b, c] and per-image head sums.
//   pack kernel   : per token row q: A = sum_oy uy(oy, qy) T1[b, oy, qx, :] (y-adjoint, <= 2/scale + 1 terms) and
//                   E = (U^T U D_in)[q] as a 3x3 stencil (U^T U of a 1-D bilinear map is tridiagonal), then
//                   dD = A - r^2 t E, written transposed in bf16 for the tensor-core contraction (wgrad.cu).
// Round 1/2a: scatter with 8 float4 atomics per lane and pixel (153 us for 16 images, order-dependent rounding).
constexpr int BWD_PART = 260;   // per-row partials: t_c [128] | gw [128] | gb [2] | bce sums [2]
constexpr int BWD_IMG = 132;    // per-image head sums: gw [128] | gb [2] | bce sums [2]
constexpr int BWD_MAXW = 128;   // largest grid side the backward supports
constexpr int BWD_TAPS = 6;     // y-adjoint weights kept in shared memory (37 -> 68 needs at most 5)

constexpr int BWD_WARPS = 8;
constexpr int BWD_GLD = 72;    // pitch of the TF32 Gram rows (B fragments: 4 rows x 8 columns -> 32 banks)
constexpr int BWD_FLD = 132;   // pitch of the TF32 feature rows (A fragments: 8 rows x 4 columns -> 32 banks)
constexpr int BWD_OLD = 136;   // pitch of the O / P1 rows (fp32)
// words of the region that first holds the two TF32 Grams and later the blended input rows + the reduction slots
__host__ __device__ constexpr size_t bwd_region_a_words(int gin_w) {
    const size_t g = 2 * 64 * BWD_GLD, r = (size_t)gin_w * 128 + BWD_WARPS * 32 * 5;
    return g > r ? g : r;
}
__host__ __device__ constexpr size_t bwd_rows_smem_bytes(int gin_w, int out_w) {
    return (bwd_region_a_words(gin_w) + (size_t)out_w * BWD_FLD + (size_t)out_w * BWD_OLD) * 4;
}

// CTA = one output row of one image.
//   phase M: O[px, :] = fh_own[px, :] . G_other for the whole row on the tensor cores (mma.sync TF32, fp32 accumulate):
//            10 (pixel tile, branch) units of 64 MMAs over 8 warps.  (Round 2b first cut: CUDA-core FFMAs with 3-pixel
//            register tiles, ~290 of the ~700 instructions per pixel.)  O only enters the gradient through the small
//            orthogonality term; TF32 (round-to-nearest when staged) changes it by ~1e-4 relative.
//   phase E: one pixel per warp pass, lane = 4 channels: d from the y-blended input rows, sigmoid gates, BCE, P1 written
//            over O in place.
//   phase X: x-adjoint of the upsample from shared memory, partial sums to `part`.
__global__ void __launch_bounds__(BWD_WARPS * 32, 2)
    decoder_bwd_rows_kernel(const float* __restrict__ d_in, const float* __restrict__ sumsq,
                            const float* __restrict__ emb, const float* __restrict__ w_fg,
                            const float* __restrict__ w_bg, const float* __restrict__ fhat,
                            const float* __restrict__ gram, const float* __restrict__ fg,
                            const float* __restrict__ bg, const float* __restrict__ target,
                            const float* __restrict__ dfg, const float* __restrict__ dbg,
                            const float* __restrict__ dortho, float* __restrict__ t1, float* __restrict__ part, int B,
                            int gin_h, int gin_w, int out_h, int out_w) {
    extern __shared__ __align__(16) float dyn_smem[];
    uint32_t* sGt = reinterpret_cast<uint32_t*>(dyn_smem);              // [2][64][BWD_GLD] (phase M)
    float* rowbuf = dyn_smem;                                            // [gin_w][128]     (phases E, X; aliases sGt)
    float4* red4 = reinterpret_cast<float4*>(dyn_smem + (size_t)gin_w * 128);  // [BWD_WARPS][32]
    float* red1 = reinterpret_cast<float*>(red4 + BWD_WARPS * 32);              // [BWD_WARPS][32]
    uint32_t* sFt = reinterpret_cast<uint32_t*>(dyn_smem + bwd_region_a_words(gin_w));  // [out_w][BWD_FLD] (phase M)
    float* sO = dyn_smem + bwd_region_a_words(gin_w) + (size_t)out_w * BWD_FLD;         // [out_w][BWD_OLD] O, then P1
    __shared__ int s_lo[BWD_MAXW], s_cnt[BWD_MAXW];  // per input column: first contributing output column, how many
    __shared__ int s_x0[BWD_MAXW], s_x1[BWD_MAXW];   // per output column: its two taps
    __shared__ float s_lx[BWD_MAXW];
    __shared__ float s_lg[2][BWD_MAXW], s_up[2][BWD_MAXW];  // this row's logits and targets / upstream gradients (fg, bg)
    const int b = blockIdx.y, oy = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int branch = lane >> 4;
    const int npix = out_h * out_w;
    const float* frow = fhat + ((size_t)b * npix + (size_t)oy * out_w) * 128;
    for (int i = threadIdx.x; i < 2 * out_w; i += BWD_WARPS * 32) {  // per-pixel scalars: no global latency in phase E
        const int br = i >= out_w, ox = i - br * out_w;
        const size_t pix = (size_t)b * npix + (size_t)oy * out_w + ox;
        s_lg[br][ox] = br == 0 ? fg[pix] : bg[pix];
        if (dfg != nullptr)
            s_up[br][ox] = br == 0 ? dfg[pix] : dbg[pix];
        else
            s_up[br][ox] = br == 0 ? target[pix] : 1.f - target[pix];
    }
    for (int i = threadIdx.x; i < 2 * 4096; i += BWD_WARPS * 32)
        sGt[(i >> 6) * BWD_GLD + (i & 63)] = to_tf32(gram[(size_t)b * 8192 + i]);
    for (int i = threadIdx.x; i < out_w * 32; i += BWD_WARPS * 32) {
        const float4 v = reinterpret_cast<const float4*>(frow)[i];
        *reinterpret_cast<uint4*>(sFt + (i >> 5) * BWD_FLD + 4 * (i & 31)) =
                make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    }
    for (int qx = threadIdx.x; qx < gin_w; qx += BWD_WARPS * 32) tap_range(qx, gin_w, out_w, s_lo[qx], s_cnt[qx]);
    for (int ox = threadIdx.x; ox < out_w; ox += BWD_WARPS * 32) bilinear_tap(ox, gin_w, out_w, s_x0[ox], s_x1[ox], s_lx[ox]);
    __syncthreads();

    // ---- phase M ----
    {
        const int g = lane >> 2, t = lane & 3;
        const int m_tiles = (out_w + 15) / 16;
        for (int u = warp; u < 2 * m_tiles; u += BWD_WARPS) {
            const int mt = u % m_tiles, br = u / m_tiles;
            const int ra = min(mt * 16 + g, out_w - 1), rb = min(mt * 16 + g + 8, out_w - 1);  // clamped tail rows
            const uint32_t* fa = sFt + ra * BWD_FLD + br * 64;
            const uint32_t* fb = sFt + rb * BWD_FLD + br * 64;
            const uint32_t* G = sGt + (br ^ 1) * 64 * BWD_GLD;
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
            for (int k0 = 0; k0 < 64; k0 += 8) {
                const uint32_t a[4] = {fa[k0 + t], fb[k0 + t], fa[k0 + t + 4], fb[k0 + t + 4]};
                const uint32_t* g0 = G + (k0 + t) * BWD_GLD + g;
                const uint32_t* g1 = G + (k0 + t + 4) * BWD_GLD + g;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(acc[nt], a, g0[nt * 8], g1[nt * 8]);
            }
            const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (r0 < out_w)
                    *reinterpret_cast<float2*>(sO + r0 * BWD_OLD + br * 64 + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
                if (r1 < out_w)
                    *reinterpret_cast<float2*>(sO + r1 * BWD_OLD + br * 64 + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
            }
        }
    }
    __syncthreads();  // the Grams are dead: their region now takes the blended input rows

    const float* d_img = d_in + (size_t)b * gin_h * gin_w * 128;
    {   // same arithmetic as decoder_head_kernel, so d is reproduced bit for bit
        int y0, y1;
        float ly;
        bilinear_tap(oy, gin_h, out_h, y0, y1, ly);
        const float hy = 1.f - ly;
        const float4* ra = reinterpret_cast<const float4*>(d_img + (size_t)y0 * gin_w * 128);
        const float4* rb = reinterpret_cast<const float4*>(d_img + (size_t)y1 * gin_w * 128);
        for (int i = threadIdx.x; i < gin_w * 32; i += BWD_WARPS * 32) {
            const float4 u = ra[i], v = rb[i];
            reinterpret_cast<float4*>(rowbuf)[i] =
                    make_float4(hy * u.x + ly * v.x, hy * u.y + ly * v.y, hy * u.z + ly * v.z, hy * u.w + ly * v.w);
        }
    }
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
    float4 gw = make_float4(0.f, 0.f, 0.f, 0.f), tacc = make_float4(0.f, 0.f, 0.f, 0.f);
    float gb = 0.f, lacc = 0.f;
    __syncthreads();

    // ---- phase E ----  (the features of the next pixel are requested before the current one is worked on)
    float4 ff_next = warp < out_w ? reinterpret_cast<const float4*>(frow + (size_t)warp * 128)[lane]
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ox = warp; ox < out_w; ox += BWD_WARPS) {
        const float4 ff = ff_next;
        if (ox + BWD_WARPS < out_w) ff_next = reinterpret_cast<const float4*>(frow + (size_t)(ox + BWD_WARPS) * 128)[lane];
        const float lx = s_lx[ox], hx = 1.f - lx;
        const float4 u = reinterpret_cast<const float4*>(rowbuf + (size_t)s_x0[ox] * 128)[lane];
        const float4 v = reinterpret_cast<const float4*>(rowbuf + (size_t)s_x1[ox] * 128)[lane];
        const float4 dd = make_float4(hx * u.x + lx * v.x, hx * u.y + lx * v.y, hx * u.z + lx * v.z, hx * u.w + lx * v.w);
        float4 fo;  // other branch, same column
        fo.x = __shfl_xor_sync(0xffffffffu, ff.x, 16), fo.y = __shfl_xor_sync(0xffffffffu, ff.y, 16);
        fo.z = __shfl_xor_sync(0xffffffffu, ff.z, 16), fo.w = __shfl_xor_sync(0xffffffffu, ff.w, 16);
        const float sp = 0.5f * warp_sum(ff.x * fo.x + ff.y * fo.y + ff.z * fo.z + ff.w * fo.w);
        float4* oslot = reinterpret_cast<float4*>(sO + (size_t)ox * BWD_OLD) + lane;
        float4 Ot = *oslot;
        Ot.x -= sp * fo.x, Ot.y -= sp * fo.y, Ot.z -= sp * fo.z, Ot.w -= sp * fo.w;
        const float logit = s_lg[branch][ox];
        float dlog;
        if (dfg != nullptr) {  // upstream gradients given (autograd entry)
            dlog = s_up[branch][ox];
        } else {               // fused BCE-with-logits against the merged pseudo label
            const float tt = s_up[branch][ox];
            // one exponential serves the sigmoid and the softplus; evaluated by every lane (no divergent slow path)
            const float en = __expf(-fabsf(logit)), inv1 = __fdividef(1.f, 1.f + en);
            const float sig = logit >= 0.f ? inv1 : en * inv1;
            dlog = (sig - tt) * inv_n;
            const float term = fmaxf(logit, 0.f) - logit * tt + __logf(1.f + en);
            if ((lane & 15) == 0) lacc += term;
        }
        float4 sg, a, da, dfh, p1;
        sg.x = __fdividef(1.f, 1.f + __expf(-ff.x * dd.x)), sg.y = __fdividef(1.f, 1.f + __expf(-ff.y * dd.y));
        sg.z = __fdividef(1.f, 1.f + __expf(-ff.z * dd.z)), sg.w = __fdividef(1.f, 1.f + __expf(-ff.w * dd.w));
        a.x = sg.x + dd.x, a.y = sg.y + dd.y, a.z = sg.z + dd.z, a.w = sg.w + dd.w;
        da.x = dlog * wh.x, da.y = dlog * wh.y, da.z = dlog * wh.z, da.w = dlog * wh.w;
        gw.x += dlog * a.x, gw.y += dlog * a.y, gw.z += dlog * a.z, gw.w += dlog * a.w;
        if ((lane & 15) == 0) gb += dlog;
        const float4 ds = make_float4(sg.x * (1.f - sg.x), sg.y * (1.f - sg.y), sg.z * (1.f - sg.z), sg.w * (1.f - sg.w));
        dfh.x = da.x * ds.x * dd.x + oc * Ot.x, dfh.y = da.y * ds.y * dd.y + oc * Ot.y;
        dfh.z = da.z * ds.z * dd.z + oc * Ot.z, dfh.w = da.w * ds.w * dd.w + oc * Ot.w;
        tacc.x += dfh.x * ff.x, tacc.y += dfh.y * ff.y, tacc.z += dfh.z * ff.z, tacc.w += dfh.w * ff.w;
        p1.x = da.x * (1.f + ds.x * ff.x) + r.x * dfh.x, p1.y = da.y * (1.f + ds.y * ff.y) + r.y * dfh.y;
        p1.z = da.z * (1.f + ds.z * ff.z) + r.z * dfh.z, p1.w = da.w * (1.f + ds.w * ff.w) + r.w * dfh.w;
        *oslot = p1;  // P1 over O, same lane, same slot
    }
    float* sP1 = sO;
    // fixed-order block reduction of the per-lane accumulators -> this row's slot of `part`
    float* prow = part + ((size_t)b * out_h + oy) * BWD_PART;
    red4[warp * 32 + lane] = tacc;
    red1[warp * 32 + lane] = (lane & 15) == 0 ? gb : 0.f;
    __syncthreads();  // also publishes sP1
    if (warp == 0) {
        float4 t = red4[lane];
        float g1 = red1[lane];
        for (int w = 1; w < BWD_WARPS; ++w) {
            t.x += red4[w * 32 + lane].x, t.y += red4[w * 32 + lane].y, t.z += red4[w * 32 + lane].z, t.w += red4[w * 32 + lane].w;
            g1 += red1[w * 32 + lane];
        }
        reinterpret_cast<float4*>(prow)[lane] = t;
        if ((lane & 15) == 0) prow[256 + (lane >> 4)] = g1;
    }
    __syncthreads();
    red4[warp * 32 + lane] = gw;
    red1[warp * 32 + lane] = lacc;
    __syncthreads();
    if (warp == 0) {
        float4 t = red4[lane];
        float l = red1[lane];
        for (int w = 1; w < BWD_WARPS; ++w) {
            t.x += red4[w * 32 + lane].x, t.y += red4[w * 32 + lane].y, t.z += red4[w * 32 + lane].z, t.w += red4[w * 32 + lane].w;
            l += red1[w * 32 + lane];
        }
        reinterpret_cast<float4*>(prow + 128)[lane] = t;
        if ((lane & 15) == 0) prow[258 + (lane >> 4)] = l;
    }
    // x-adjoint of the upsample inside shared memory, ascending output column => fixed summation order
    float* trow = t1 + ((size_t)b * out_h + oy) * gin_w * 128;
    for (int idx = threadIdx.x; idx < gin_w * 32; idx += BWD_WARPS * 32) {
        const int qx = idx >> 5, l4 = idx & 31;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int lo = s_lo[qx], n = s_cnt[qx];
        for (int k = 0; k < n; ++k) {
            const int ox = lo + k;
            const float w = (s_x0[ox] == qx ? 1.f - s_lx[ox] : 0.f) + (s_x1[ox] == qx ? s_lx[ox] : 0.f);
            const float4 v = reinterpret_cast<const float4*>(sP1 + (size_t)ox * BWD_OLD)[l4];
            acc.x += w * v.x, acc.y += w * v.y, acc.z += w * v.z, acc.w += w * v.w;
        }
        reinterpret_cast<float4*>(trow + (size_t)qx * 128)[l4] = acc;
    }
}

// per image: row partials summed in row order -> tsum [B,128] and the per-image head sums [B, BWD_IMG]
__global__ void __launch_bounds__(288)
    decoder_bwd_reduce_kernel(const float* __restrict__ part, float* __restrict__ tsum, float* __restrict__ img_part,
                              int out_h) {
    const int b = blockIdx.x, t = threadIdx.x;
    if (t >= BWD_PART) return;
    const float* p = part + (size_t)b * out_h * BWD_PART + t;
    float s = 0.f;
    int oy = 0;
    for (; oy + 8 <= out_h; oy += 8) {  // 8 loads in flight, added in row order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[(size_t)(oy + u) * BWD_PART];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; oy < out_h; ++oy) s += p[(size_t)oy * BWD_PART];
    if (t < 128)
        tsum[(size_t)b * 128 + t] = s;
    else
        img_part[(size_t)b * BWD_IMG + (t - 128)] = s;
}

// dD = A - (r_c^2 t_c) E for 64 token rows per CTA, written transposed as bf16 (dDt [128, t_pad], zero beyond T) for the
// tensor-core contraction; bpart[blk, c] = this block's column sums (db_dec, reduced later in a fixed order).
// A = y-adjoint of T1, E = (U^T U D_in) as a 3x3 stencil with the tridiagonal 1-D factors My, Mx.
// Block 0 also finishes the head-weight / bias gradients and the two BCE means (sum over images, image order).
__global__ void __launch_bounds__(256)
    decoder_bwd_pack_kernel(const float* __restrict__ t1, const float* __restrict__ d_in,
                            const float* __restrict__ tsum, const float* __restrict__ sumsq,
                            const float* __restrict__ emb, const float* __restrict__ img_part,
                            __nv_bfloat16* __restrict__ dDt, float* __restrict__ bpart, float* __restrict__ g_wfg,
                            float* __restrict__ g_bfg, float* __restrict__ g_wbg, float* __restrict__ g_bbg,
                            float* __restrict__ loss2, int B, int gin_h, int gin_w, int out_h, int out_w,
                            int total_rows, int t_pad) {
    __shared__ __align__(16) float tile[64][132];
    __shared__ __align__(16) float colsum[8][128];
    __shared__ int s_ylo[BWD_MAXW], s_ycnt[BWD_MAXW];
    __shared__ float s_wy[BWD_MAXW][BWD_TAPS];  // weights of the first BWD_TAPS contributing output rows
    __shared__ float s_my[BWD_MAXW][3], s_mx[BWD_MAXW][3];
    const int r0 = blockIdx.x * 64;
    const int P = gin_h * gin_w;
    {   // tables for the grid rows this 64-token block touches (it may run over the end of an image: modulo gin_h)
        const int qy_first = (r0 % P) / gin_w;
        const int n_qy = 63 / gin_w + 2 < gin_h ? 63 / gin_w + 2 : gin_h;
        for (int j = threadIdx.x; j < n_qy; j += 256) {
            const int q = (qy_first + j) % gin_h;
            int lo, cnt;
            tap_range(q, gin_h, out_h, lo, cnt);
            s_ylo[q] = lo, s_ycnt[q] = cnt;
            for (int k = 0; k < BWD_TAPS; ++k) s_wy[q][k] = k < cnt ? tap_weight(lo + k, q, gin_h, out_h) : 0.f;
            gram_row_1d(q, gin_h, out_h, s_my[q]);
        }
        for (int q = (int)threadIdx.x - 64; q >= 0 && q < gin_w; q += 192) gram_row_1d(q, gin_w, out_w, s_mx[q]);
    }
    __syncthreads();
    // warp = 8 consecutive token rows, lane = 4 channels: every load is one coalesced 512-byte row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 e4 = __ldg(reinterpret_cast<const float4*>(emb) + lane);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), coef = acc;
    int cur_b = -1;
#pragma unroll 1
    for (int i = warp * 8; i < warp * 8 + 8; ++i) {
        const int row = r0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < total_rows) {
            const int b = row / P, q = row - b * P;
            const int qy = q / gin_w, qx = q - qy * gin_w;
            if (b != cur_b) {  // r_c^2 t_c of this image
                const float4 sq = reinterpret_cast<const float4*>(sumsq + (size_t)b * 128)[lane];
                const float4 ts = reinterpret_cast<const float4*>(tsum + (size_t)b * 128)[lane];
                float g;
                g = e4.x / fmaxf(fabsf(e4.x) * sqrtf(sq.x), 1e-12f), coef.x = g * g * ts.x;
                g = e4.y / fmaxf(fabsf(e4.y) * sqrtf(sq.y), 1e-12f), coef.y = g * g * ts.y;
                g = e4.z / fmaxf(fabsf(e4.z) * sqrtf(sq.z), 1e-12f), coef.z = g * g * ts.z;
                g = e4.w / fmaxf(fabsf(e4.w) * sqrtf(sq.w), 1e-12f), coef.w = g * g * ts.w;
                cur_b = b;
            }
            float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
            const int lo = s_ylo[qy], n = s_ycnt[qy];
            const float4* tcol = reinterpret_cast<const float4*>(t1 + (((size_t)b * out_h + lo) * gin_w + qx) * 128) + lane;
            const size_t tstride = (size_t)gin_w * 32;
#pragma unroll
            for (int k = 0; k < BWD_TAPS; ++k)
                if (k < n) {
                    const float w = s_wy[qy][k];
                    const float4 t = tcol[k * tstride];
                    A.x += w * t.x, A.y += w * t.y, A.z += w * t.z, A.w += w * t.w;
                }
            for (int k = BWD_TAPS; k < n; ++k) {
                const float w = tap_weight(lo + k, qy, gin_h, out_h);
                const float4 t = tcol[k * tstride];
                A.x += w * t.x, A.y += w * t.y, A.z += w * t.z, A.w += w * t.w;
            }
            float4 E = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* d_img = reinterpret_cast<const float4*>(d_in + (size_t)b * P * 128) + lane;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = qy + dy;
                if (yy < 0 || yy >= gin_h) continue;
                float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = qx + dx;
                    if (xx < 0 || xx >= gin_w) continue;
                    const float m = s_mx[qx][dx + 1];
                    const float4 t = d_img[((size_t)yy * gin_w + xx) * 32];
                    rs.x += m * t.x, rs.y += m * t.y, rs.z += m * t.z, rs.w += m * t.w;
                }
                const float m = s_my[qy][dy + 1];
                E.x += m * rs.x, E.y += m * rs.y, E.z += m * rs.z, E.w += m * rs.w;
            }
            v = make_float4(A.x - coef.x * E.x, A.y - coef.y * E.y, A.z - coef.z * E.z, A.w - coef.w * E.w);
        }
        reinterpret_cast<float4*>(tile[i])[lane] = v;
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    reinterpret_cast<float4*>(colsum[warp])[lane] = acc;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int c = threadIdx.x;
        float t = colsum[0][c];
        for (int w = 1; w < 8; ++w) t += colsum[w][c];
        bpart[(size_t)blockIdx.x * 128 + c] = t;
    }
    // transposed store: thread -> (channel cc, 32-token half hh): 32 bf16 = 64 bytes
    const int cc = threadIdx.x >> 1, hh = threadIdx.x & 1;
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(tile[hh * 32 + 2 * i][cc], tile[hh * 32 + 2 * i + 1][cc]);
    uint4* dst = reinterpret_cast<uint4*>(dDt + (size_t)cc * t_pad + r0 + hh * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    if (blockIdx.x == 0 && threadIdx.x < BWD_IMG) {
        const int t = threadIdx.x;
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += img_part[(size_t)b * BWD_IMG + t];
        if (t < 64)
            g_wfg[t] = s;
        else if (t < 128)
            g_wbg[t - 64] = s;
        else if (t == 128)
            g_bfg[0] = s;
        else if (t == 129)
            g_bbg[0] = s;
        else
            loss2[t - 130] = s / ((float)B * (float)(out_h * out_w));
    }
}

static size_t bwd_front_bytes(int B, int gin_w, int out_h) {  // T1 | row partials | tsum | per-image sums, 256-aligned
    const size_t n = (size_t)B * out_h * gin_w * 128 + (size_t)B * out_h * BWD_PART + (size_t)B * 128 + (size_t)B * BWD_IMG;
    return (n * 4 + 255) / 256 * 256;
}

size_t decoder_backward_workspace_bytes(int B, int gin_h, int gin_w, int out_h, int out_w) {
    (void)out_w;
    const size_t rows = (size_t)B * gin_h * gin_w;
    // dDt (bf16, transposed), split-K partial tiles, column-sum partials (wgrad.cu); dim <= 1024
    size_t wg = 0;
    for (int dim = 256; dim <= 1024; dim += 256) {
        const size_t n = wgrad_workspace_bytes((int)rows, dim);
        wg = n > wg ? n : wg;
    }
    return bwd_front_bytes(B, gin_w, out_h) + wg + 4096;
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
    UCOD_REQUIRE(ws_bytes >= decoder_backward_workspace_bytes(B, gin_h, gin_w, out_h, out_w),
                 "decoder_backward: workspace too small");
    UCOD_REQUIRE(gin_h <= BWD_MAXW && gin_w <= BWD_MAXW && out_h <= BWD_MAXW && out_w <= BWD_MAXW,
                 "decoder_backward: grids up to %d x %d", BWD_MAXW, BWD_MAXW);
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
    float* t1 = static_cast<float*>(workspace);
    float* part = t1 + (size_t)B * out_h * gin_w * 128;
    float* tsum = part + (size_t)B * out_h * BWD_PART;
    float* img_part = tsum + (size_t)B * 128;
    const size_t front = bwd_front_bytes(B, gin_w, out_h);
    uint8_t* wg_ws = static_cast<uint8_t*>(workspace) + front;
    WgradPlan plan;
    if (int rc = wgrad_plan((int)rows, w.dim, wg_ws, ws_bytes - front, &plan)) return rc;
    // every buffer below is fully written by its producer: nothing to zero
    {
        const size_t smem = bwd_rows_smem_bytes(gin_w, out_w);
        static size_t configured = 0;
        if (smem > configured) {
            UCOD_CHECK_CUDA(cudaFuncSetAttribute(decoder_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem));
            configured = smem;
        }
        ProfScope ps(KC_DECODER, stream, (double)B * npix * 128 * 4 + (double)rows * 128 * 4 + (double)B * out_h * gin_w * 512);
        decoder_bwd_rows_kernel<<<dim3(out_h, B), BWD_WARPS * 32, smem, stream>>>(d_in, sumsq, w.emb, w.w_fg, w.w_bg, fhat, gram, fg,
                                                                       bg, target, dfg, dbg, dortho, t1, part, B, gin_h,
                                                                       gin_w, out_h, out_w);
    }
    decoder_bwd_reduce_kernel<<<B, 288, 0, stream>>>(part, tsum, img_part, out_h);
    {
        ProfScope ps(KC_DECODER, stream, (double)B * out_h * gin_w * 512 + (double)rows * 128 * (4 + 2));
        decoder_bwd_pack_kernel<<<plan.n_kblocks, 256, 0, stream>>>(t1, d_in, tsum, sumsq, w.emb, img_part, plan.dDt,
                                                                    plan.bpart, g.w_fg, g.b_fg, g.w_bg, g.b_bg, loss2, B,
                                                                    gin_h, gin_w, out_h, out_w, (int)rows, plan.t_pad);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return wgrad_contract(plan, keys_bf16, (int)rows, w.dim, g.w_dec, g.b_dec, stream);
}

// loss = bce_fg + bce_bg + ortho - dis_loss (loop_UCOD_DPL.py:176-180; dis_loss == NULL in the finetune epochs)
__global__ void train_loss_kernel(const float* __restrict__ loss2, const float* __restrict__ ortho,
                                  const float* __restrict__ dis_loss, float* __restrict__ out) {
    float v = (loss2[0] + loss2[1]) + ortho[0];
    if (dis_loss != nullptr) v -= dis_loss[0];
    out[0] = v;
}
int train_loss(const float* loss2, const float* ortho, const float* dis_loss, float* out, cudaStream_t stream) {
    UCOD_REQUIRE(loss2 && ortho && out, "train_loss: null argument");
    train_loss_kernel<<<1, 1, 0, stream>>>(loss2, ortho, dis_loss, out);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
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
// (fp32 sigmoid(x) > 0.5  <=>  x > 1.5 * 2^-24, probed on torch CPU; pinned in tests/test_oracle_golden.py::test_sigmoid_half_threshold_is_pinned).
// One thread = 4 consecutive output pixels (32-bit store of 4 mask bytes / float4 store).
// ------------------------------------------------------------------------------------------------
#define UCOD_SIGMOID_HALF_THRESHOLD 0x1.8p-24f

constexpr int UP_ROWS = 8;       // output rows per CTA
constexpr int UP_MAXW = 16384;   // widest output the column-tap table in shared memory covers (12 bytes per column)
template <int MODE>
__global__ void __launch_bounds__(256)
    upsample_bilinear_kernel(const float* __restrict__ in, void* __restrict__ out, int in_h, int in_w, int out_h,
                             int out_w) {
    // CTA = UP_ROWS output rows of one image.  The column taps are computed once per CTA (a float division each) into
    // shared memory; round 2a recomputed five taps per thread for four pixels and was issue-bound (90 % issue active,
    // 0.3 TB/s).  One thread = 4 consecutive output pixels, packed store.
    extern __shared__ __align__(16) uint8_t up_smem[];
    int* s_x0 = reinterpret_cast<int*>(up_smem);
    int* s_x1 = s_x0 + out_w;
    float* s_lx = reinterpret_cast<float*>(s_x1 + out_w);
    const int b = blockIdx.y;
    for (int ox = threadIdx.x; ox < out_w; ox += blockDim.x) bilinear_tap(ox, in_w, out_w, s_x0[ox], s_x1[ox], s_lx[ox]);
    __syncthreads();
    const float* src = in + (size_t)b * in_h * in_w;
    const int oy_end = min(out_h, ((int)blockIdx.x + 1) * UP_ROWS);
    for (int oy = blockIdx.x * UP_ROWS; oy < oy_end; ++oy) {
        int y0, y1;
        float ly;
        bilinear_tap(oy, in_h, out_h, y0, y1, ly);
        const float hy = 1.f - ly;
        const float* r0 = src + y0 * in_w;
        const float* r1 = src + y1 * in_w;
        for (int ox0 = threadIdx.x * 4; ox0 < out_w; ox0 += blockDim.x * 4) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ox = ox0 + i < out_w ? ox0 + i : out_w - 1;
                const int x0 = s_x0[ox], x1 = s_x1[ox];
                const float lx = s_lx[ox], hx = 1.f - lx;
                float t00 = __ldg(r0 + x0), t01 = __ldg(r0 + x1), t10 = __ldg(r1 + x0), t11 = __ldg(r1 + x1);
                if constexpr (MODE == 2) {  // probabilities first, then interpolate (loop_CORAL.py:331-338)
                    t00 = 1.f / (1.f + expf(-t00)), t01 = 1.f / (1.f + expf(-t01));
                    t10 = 1.f / (1.f + expf(-t10)), t11 = 1.f / (1.f + expf(-t11));
                }
                v[i] = hy * (hx * t00 + lx * t01) + ly * (hx * t10 + lx * t11);
            }
            const size_t o = ((size_t)b * out_h + oy) * out_w + ox0;
            const int nv = min(4, out_w - ox0);
            if constexpr (MODE == 0) {
                float* dst = static_cast<float*>(out) + o;
                if (nv == 4 && (o & 3) == 0) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    for (int i = 0; i < nv; ++i) dst[i] = v[i];
                }
            } else {
                uint8_t* dst = static_cast<uint8_t*>(out) + o;
                uint8_t m[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) m[i] = (MODE == 1 ? v[i] > UCOD_SIGMOID_HALF_THRESHOLD : v[i] > 0.5f) ? 1 : 0;
                if (nv == 4 && (o & 3) == 0) {
                    *reinterpret_cast<uint32_t*>(dst) = (uint32_t)m[0] | ((uint32_t)m[1] << 8) | ((uint32_t)m[2] << 16) | ((uint32_t)m[3] << 24);
                } else if (nv == 4 && (o & 1) == 0) {
                    reinterpret_cast<uint16_t*>(dst)[0] = (uint16_t)(m[0] | (m[1] << 8));
                    reinterpret_cast<uint16_t*>(dst)[1] = (uint16_t)(m[2] | (m[3] << 8));
                } else {
                    for (int i = 0; i < nv; ++i) dst[i] = m[i];
                }
            }
        }
    }
}

template <int MODE>
static int launch_upsample(const float* in, void* out, int B, int in_h, int in_w, int out_h, int out_w,
                           cudaStream_t stream) {
    const size_t smem = (size_t)out_w * 12;
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        UCOD_CHECK_CUDA(cudaFuncSetAttribute(upsample_bilinear_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        configured = smem;
    }
    upsample_bilinear_kernel<MODE><<<dim3(ceil_div(out_h, UP_ROWS), B), 256, smem, stream>>>(in, out, in_h, in_w, out_h,
                                                                                                 out_w);
    return 0;
}

int upsample_bilinear(const float* in, void* out, int B, int in_h, int in_w, int out_h, int out_w, int binarize,
                      cudaStream_t stream) {
    UCOD_REQUIRE(in && out && B > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "upsample: bad argument");
    UCOD_REQUIRE(out_w <= UP_MAXW, "upsample: outputs up to %d pixels wide", UP_MAXW);
    ProfScope ps(KC_RESAMPLE, stream, (double)B * in_h * in_w * 4 + (double)B * out_h * out_w * (binarize ? 1 : 4));
    UCOD_REQUIRE(binarize >= 0 && binarize <= 3, "upsample: binarize mode %d unknown", binarize);
    if (binarize == 1)
        launch_upsample<1>(in, out, B, in_h, in_w, out_h, out_w, stream);
    else if (binarize == 2)
        launch_upsample<2>(in, out, B, in_h, in_w, out_h, out_w, stream);
    else if (binarize == 3)
        launch_upsample<3>(in, out, B, in_h, in_w, out_h, out_w, stream);
    else
        launch_upsample<0>(in, out, B, in_h, in_w, out_h, out_w, stream);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
