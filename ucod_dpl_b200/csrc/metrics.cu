// COD metric suite on the device, fp64 (reference engine/utils/metrics/metric.py:19-74,128-531,
// numpy/scipy float64 on the CPU: `_prepare_data`, ACC, IoU, MAE, S-measure, E-measure (adaptive + 256-threshold
// curve), F-measure (adaptive + curve) and the weighted F-measure with its exact Euclidean feature transform).
// Once the model path runs at ~0.35 ms / image the CPU suite (~50-100 ms / image) is >99 % of the eval wall time.
//
// One launch sequence handles a batch of equally sized (gt, pred) pairs; every image is independent.
//   1. min/max of gt and pred                                  (normalisation of `_prepare_data`)
//   2. one pass of per-pixel sums + two 256-bin histograms     (ACC, IoU, MAE, S-object, centroid, curves)
//   3. centroid / adaptive threshold, then quadrant sums       (S-region SSIM, adaptive E / F)
//   4. exact Euclidean feature transform with scipy's tie rule (smallest x, then smallest y among the nearest
//      foreground pixels — probed against scipy.ndimage.distance_transform_edt), 7x7 Gaussian, weighted sums
//   5. one CTA per image turns the accumulators into the measures and the two 256-point curves.
#include "metrics.cuh"

#include "prof.cuh"

namespace ucod {

namespace {

constexpr double M_EPS = 2.220446049250313e-16;  // np.spacing(1)

// per-image accumulator block (doubles)
enum Acc : int {
    A_GMIN = 0, A_GMAX, A_PMIN, A_PMAX,          // written as doubles by the min/max kernel
    A_SUM_P, A_CNT_G, A_SUM_ABS, A_EQ, A_AND, A_OR,
    A_FG_P, A_FG_P2, A_BG_Q, A_BG_Q2, A_SUM_Y, A_SUM_X,
    A_CX, A_CY, A_THR,                            // derived
    A_QUAD,                                       // 4 quadrants x {n, sp, sg, sp2, spg} = 20
    A_ADP_B = A_QUAD + 20, A_ADP_BG,              // #(p >= thr), #(p >= thr & g)
    A_W_SUM_G, A_W_EW_G, A_W_EW_BG,               // weighted F sums
    A_COUNT
};
constexpr int ACC_STRIDE = 64;

__device__ __forceinline__ double block_reduce(double v, double* red, int op /*0 sum,1 min,2 max*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, v, o);
        v = op == 0 ? v + t : (op == 1 ? fmin(v, t) : fmax(v, t));
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = op == 0 ? r + red[i] : (op == 1 ? fmin(r, red[i]) : fmax(r, red[i]));
    return r;
}

__global__ void __launch_bounds__(1024) metrics_minmax_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                             int n, double* __restrict__ acc) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    double gmin = 1e300, gmax = -1e300, pmin = 1e300, pmax = -1e300;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double g = gt[(size_t)b * n + i], p = pred[(size_t)b * n + i];
        gmin = fmin(gmin, g), gmax = fmax(gmax, g), pmin = fmin(pmin, p), pmax = fmax(pmax, p);
    }
    gmin = block_reduce(gmin, red, 1), gmax = block_reduce(gmax, red, 2);
    pmin = block_reduce(pmin, red, 1), pmax = block_reduce(pmax, red, 2);
    if (threadIdx.x == 0) {
        double* a = acc + (size_t)b * ACC_STRIDE;
        a[A_GMIN] = gmin, a[A_GMAX] = gmax, a[A_PMIN] = pmin, a[A_PMAX] = pmax;
    }
}

// `_prepare_data` (metric.py:128-136)
__device__ __forceinline__ void prep(const double* a, float graw, float praw, double& p, bool& g) {
    const double gmin = a[A_GMIN], gmax = a[A_GMAX], pmin = a[A_PMIN], pmax = a[A_PMAX];
    double gv = graw;
    if (gmax != gmin) gv = (gv - gmin) / (gmax - gmin);
    g = gv > 0.5;
    if (pmax != pmin) p = ((double)praw - pmin) / (pmax - pmin);
    else p = (double)(long long)praw;  // astype(int): truncation
}

__global__ void __launch_bounds__(256) metrics_stats_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                           int h, int w, double* __restrict__ acc,
                                                           unsigned int* __restrict__ hist /*[B][2][256]*/) {
    __shared__ double red[8];
    __shared__ unsigned int s_h[512];
    const int b = blockIdx.y, n = h * w;
    const double* a = acc + (size_t)b * ACC_STRIDE;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double p;
        bool g;
        prep(a, gt[(size_t)b * n + i], pred[(size_t)b * n + i], p, g);
        const double gd = g ? 1.0 : 0.0;
        v[0] += p, v[1] += gd, v[2] += fabs(p - gd), v[3] += (p == gd) ? 1.0 : 0.0;
        v[4] += (p != 0.0 && g) ? 1.0 : 0.0, v[5] += (p != 0.0 || g) ? 1.0 : 0.0;
        if (g) {
            v[6] += p, v[7] += p * p, v[10] += (double)(i / w), v[11] += (double)(i % w);
        } else {
            v[8] += 1.0 - p, v[9] += (1.0 - p) * (1.0 - p);
        }
        const int q = (int)(unsigned char)(p * 255.0);  // (pred * 255).astype(np.uint8)
        atomicAdd(&s_h[(g ? 0 : 256) + q], 1u);
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const double r = block_reduce(v[k], red, 0);
        if (threadIdx.x == 0 && r != 0.0) atomicAdd(acc + (size_t)b * ACC_STRIDE + A_SUM_P + k, r);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x)
        if (s_h[i]) atomicAdd(hist + (size_t)b * 512 + i, s_h[i]);
}

__global__ void metrics_derive_kernel(double* __restrict__ acc, int B, int h, int w) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double* a = acc + (size_t)b * ACC_STRIDE;
    const double cnt = a[A_CNT_G];
    double cx, cy;
    if (cnt == 0) cx = rint((double)w / 2), cy = rint((double)h / 2);
    else cy = rint(a[A_SUM_Y] / cnt), cx = rint(a[A_SUM_X] / cnt);  // np.round: half to even
    a[A_CX] = (double)((int)cx + 1), a[A_CY] = (double)((int)cy + 1);
    a[A_THR] = fmin(2.0 * a[A_SUM_P] / ((double)h * w), 1.0);
}

__global__ void __launch_bounds__(256) metrics_region_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                            int h, int w, double* __restrict__ acc) {
    __shared__ double red[8];
    const int b = blockIdx.y, n = h * w;
    const double* a = acc + (size_t)b * ACC_STRIDE;
    const int cx = (int)a[A_CX], cy = (int)a[A_CY];
    const double thr = a[A_THR];
    double v[22];
#pragma unroll
    for (int k = 0; k < 22; ++k) v[k] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double p;
        bool g;
        prep(a, gt[(size_t)b * n + i], pred[(size_t)b * n + i], p, g);
        const int y = i / w, x = i - y * w;
        const int q = (y < cy ? 0 : 2) + (x < cx ? 0 : 1);
        const double gd = g ? 1.0 : 0.0;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
            if (q == qq) {
                v[qq * 5 + 0] += 1.0, v[qq * 5 + 1] += p, v[qq * 5 + 2] += gd, v[qq * 5 + 3] += p * p, v[qq * 5 + 4] += p * gd;
            }
        if (p >= thr) {
            v[20] += 1.0;
            if (g) v[21] += 1.0;
        }
    }
#pragma unroll
    for (int k = 0; k < 22; ++k) {
        const double r = block_reduce(v[k], red, 0);
        if (threadIdx.x == 0 && r != 0.0) atomicAdd(acc + (size_t)b * ACC_STRIDE + A_QUAD + k, r);
    }
}

// ---- exact Euclidean feature transform of the foreground set (nearest fg pixel of every pixel) ----
// pass A: per column, vertical distance to the nearest fg pixel of that column (ties: the upper pixel)
__global__ void metrics_edt_cols_kernel(const float* __restrict__ gt, const double* __restrict__ acc, int h, int w,
                                        int* __restrict__ vdist /*[B][h][w] signed: y' - y, INT_MAX = none*/) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (x >= w) return;
    const double* a = acc + (size_t)b * ACC_STRIDE;
    const double gmin = a[A_GMIN], gmax = a[A_GMAX];
    const float* g = gt + (size_t)b * h * w;
    int* vd = vdist + (size_t)b * h * w;
    auto is_fg = [&](int y) {
        double gv = g[(size_t)y * w + x];
        if (gmax != gmin) gv = (gv - gmin) / (gmax - gmin);
        return gv > 0.5;
    };
    int last = -1;
    for (int y = 0; y < h; ++y) {  // nearest fg at or above
        if (is_fg(y)) last = y;
        vd[(size_t)y * w + x] = last < 0 ? 0x7fffffff : last - y;
    }
    last = -1;
    for (int y = h - 1; y >= 0; --y) {  // nearest fg below: take it only if strictly closer
        if (is_fg(y)) last = y;
        if (last >= 0) {
            const int up = vd[(size_t)y * w + x];
            const int dn = last - y;
            if (up == 0x7fffffff || dn < -up) vd[(size_t)y * w + x] = dn;
        }
    }
}
// pass B: per pixel, search the columns outwards; ties -> smaller x', then (from pass A) smaller y'
__global__ void metrics_edt_rows_kernel(const int* __restrict__ vdist, int h, int w, double* __restrict__ dst,
                                        int* __restrict__ feat /*[B][h][w] linear index of the nearest fg pixel or -1*/) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= w) return;
    const int* vd = vdist + ((size_t)b * h + y) * w;
    long long best = 0x7fffffffffffffffll;
    int bx = -1, bdy = 0;
    for (int dx = 0; dx < w; ++dx) {
        if ((long long)dx * dx > best) break;
        const int xs[2] = {x - dx, x + dx};
        for (int s = 0; s < (dx == 0 ? 1 : 2); ++s) {
            const int xc = xs[s];
            if (xc < 0 || xc >= w) continue;
            const int dy = vd[xc];
            if (dy == 0x7fffffff) continue;
            const long long d2 = (long long)dx * dx + (long long)dy * dy;
            if (d2 < best || (d2 == best && xc < bx)) best = d2, bx = xc, bdy = dy;
        }
    }
    const size_t o = ((size_t)b * h + y) * w + x;
    dst[o] = bx < 0 ? 0.0 : sqrt((double)best);
    feat[o] = bx < 0 ? -1 : (y + bdy) * w + bx;
}
// E_t (error carried over from the nearest fg pixel for background pixels)
__global__ void metrics_et_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                  const double* __restrict__ acc, const int* __restrict__ feat, int n,
                                  double* __restrict__ et) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= n) return;
    const double* a = acc + (size_t)b * ACC_STRIDE;
    double p;
    bool g;
    prep(a, gt[(size_t)b * n + i], pred[(size_t)b * n + i], p, g);
    if (!g) {
        const int f = feat[(size_t)b * n + i];
        if (f >= 0) {
            double pf;
            bool gf;
            prep(a, gt[(size_t)b * n + f], pred[(size_t)b * n + f], pf, gf);
            et[(size_t)b * n + i] = fabs(pf - (gf ? 1.0 : 0.0));
            return;
        }
    }
    et[(size_t)b * n + i] = fabs(p - (g ? 1.0 : 0.0));
}
__constant__ double c_gauss7[49];
__global__ void __launch_bounds__(256) metrics_wfm_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                         const double* __restrict__ et, const double* __restrict__ dst,
                                                         int h, int w, double* __restrict__ acc) {
    __shared__ double red[8];
    const int b = blockIdx.y, n = h * w;
    const double* a = acc + (size_t)b * ACC_STRIDE;
    const double* e_t = et + (size_t)b * n;
    double v[3] = {0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double p;
        bool g;
        prep(a, gt[(size_t)b * n + i], pred[(size_t)b * n + i], p, g);
        const int y = i / w, x = i - y * w;
        double ea = 0.0;
        for (int ky = -3; ky <= 3; ++ky) {
            const int yy = y + ky;
            if (yy < 0 || yy >= h) continue;
            for (int kx = -3; kx <= 3; ++kx) {
                const int xx = x + kx;
                if (xx < 0 || xx >= w) continue;
                ea += e_t[(size_t)yy * w + xx] * c_gauss7[(3 - ky) * 7 + (3 - kx)];
            }
        }
        const double e = fabs(p - (g ? 1.0 : 0.0));
        const double mn = (g && ea < e) ? ea : e;
        const double bw = g ? 1.0 : 2.0 - exp(log(0.5) / 5.0 * dst[(size_t)b * n + i]);
        const double ew = mn * bw;
        if (g) v[0] += 1.0, v[1] += ew;
        else v[2] += ew;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double r = block_reduce(v[k], red, 0);
        if (threadIdx.x == 0 && r != 0.0) atomicAdd(acc + (size_t)b * ACC_STRIDE + A_W_SUM_G + k, r);
    }
}

__device__ double em_from_counts(double fg_fg, double fg_bg, double gt_fg, double size) {
    const double pred_fg = fg_fg + fg_bg, pred_bg = size - pred_fg;
    double s;
    if (gt_fg == 0) s = pred_bg;
    else if (gt_fg == size) s = pred_fg;
    else {
        const double bg_fg = gt_fg - fg_fg, bg_bg = pred_bg - bg_fg;
        const double mp = pred_fg / size, mg = gt_fg / size;
        const double ca[4] = {1 - mp, 1 - mp, 0 - mp, 0 - mp}, cb[4] = {1 - mg, 0 - mg, 1 - mg, 0 - mg};
        const double part[4] = {fg_fg, fg_bg, bg_fg, bg_bg};
        s = 0;
        for (int i = 0; i < 4; ++i) {
            const double al = 2 * (ca[i] * cb[i]) / (ca[i] * ca[i] + cb[i] * cb[i] + M_EPS);
            s += ((al + 1) * (al + 1) / 4) * part[i];
        }
    }
    return s / (size - 1 + M_EPS);
}
__device__ double ssim_from_sums(const double* q) {
    const double n = q[0];
    const double x = q[1] / n, y = q[2] / n;
    const double sx = (q[3] - n * x * x) / (n - 1), sy = (q[2] - n * y * y) / (n - 1), sxy = (q[4] - n * x * y) / (n - 1);
    const double al = 4 * x * y * sxy, be = (x * x + y * y) * (sx + sy);
    if (al != 0) return al / (be + M_EPS);
    return be == 0 ? 1.0 : 0.0;
}
__device__ double s_object(double sum, double sum2, double n) {
    const double x = sum / n;
    double var = (sum2 - n * x * x) / (n - 1);
    if (var < 0) var = 0;
    return 2 * x / (x * x + 1 + sqrt(var) + M_EPS);
}

__global__ void __launch_bounds__(256) metrics_finalize_kernel(const double* __restrict__ acc,
                                                              const unsigned int* __restrict__ hist, int h, int w,
                                                              double* __restrict__ out) {
    const int b = blockIdx.x, t = threadIdx.x;
    const double* a = acc + (size_t)b * ACC_STRIDE;
    const unsigned int* hf = hist + (size_t)b * 512;
    double* o = out + (size_t)b * METRICS_OUT;
    const double size = (double)h * w, gt_fg = a[A_CNT_G];
    // cumulative counts for threshold index t: bins 255 .. 255 - t
    double fg = 0, bg = 0;
    for (int k = 255; k >= 255 - t; --k) fg += hf[k], bg += hf[256 + k];
    o[7 + t] = em_from_counts(fg, bg, gt_fg, size);
    {
        double ps = fg + bg;
        if (ps == 0) ps = 1;
        const double T = gt_fg > 1 ? gt_fg : 1;
        const double prec = fg / ps, rec = fg / T;
        const double num = 1.3 * prec * rec;
        const double den = num == 0 ? 1 : 0.3 * prec + rec;
        o[7 + 256 + t] = num / den;
    }
    if (t != 0) return;
    o[0] = a[A_EQ] / size;
    o[1] = a[A_OR] == 0 ? 1.0 : a[A_AND] / a[A_OR];
    o[2] = a[A_SUM_ABS] / size;
    // S-measure
    const double ymean = gt_fg / size, pmean = a[A_SUM_P] / size;
    double sm;
    if (ymean == 0) sm = 1 - pmean;
    else if (ymean == 1) sm = pmean;
    else {
        const double obj = ymean * s_object(a[A_FG_P], a[A_FG_P2], gt_fg) +
                           (1 - ymean) * s_object(a[A_BG_Q], a[A_BG_Q2], size - gt_fg);
        const double cx = a[A_CX], cy = a[A_CY];
        const double w1 = cx * cy / size, w2 = cy * (w - cx) / size, w3 = (h - cy) * cx / size, w4 = 1 - w1 - w2 - w3;
        const double reg = w1 * ssim_from_sums(a + A_QUAD) + w2 * ssim_from_sums(a + A_QUAD + 5) +
                           w3 * ssim_from_sums(a + A_QUAD + 10) + w4 * ssim_from_sums(a + A_QUAD + 15);
        sm = 0.5 * obj + 0.5 * reg;
        sm = sm > 0 ? sm : 0;  // python max(0, nan) == 0 as well
    }
    o[3] = sm;
    // adaptive E / F
    const double bcnt = a[A_ADP_B], bgcnt = a[A_ADP_BG];
    o[4] = em_from_counts(bgcnt, bcnt - bgcnt, gt_fg, size);
    if (bgcnt == 0) o[5] = 0;
    else {
        const double pre = bgcnt / bcnt, rec = bgcnt / gt_fg;
        o[5] = 1.3 * pre * rec / (0.3 * pre + rec);
    }
    // weighted F
    if (gt_fg == 0) o[6] = 0;
    else {
        const double tpw = a[A_W_SUM_G] - a[A_W_EW_G], fpw = a[A_W_EW_BG];
        const double R = 1 - a[A_W_EW_G] / a[A_W_SUM_G];
        const double P = tpw / (tpw + fpw + M_EPS);
        o[6] = 2 * R * P / (R + P + M_EPS);
    }
}

size_t align256m(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

size_t cod_metrics_workspace_bytes(int B, int h, int w) {
    const size_t n = (size_t)B * h * w;
    return align256m((size_t)B * ACC_STRIDE * 8) + align256m((size_t)B * 512 * 4) + align256m(n * 4) * 2 +
           align256m(n * 8) * 2 + 1024;
}

int cod_metrics(const float* gt, const float* pred, int B, int h, int w, double* out, void* workspace, size_t ws_bytes,
                cudaStream_t stream) {
    UCOD_REQUIRE(gt && pred && out && workspace, "cod_metrics: null argument");
    UCOD_REQUIRE(B > 0 && h > 0 && w > 0 && (long long)h * w < (1ll << 30), "cod_metrics: bad geometry");
    UCOD_REQUIRE(ws_bytes >= cod_metrics_workspace_bytes(B, h, w), "cod_metrics: workspace too small");
    static bool gauss_ready = false;
    if (!gauss_ready) {  // matlab_style_gauss2D((7,7), sigma=5) (metric.py:512-525)
        double k[49], s = 0;
        for (int y = -3; y <= 3; ++y)
            for (int x = -3; x <= 3; ++x) s += (k[(y + 3) * 7 + x + 3] = exp(-(double)(x * x + y * y) / 50.0));
        for (int i = 0; i < 49; ++i) k[i] /= s;
        UCOD_CHECK_CUDA(cudaMemcpyToSymbol(c_gauss7, k, sizeof(k)));
        gauss_ready = true;
    }
    const size_t n = (size_t)B * h * w;
    uint8_t* p = static_cast<uint8_t*>(workspace);
    double* acc = reinterpret_cast<double*>(p);
    p += align256m((size_t)B * ACC_STRIDE * 8);
    unsigned int* hist = reinterpret_cast<unsigned int*>(p);
    p += align256m((size_t)B * 512 * 4);
    int* vdist = reinterpret_cast<int*>(p);
    p += align256m(n * 4);
    int* feat = reinterpret_cast<int*>(p);
    p += align256m(n * 4);
    double* dst = reinterpret_cast<double*>(p);
    p += align256m(n * 8);
    double* et = reinterpret_cast<double*>(p);
    UCOD_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)B * ACC_STRIDE * 8, stream));
    UCOD_CHECK_CUDA(cudaMemsetAsync(hist, 0, (size_t)B * 512 * 4, stream));
    const int npx = h * w;
    const int chunks = ceil_div(npx, 256 * 16) < 64 ? ceil_div(npx, 256 * 16) : 64;
    ProfScope ps(KC_OTHER, stream, (double)n * 8 * 6);
    metrics_minmax_kernel<<<B, 1024, 0, stream>>>(gt, pred, npx, acc);
    metrics_stats_kernel<<<dim3(chunks, B), 256, 0, stream>>>(gt, pred, h, w, acc, hist);
    metrics_derive_kernel<<<ceil_div(B, 64), 64, 0, stream>>>(acc, B, h, w);
    metrics_region_kernel<<<dim3(chunks, B), 256, 0, stream>>>(gt, pred, h, w, acc);
    metrics_edt_cols_kernel<<<dim3(ceil_div(w, 64), B), 64, 0, stream>>>(gt, acc, h, w, vdist);
    metrics_edt_rows_kernel<<<dim3(ceil_div(w, 128), h, B), 128, 0, stream>>>(vdist, h, w, dst, feat);
    metrics_et_kernel<<<dim3(ceil_div(npx, 256), B), 256, 0, stream>>>(gt, pred, acc, feat, npx, et);
    metrics_wfm_kernel<<<dim3(chunks, B), 256, 0, stream>>>(gt, pred, et, dst, h, w, acc);
    metrics_finalize_kernel<<<B, 256, 0, stream>>>(acc, hist, h, w, out);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
