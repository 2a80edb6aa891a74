// Look-Twice device stages (engine/runner/loop_UCOD_DPL.py:326-417).
//
//  1. lt_boxes: 8-connected component labelling of the binarised 518^2 / 296^2 mask (union-find in global memory,
//     atomicMin linking), per-component area / bounding box / OpenCV ordering key, then the reference's box logic
//     (`process_preds` :366-384, `expand_bbox` :399-417) in IEEE fp64 with explicit round-to-nearest ops so that
//     `int()` truncations match CPython bit for bit.  OpenCV numbers components by their first 2x2 block in
//     block-raster order; that key is reproduced so ties in the final stable sort fall the same way.
//  2. roi_crop_resize: PIL `crop` + torchvision `Resize` (= Pillow ImagingResample, antialiased triangle filter):
//     fp64 coefficients -> 22-bit fixed point, horizontal pass then vertical pass, uint8 rounding after each.
//  3. paste_bicubic: `ToPILImage` + `Image.resize` (Pillow default BICUBIC, a=-0.5) of the second-pass 37^2 mask and
//     `paste` into the full-size mask, boxes applied in the reference's order.
// All three are HBM-/latency-bound integer work: coalesced row-major passes, warp-aggregated atomics.
#include "looktwice.cuh"

#include "prof.cuh"

namespace ucod {

namespace {

// ------------------------------------------------------------------------------------------------
// union-find connected components
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* L, int i) {
    const volatile int* V = L;
    while (true) {
        const int p = V[i];
        if (p == i) return i;
        i = p;
    }
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&L[a], b);  // link larger root under the smaller one
        if (old == a) return;
        a = old;
    }
}

__global__ void ccl_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, int n_img, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    L[i] = mask[i] ? (int)(i % n_img) : -1;
}

__global__ void ccl_merge_kernel(int* __restrict__ L, int H, int W, int B) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    if (x >= W) return;
    int* Lb = L + (size_t)b * H * W;
    const int i = y * W + x;
    if (Lb[i] < 0) return;
    if (x > 0 && Lb[i - 1] >= 0) uf_union(Lb, i, i - 1);
    if (y > 0) {
        const int u = i - W;
        if (Lb[u] >= 0) uf_union(Lb, i, u);
        if (x > 0 && Lb[u - 1] >= 0) uf_union(Lb, i, u - 1);
        if (x + 1 < W && Lb[u + 1] >= 0) uf_union(Lb, i, u + 1);
    }
}

// final root per pixel + per-root area (warp-aggregated atomics)
__global__ void ccl_flatten_area_kernel(int* __restrict__ L, int* __restrict__ area, int n_img, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int root = -1;
    size_t base = 0;
    if (i < total) {
        base = (i / n_img) * (size_t)n_img;
        const int l = L[i];
        if (l >= 0) {
            root = uf_find(L + base, (int)(i - base));
            L[i] = root;
        }
    }
    // key must distinguish images: a warp may straddle an image boundary
    const long long key = root < 0 ? -1ll : (long long)(base + root);
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    if (root >= 0) {
        const int leader = __ffs(grp) - 1;
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(&area[base + root], __popc(grp));
    }
}

constexpr int LT_MAXBIG = 128;  // components with area fraction > 0.01: at most 99

struct ImgSummary {
    int n_comp;
    int max_area;
    int n_big;
    int pad;
};

// roots: count components, track the largest, register the "big" ones (p > 0.01) in a compact table
__global__ void ccl_roots_kernel(const int* __restrict__ L, int* __restrict__ area, ImgSummary* __restrict__ summ,
                                 int* __restrict__ big_root, int* __restrict__ big_area, int n_img, size_t total,
                                 double inv_gate /* p > 0.01 test uses area / n_img in fp64 */) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = (int)(i / n_img);
    const int li = (int)(i - (size_t)b * n_img);
    if (L[i] != li) return;
    const int a = area[i];
    atomicAdd(&summ[b].n_comp, 1);
    atomicMax(&summ[b].max_area, a);
    const double p = __ddiv_rn((double)a, (double)n_img);
    if (p > inv_gate) {
        const int slot = atomicAdd(&summ[b].n_big, 1);
        if (slot < LT_MAXBIG) {
            big_root[b * LT_MAXBIG + slot] = li;
            big_area[b * LT_MAXBIG + slot] = a;
            area[i] = -(slot + 1);  // pixels find their slot through their root
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Run-based connected components in shared memory: ONE CTA labels one whole mask without ever materialising a label
// image.  The u8 mask (the only HBM-sized read: H*W bytes) is packed to one bit per pixel with warp ballots; rows are
// cut into runs of foreground pixels; runs of adjacent rows that touch (8-connectivity: overlap after growing by one
// pixel) are united with a union-find over RUN indices (atomicMin linking in shared memory); component areas are
// accumulated per root, the big components (area fraction > 0.01) get bounding box and OpenCV first-block key from
// their runs.  Outputs are the per-image summary / big-component tables the box kernel (`lt_boxes_kernel`) consumes.
// Shared memory: bit image (4*H*ceil(W/32) B, later reused for the per-root accumulators) + 8 B per run.
// Capacity: `rmax` runs and as many components as the bit image has words; a mask beyond that sets flag 4 in the
// summary (nbox = -3) and the host re-runs that batch through the global-memory labeller.  Masks that come from a
// bilinearly up-sampled fs x fs logit map have at most fs/2+1 runs per row and (fs/2)^2 components — 18 130 runs and
// 1 156 components for 68 -> 518 — so the Look-Twice pipeline never overflows.
// ------------------------------------------------------------------------------------------------
struct BigStats {
    int x0, x1, y0, y1, key;
};

__device__ __forceinline__ int rl_find(const int* parent, int i) {
    const volatile int* V = parent;
    while (true) {
        const int p = V[i];
        if (p == i) return i;
        i = p;
    }
}
__device__ __forceinline__ void rl_union(int* parent, int a, int b) {
    while (true) {
        a = rl_find(parent, a);
        b = rl_find(parent, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(1024, 1)
    ccl_runs_kernel(const uint8_t* __restrict__ mask, int H, int W, int rmax, ImgSummary* __restrict__ summ,
                    int* __restrict__ big_area, BigStats* __restrict__ st, int* __restrict__ labels_out,
                    double big_gate) {
    extern __shared__ __align__(16) uint8_t ccl_smem[];
    const int WW = (W + 31) >> 5;
    uint32_t* bits = reinterpret_cast<uint32_t*>(ccl_smem);              // [H * WW]; later: acc[] per root slot
    int* acc = reinterpret_cast<int*>(ccl_smem);
    int* row_off = reinterpret_cast<int*>(bits + (size_t)H * WW);       // [H + 1]
    int* parent = row_off + (H + 1);                                     // [rmax]
    uint16_t* rs = reinterpret_cast<uint16_t*>(parent + rmax);           // [rmax] run start x
    uint16_t* re = rs + rmax;                                            // [rmax] run end x (inclusive)
    __shared__ int s_scan[33];
    __shared__ int s_total, s_roots, s_max_area, s_nbig, s_flags;
    __shared__ BigStats s_big[LT_MAXBIG];
    __shared__ int s_big_area[LT_MAXBIG];

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int stage_bytes = ((W + 15 + 16) >> 4) << 4;  // one staged row per warp (16-byte vectors incl. head / tail)
    if (tid == 0) s_max_area = 0, s_nbig = 0, s_flags = 0, s_roots = 0;
    for (int i = tid; i < LT_MAXBIG; i += blockDim.x) {
        s_big[i].x0 = W, s_big[i].x1 = -1, s_big[i].y0 = H, s_big[i].y1 = -1, s_big[i].key = 0x7fffffff;
        s_big_area[i] = 0;
    }

    // ---- 1. bit image: one warp per row.  The row is staged in shared memory with 16-byte loads from the
    //         aligned-down address (all of a row's loads are in flight together: one memory round trip per row instead
    //         of one per 32 pixels), then one ballot per 32 pixels turns it into words ----
    {
        uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(re + rmax) + 15) & ~uintptr_t(15)) +
                         (size_t)warp * stage_bytes;
        const size_t total_bytes = (size_t)gridDim.x * H * W;
        for (int y = warp; y < H; y += nwarps) {
            const size_t off = ((size_t)b * H + y) * W;       // byte offset of the row in the batch tensor
            const uint8_t* row = mask + off;
            const int head = (int)(reinterpret_cast<uintptr_t>(row) & 15);
            const int nvec = (head + W + 15) >> 4;
            for (int v = lane; v < nvec; v += 32) {
                // bytes [16 v - head, 16 v - head + 16) of the row; stay inside the tensor at both ends
                const long long lo = (long long)off + 16ll * v - head;
                if (lo >= 0 && (size_t)lo + 16 <= total_bytes) {
                    reinterpret_cast<uint4*>(stage)[v] = __ldg(reinterpret_cast<const uint4*>(mask + lo));
                } else {
                    for (int k = 0; k < 16; ++k) {
                        const long long a = lo + k;
                        stage[16 * v + k] = (a >= 0 && (size_t)a < total_bytes) ? mask[a] : 0;
                    }
                }
            }
            __syncwarp();
            for (int w0 = 0; w0 < WW; ++w0) {
                const int x = w0 * 32 + lane;
                const uint32_t word = __ballot_sync(0xffffffffu, x < W && stage[head + x] != 0);
                if (lane == 0) bits[y * WW + w0] = word;
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- 2. runs per row (count, exclusive scan over rows, extraction) ----
    int my_cnt = 0;
    if (tid < H) {
        uint32_t carry = 0;
        for (int w0 = 0; w0 < WW; ++w0) {
            const uint32_t v = bits[tid * WW + w0];
            my_cnt += __popc(v & ~((v << 1) | carry));
            carry = v >> 31;
        }
    }
    {   // block-wide exclusive scan of my_cnt over tid (H <= blockDim.x)
        int incl = my_cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = lane < nwarps ? s_scan[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            s_scan[lane] = v;  // inclusive totals per warp
            if (lane == 31) s_total = v;
        }
        __syncthreads();
        const int base = (warp > 0 ? s_scan[warp - 1] : 0) + incl - my_cnt;
        if (tid < H) row_off[tid] = base;
        if (tid == 0) row_off[H] = s_total;
    }
    __syncthreads();
    const int n_runs = s_total;
    if (n_runs > rmax) {  // out of run slots: flagged, nothing else is written for this image
        if (tid == 0) {
            summ[b].n_comp = 0, summ[b].max_area = 0, summ[b].n_big = 0, summ[b].pad = 4;
        }
        return;
    }
    if (tid < H) {
        int o = row_off[tid];
        bool in_run = false;
        for (int w0 = 0; w0 < WW; ++w0) {
            uint32_t v = bits[tid * WW + w0];
            int pos = 0;  // bits below `pos` are consumed
            while (pos < 32) {
                const uint32_t rest = pos ? (v >> pos) : v;
                if (!in_run) {
                    if (rest == 0) break;
                    const int s = __ffs(rest) - 1;
                    pos += s;
                    rs[o] = (uint16_t)(w0 * 32 + pos);
                    in_run = true;
                } else {
                    const uint32_t inv = ~rest & (pos ? (0xffffffffu >> pos) : 0xffffffffu);
                    if (inv == 0) {  // run continues to the end of the word
                        pos = 32;
                        break;
                    }
                    const int e = __ffs(inv) - 1;
                    pos += e;
                    re[o++] = (uint16_t)(w0 * 32 + pos - 1);
                    in_run = false;
                }
            }
        }
        if (in_run) re[o++] = (uint16_t)(W - 1);  // bits past W are zero, so this only happens at the row end
    }
    for (int i = tid; i < n_runs; i += blockDim.x) parent[i] = i;
    __syncthreads();

    // ---- 3. unite touching runs of adjacent rows (two-pointer walk per row pair) ----
    if (tid >= 1 && tid < H) {
        int i = row_off[tid - 1], j = row_off[tid];
        const int ie = row_off[tid], je = row_off[tid + 1];
        while (i < ie && j < je) {
            const int as = rs[i], ae = re[i], bs = rs[j], be = re[j];
            if (as <= be + 1 && bs <= ae + 1) rl_union(parent, i, j);
            if (ae < be) ++i; else ++j;
        }
    }
    __syncthreads();
    // ---- 4. flatten, count roots, give every root an accumulator slot ----
    for (int i = tid; i < n_runs; i += blockDim.x) parent[i] = rl_find(parent, i);
    __syncthreads();
    const int acc_cap = H * WW;
    for (int i0 = 0; i0 < n_runs; i0 += blockDim.x) {  // roots get slots in run order (chunked block scan)
        const int i = i0 + tid;
        const int is_root = (i < n_runs && parent[i] == i) ? 1 : 0;
        int incl = is_root;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = lane < nwarps ? s_scan[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            s_scan[lane] = v;
        }
        __syncthreads();
        const int slot = s_roots + (warp > 0 ? s_scan[warp - 1] : 0) + incl - is_root;
        if (is_root && slot < acc_cap) parent[i] = -1 - slot;
        __syncthreads();
        if (tid == 0) s_roots += s_scan[31];
        __syncthreads();
    }
    const int n_roots = s_roots;
    if (n_roots > acc_cap) {
        if (tid == 0) summ[b].n_comp = 0, summ[b].max_area = 0, summ[b].n_big = 0, summ[b].pad = 4;
        return;
    }
    for (int i = tid; i < n_roots; i += blockDim.x) acc[i] = 0;  // (the bit image is no longer needed)
    __syncthreads();
    auto slot_of = [&](int i) {
        const int p = parent[i];
        return p < 0 ? -1 - p : -1 - parent[p];
    };
    // ---- 5. areas ----
    for (int i = tid; i < n_runs; i += blockDim.x) atomicAdd(&acc[slot_of(i)], (int)re[i] - (int)rs[i] + 1);
    __syncthreads();
    // ---- 6. largest component, big components ----
    const double hw = (double)((long long)H * W);
    for (int k = tid; k < n_roots; k += blockDim.x) {
        const int a = acc[k];
        atomicMax(&s_max_area, a);
        if (__ddiv_rn((double)a, hw) > big_gate) {
            const int bslot = atomicAdd(&s_nbig, 1);
            if (bslot < LT_MAXBIG) {
                s_big_area[bslot] = a;
                acc[k] = -(bslot + 1);
            }
        }
    }
    __syncthreads();
    // ---- 7. bounding box / OpenCV first-block key of the big components, from their runs (thread per row) ----
    if (tid < H && s_nbig > 0) {
        const int half_w = (W + 1) >> 1;
        for (int i = row_off[tid]; i < row_off[tid + 1]; ++i) {
            const int a = acc[slot_of(i)];
            if (a >= 0) continue;
            BigStats* g = &s_big[-a - 1];
            atomicMin(&g->x0, (int)rs[i]);
            atomicMax(&g->x1, (int)re[i]);
            atomicMin(&g->y0, tid);
            atomicMax(&g->y1, tid);
            atomicMin(&g->key, (tid >> 1) * half_w + ((int)rs[i] >> 1));
        }
    }
    __syncthreads();
    if (tid == 0) {
        summ[b].n_comp = n_roots, summ[b].max_area = s_max_area, summ[b].n_big = s_nbig, summ[b].pad = 0;
    }
    for (int i = tid; i < LT_MAXBIG; i += blockDim.x) {
        st[b * LT_MAXBIG + i] = s_big[i];
        big_area[b * LT_MAXBIG + i] = s_big_area[i];
    }
    // ---- 8. optional label image: every pixel gets the smallest raster index of its component, background -1 ----
    if (labels_out != nullptr) {
        int* L = labels_out + (size_t)b * H * W;
        for (int i = tid; i < H * W; i += blockDim.x) L[i] = -1;
        __syncthreads();
        if (tid < H) {
            for (int i = row_off[tid]; i < row_off[tid + 1]; ++i) {
                const int p = parent[i];
                const int root = p < 0 ? i : p;
                int lo = 0, hi = H;  // row of the root run: last y with row_off[y] <= root
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (row_off[mid] <= root) lo = mid; else hi = mid;
                }
                const int lab = lo * W + (int)rs[root];
                for (int x = rs[i]; x <= re[i]; ++x) L[tid * W + x] = lab;
            }
        }
    }
}

static size_t ccl_runs_smem_bytes(int H, int W, int rmax) {
    const size_t ww = (size_t)((W + 31) >> 5);
    const size_t stage = (size_t)(((W + 15 + 16) >> 4) << 4) * 32;  // per-warp row staging
    return (size_t)H * ww * 4 + (size_t)(H + 1) * 4 + (size_t)rmax * 8 + stage + 16;
}
// run capacity that fits the 227 KB of one CTA (0: this geometry cannot use the shared-memory labeller)
static int ccl_runs_capacity(int H, int W) {
    if (H > 1024 || W > 65535) return 0;
    const size_t fixed = ccl_runs_smem_bytes(H, W, 0) + 6 * 1024;  // + the kernel's static shared arrays
    const size_t budget = 227 * 1024;
    if (fixed + 8 * 1024 > budget) return 0;
    return (int)((budget - fixed) / 8) & ~7;
}

__global__ void ccl_bbox_kernel(const int* __restrict__ L, const int* __restrict__ area, BigStats* __restrict__ st,
                                int H, int W, int B) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    int slot = -1;
    if (x < W) {
        const size_t base = (size_t)b * H * W;
        const int r = L[base + (size_t)y * W + x];
        if (r >= 0) {
            const int a = area[base + r];
            if (a < 0) slot = -a - 1;
        }
    }
    const unsigned grp = __match_any_sync(0xffffffffu, slot);
    if (slot >= 0) {
        // all lanes of a warp share y and b: reduce x over the group, one set of atomics per group
        const int xmin = __reduce_min_sync(grp, x);
        const int xmax = __reduce_max_sync(grp, x);
        if ((int)(threadIdx.x & 31) == __ffs(grp) - 1) {
            BigStats* s = st + b * LT_MAXBIG + slot;
            atomicMin(&s->x0, xmin);
            atomicMax(&s->x1, xmax);
            atomicMin(&s->y0, y);
            atomicMax(&s->y1, y);
            atomicMin(&s->key, (y >> 1) * ((W + 1) >> 1) + (xmin >> 1));
        }
    }
}

__global__ void lt_init_stats_kernel(BigStats* st, int n, int H, int W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st[i].x0 = W, st[i].x1 = -1, st[i].y0 = H, st[i].y1 = -1, st[i].key = 0x7fffffff;
}

// One thread per image: process_preds' branchy tail + expand_bbox in exact fp64.
__global__ void lt_boxes_kernel(const ImgSummary* __restrict__ summ, const int* __restrict__ big_area,
                                const BigStats* __restrict__ st, int* __restrict__ boxes, int* __restrict__ nbox,
                                int* __restrict__ status, int B, int H, int W, double look_twice_th, int dynamic,
                                double const_scale) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int* out = boxes + (size_t)b * LT_MAXBIG * 4;
    const ImgSummary s = summ[b];
    status[b] = s.pad & 4;
    if (s.pad & 4) {  // the shared-memory labeller ran out of run / component slots for this mask
        nbox[b] = -3;
        return;
    }
    if (s.n_comp == 0) {  // loop_UCOD_DPL.py:369-370
        out[0] = 129, out[1] = 129, out[2] = 259, out[3] = 259;
        nbox[b] = 1;
        return;
    }
    const double hw = (double)((long long)H * W);
    const double p_max = __ddiv_rn((double)s.max_area, hw);
    if (!(p_max < look_twice_th)) {  // :372 else-branch -> None
        nbox[b] = -1;
        return;
    }
    int n = s.n_big < LT_MAXBIG ? s.n_big : LT_MAXBIG;
    if (s.n_big > LT_MAXBIG) status[b] |= 2;
    // order slots by OpenCV label order (first-block key), insertion sort on a local index array
    int order[LT_MAXBIG];
    for (int i = 0; i < n; ++i) {
        int j = i;
        const int k = st[b * LT_MAXBIG + i].key;
        while (j > 0 && st[b * LT_MAXBIG + order[j - 1]].key > k) {
            order[j] = order[j - 1];
            --j;
        }
        order[j] = i;
    }
    int bx[LT_MAXBIG][4];
    for (int t = 0; t < n; ++t) {
        const int slot = order[t];
        const BigStats g = st[b * LT_MAXBIG + slot];
        const int x = g.x0, y = g.y0, w = g.x1 - g.x0 + 1, h = g.y1 - g.y0 + 1;
        double scale = const_scale;
        if (dynamic) {
            const double fr = __ddiv_rn((double)big_area[b * LT_MAXBIG + slot], (double)(h * w));
            const double br = __ddiv_rn((double)(h * y), hw);  // (sic) h*y, loop_UCOD_DPL.py:404
            const double arg = __dadd_rn(__dsub_rn(1.0, __ddiv_rn(br, fr)), 1.0);
            if (arg < 0.0) {  // math.sqrt raises ValueError in the reference
                status[b] |= 1;
                nbox[b] = -2;
                return;
            }
            scale = __dsqrt_rn(arg);
        }
        const double new_w = __dmul_rn((double)w, scale);
        const double new_h = __dmul_rn((double)h, scale);
        double new_x = __dsub_rn((double)x, __ddiv_rn(__dsub_rn(new_w, (double)w), 2.0));
        double new_y = __dsub_rn((double)y, __ddiv_rn(__dsub_rn(new_h, (double)h), 2.0));
        new_x = new_x > 0.0 ? new_x : 0.0;  // max(0, new_x)
        if (__dadd_rn(new_x, new_w) > (double)H) new_x = __dsub_rn((double)H, new_w);  // img_width := h (:379)
        new_y = new_y > 0.0 ? new_y : 0.0;
        if (__dadd_rn(new_y, new_h) > (double)W) new_y = __dsub_rn((double)W, new_h);  // img_height := w
        bx[t][0] = (int)new_x, bx[t][1] = (int)new_y, bx[t][2] = (int)new_w, bx[t][3] = (int)new_h;
    }
    // stable sort by -w*h
    int ord2[LT_MAXBIG];
    for (int i = 0; i < n; ++i) {
        int j = i;
        const long long k = -(long long)bx[i][2] * bx[i][3];
        while (j > 0 && -(long long)bx[ord2[j - 1]][2] * bx[ord2[j - 1]][3] > k) {
            ord2[j] = ord2[j - 1];
            --j;
        }
        ord2[j] = i;
    }
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 4; ++c) out[i * 4 + c] = bx[ord2[i]][c];
    nbox[b] = n;
}

// ------------------------------------------------------------------------------------------------
// Pillow ImagingResample (8 bpc): coefficient tables
// ------------------------------------------------------------------------------------------------
constexpr int RS_KMAX = 40;          // default tap capacity per output sample (down-scale up to ~19x for bilinear)
constexpr int RS_PRECISION_BITS = 22;

__device__ __forceinline__ double filt_bilinear(double x) {
    if (x < 0.0) x = -x;
    if (x < 1.0) return __dsub_rn(1.0, x);
    return 0.0;
}
__device__ __forceinline__ double filt_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) {
        // ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        double t = __dsub_rn(__dmul_rn(a + 2.0, x), a + 3.0);
        t = __dmul_rn(__dmul_rn(t, x), x);
        return __dadd_rn(t, 1.0);
    }
    if (x < 2.0) {
        // (((x - 5) * x + 8) * x - 4) * a
        double t = __dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0);
        t = __dsub_rn(__dmul_rn(t, x), 4.0);
        return __dmul_rn(t, a);
    }
    return 0.0;
}

// job geometry for one axis: in_size source samples -> out_size samples (box = whole source)
// bounds[(job*2+axis)*out_cap + xx] = {xmin, count} ; coef[((job*2+axis)*kmax + k)*out_cap + xx]
// (tap-major, so that neighbouring output samples read neighbouring coefficients: coalesced in the horizontal passes)
// kmax = tap capacity per output sample (host: from the largest possible down-scale); a job needing more sets err bit 0
// and gets count 0 (its output is then the rounding constant only — the caller raises on the flag).
// The filter is evaluated twice (sum, then normalised taps) instead of keeping the taps in a local array.
__global__ void resample_coeffs_kernel(const int* __restrict__ in_sizes /*[njobs,2] (w,h)*/,
                                       const int* __restrict__ out_sizes /*[njobs,2] (w,h)*/, int njobs, int out_cap,
                                       int kmax, int bicubic, int2* __restrict__ bounds, int* __restrict__ coef,
                                       int* __restrict__ err, const int* __restrict__ njobs_dev) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int axis = blockIdx.y, job = blockIdx.z;
    if (njobs_dev != nullptr && job >= __ldg(njobs_dev)) return;
    const int in_size = in_sizes[job * 2 + axis], out_size = out_sizes[job * 2 + axis];
    if (xx >= out_size || xx >= out_cap || in_size <= 0) return;
    const double fsupport = bicubic ? 2.0 : 1.0;
    const double scale = __ddiv_rn((double)(float)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = __dmul_rn(fsupport, filterscale);
    const int ksize = (int)ceil(support) * 2 + 1;
    const size_t o = ((size_t)job * 2 + axis) * out_cap + xx;
    if (ksize > kmax) {
        atomicOr(err, 1);
        bounds[o] = make_int2(0, 0);
        return;
    }
    const double center = __dadd_rn(0.0, __dmul_rn((double)xx + 0.5, scale));
    const double ss = __ddiv_rn(1.0, filterscale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
        const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        ww = __dadd_rn(ww, bicubic ? filt_bicubic(arg) : filt_bilinear(arg));
    }
    bounds[o] = make_int2(xmin, n);
    int* kk = coef + ((size_t)job * 2 + axis) * kmax * out_cap + xx;  // stride out_cap between taps
    for (int x = 0; x < n; ++x) {
        const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        double k = bicubic ? filt_bicubic(arg) : filt_bilinear(arg);
        if (ww != 0.0) k = __ddiv_rn(k, ww);
        const double f = __dmul_rn(k, (double)(1 << RS_PRECISION_BITS));
        kk[(size_t)x * out_cap] = k < 0 ? (int)__dadd_rn(-0.5, f) : (int)__dadd_rn(0.5, f);
    }
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ---- crop + horizontal pass: tmp[job][c][r][xx], r over source rows y_first..y_last-1 of the crop ----
__global__ void crop_hpass_kernel(const uint8_t* __restrict__ images, int H0, int W0, long long img_stride,
                                  long long ch_stride, long long row_stride, long long px_stride,
                                  const int* __restrict__ jobs /*[n,5] img,x,y,w,h*/, const int2* __restrict__ bounds,
                                  const int* __restrict__ coef, uint8_t* __restrict__ tmp, int out_cap, int out_w,
                                  int out_h, int tmp_rows, int kmax, const int* __restrict__ njobs_dev) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int job = blockIdx.z;
    if (njobs_dev != nullptr && job >= __ldg(njobs_dev)) return;
    const int c = blockIdx.y % 3, r = blockIdx.y / 3;
    const int* jb = jobs + job * 5;
    const int cw = jb[3], chh = jb[4];
    if (cw <= 0 || chh <= 0 || xx >= out_w) return;
    const size_t vb = ((size_t)job * 2 + 1) * out_cap;
    const int y_first = bounds[vb].x;
    const int y_last = bounds[vb + out_h - 1].x + bounds[vb + out_h - 1].y;
    // r walks the image rows the crop touches; rows of the crop outside the image are zero after the horizontal pass
    // (PIL pads a crop with 0), so they are neither computed nor stored: the scratch is indexed by the image row.
    const int sy0 = jb[2] + y_first;
    const int sy = (sy0 > 0 ? sy0 : 0) + r;  // source row in the original image
    if (sy >= jb[2] + y_last || sy >= H0 || r >= tmp_rows) return;
    const size_t hb = ((size_t)job * 2 + 0) * out_cap + xx;
    const int2 bd = bounds[hb];
    const int* kk = coef + ((size_t)job * 2 + 0) * kmax * out_cap + xx;
    int acc = 1 << (RS_PRECISION_BITS - 1);
    const uint8_t* row = images + (size_t)jb[0] * img_stride + (size_t)c * ch_stride + (size_t)sy * row_stride;
    for (int k = 0; k < bd.y; ++k) {
        const int sx = jb[1] + bd.x + k;
        const int px = (sx >= 0 && sx < W0) ? row[(size_t)sx * px_stride] : 0;
        acc += px * __ldg(kk + (size_t)k * out_cap);
    }
    tmp[(((size_t)job * 3 + c) * tmp_rows + r) * out_w + xx] = clip8(acc);
}

// ---- vertical pass: out[job][c][yy][xx] (planar u8) ----
__global__ void crop_vpass_kernel(const uint8_t* __restrict__ tmp, const int* __restrict__ jobs,
                                  const int2* __restrict__ bounds, const int* __restrict__ coef,
                                  uint8_t* __restrict__ out, int out_cap, int out_w, int out_h, int tmp_rows, int kmax,
                                  int H0, const int* __restrict__ njobs_dev) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int job = blockIdx.z;
    if (njobs_dev != nullptr && job >= __ldg(njobs_dev)) return;
    const int c = blockIdx.y % 3, yy = blockIdx.y / 3;
    const int* jb = jobs + job * 5;
    if (xx >= out_w) return;
    uint8_t* dst = out + (((size_t)job * 3 + c) * out_h + yy) * out_w + xx;
    if (jb[3] <= 0 || jb[4] <= 0) {
        *dst = 0;
        return;
    }
    const size_t vb = ((size_t)job * 2 + 1) * out_cap;
    const int y_first = bounds[vb].x;
    const int2 bd = bounds[vb + yy];
    const int* kk = coef + ((size_t)job * 2 + 1) * kmax * out_cap + yy;  // same address for the whole block
    const uint8_t* src = tmp + ((size_t)job * 3 + c) * tmp_rows * out_w + xx;
    const int sy0 = jb[2] + y_first;
    const int base = sy0 > 0 ? sy0 : 0;  // image row stored at scratch row 0 (see crop_hpass_kernel)
    int acc = 1 << (RS_PRECISION_BITS - 1);
    for (int k = 0; k < bd.y; ++k) {
        const int sy = jb[2] + bd.x + k;
        if (sy >= 0 && sy < H0) acc += (int)src[(size_t)(sy - base) * out_w] * __ldg(kk + (size_t)k * out_cap);
    }
    *dst = clip8(acc);
}

// ---- crop + both resampling passes in one kernel.  CTA = (tile of CR_TY output rows, channel, job): the source rows
// the tile's vertical taps touch are resampled horizontally into shared memory (u8, rounded like Pillow's first pass),
// then the vertical pass reads them from there — the intermediate image never goes to HBM and the grid has one CTA per
// 16 output rows instead of one per source row (round 1: 2 x 500 k mostly idle CTAs per chunk, 0.67 ms each).
// When a tile needs more source rows than fit (large down-scales), it is processed in several vertical sub-tiles.
constexpr int CR_TY = 16;        // output rows per CTA
constexpr int CR_ROWS = 72;      // staged source rows (shared memory: CR_ROWS * out_w bytes)

__global__ void __launch_bounds__(256)
    crop_resize_fused_kernel(const uint8_t* __restrict__ images, int H0, int W0, long long img_stride,
                             long long ch_stride, long long row_stride, long long px_stride,
                             const int* __restrict__ jobs /*[n,5] img,x,y,w,h*/, const int2* __restrict__ bounds,
                             const int* __restrict__ coef, uint8_t* __restrict__ out, int out_cap, int out_w, int out_h,
                             int kmax, const int* __restrict__ njobs_dev) {
    extern __shared__ __align__(16) uint8_t cr_rows[];  // [CR_ROWS][pitch]
    const int pitch = (out_w + 3) & ~3;
    const int job = blockIdx.z;
    if (njobs_dev != nullptr && job >= __ldg(njobs_dev)) return;
    const int c = blockIdx.y;
    const int* jb = jobs + job * 5;
    const int yy0 = blockIdx.x * CR_TY;
    const int yy1 = min(out_h, yy0 + CR_TY);
    uint8_t* dst = out + ((size_t)job * 3 + c) * out_h * out_w;
    if (jb[3] <= 0 || jb[4] <= 0) {
        for (int i = threadIdx.x; i < (yy1 - yy0) * out_w; i += blockDim.x) dst[(size_t)yy0 * out_w + i] = 0;
        return;
    }
    const size_t vb = ((size_t)job * 2 + 1) * out_cap, hb = ((size_t)job * 2 + 0) * out_cap;
    const int* kv = coef + ((size_t)job * 2 + 1) * kmax * out_cap;
    const int* kh = coef + ((size_t)job * 2 + 0) * kmax * out_cap;
    const uint8_t* plane = images + (size_t)jb[0] * img_stride + (size_t)c * ch_stride;
    int ya = yy0;
    while (ya < yy1) {
        // sub-tile [ya, yb): as many output rows as their source-row span fits the staging buffer (at least one)
        const int lo = bounds[vb + ya].x;
        int yb = ya + 1;
        while (yb < yy1 && bounds[vb + yb].x + bounds[vb + yb].y - lo <= CR_ROWS) ++yb;
        const int hi = bounds[vb + yb - 1].x + bounds[vb + yb - 1].y;   // crop rows [lo, hi)
        const int span = min(hi - lo, CR_ROWS);  // (one output row never needs more than kmax <= CR_ROWS rows: checked by the host)
        __syncthreads();
        // horizontal pass of the staged rows; (row, column) walked incrementally: no integer division per pixel, 32-bit
        // offsets inside the image plane (64-bit index arithmetic tripled the instruction count of the tap loop)
        {
            const int rs = (int)row_stride, ps = (int)px_stride;
            int r = 0, xx = threadIdx.x;
            while (xx >= out_w) xx -= out_w, ++r;
            while (r < span) {
                const int sy = jb[2] + lo + r;
                int acc = 1 << (RS_PRECISION_BITS - 1);
                if (sy >= 0 && sy < H0) {
                    const int2 bd = bounds[hb + xx];
                    const int* kk = kh + xx;
                    const int sx0 = jb[1] + bd.x;
                    if (sx0 >= 0 && sx0 + bd.y <= W0) {  // all taps inside the image: no per-tap test
                        const uint8_t* px = plane + (sy * rs + sx0 * ps);
                        int ko = 0, po = 0;
                        for (int k = 0; k < bd.y; ++k, ko += out_cap, po += ps) acc += (int)px[po] * __ldg(kk + ko);
                    } else {
                        const uint8_t* row = plane + sy * rs;
                        int ko = 0;
                        for (int k = 0; k < bd.y; ++k, ko += out_cap) {
                            const int sx = sx0 + k;
                            const int px = (sx >= 0 && sx < W0) ? row[sx * ps] : 0;
                            acc += px * __ldg(kk + ko);
                        }
                    }
                }
                cr_rows[r * pitch + xx] = clip8(acc);
                xx += blockDim.x;
                while (xx >= out_w) xx -= out_w, ++r;
            }
        }
        __syncthreads();
        // vertical pass: one thread = 4 consecutive pixels of one output row (32-bit shared-memory reads of the staged
        // rows, whose pitch is a multiple of 4; the row's taps are the same for the whole row)
        {
            const int groups = (out_w + 3) >> 2;
            int yy = ya, g = threadIdx.x;
            while (g >= groups) g -= groups, ++yy;
            while (yy < yb) {
                const int2 bd = bounds[vb + yy];
                const int* kk = kv + yy;
                const uint8_t* col = cr_rows + (bd.x - lo) * pitch + 4 * g;
                const int half = 1 << (RS_PRECISION_BITS - 1);
                int a0 = half, a1 = half, a2 = half, a3 = half;
                int ko = 0, co = 0;
                for (int k = 0; k < bd.y; ++k, ko += out_cap, co += pitch) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(col + co);
                    const int cf = __ldg(kk + ko);
                    a0 += (int)(w & 0xffu) * cf, a1 += (int)((w >> 8) & 0xffu) * cf;
                    a2 += (int)((w >> 16) & 0xffu) * cf, a3 += (int)(w >> 24) * cf;
                }
                uint8_t* d = dst + (yy * out_w + 4 * g);
                const int nv = min(4, out_w - 4 * g);
                const uint8_t v0 = clip8(a0), v1 = clip8(a1), v2 = clip8(a2), v3 = clip8(a3);
                if (nv == 4 && (reinterpret_cast<uintptr_t>(d) & 3) == 0) {
                    *reinterpret_cast<uint32_t*>(d) = (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24);
                } else if (nv == 4 && (reinterpret_cast<uintptr_t>(d) & 1) == 0) {
                    reinterpret_cast<uint16_t*>(d)[0] = (uint16_t)(v0 | (v1 << 8));
                    reinterpret_cast<uint16_t*>(d)[1] = (uint16_t)(v2 | (v3 << 8));
                } else {
                    d[0] = v0;
                    if (nv > 1) d[1] = v1;
                    if (nv > 2) d[2] = v2;
                    if (nv > 3) d[3] = v3;
                }
                g += blockDim.x;
                while (g >= groups) g -= groups, ++yy;
            }
        }
        ya = yb;
    }
}

// ---- paste: horizontal pass over the binarised g x g prediction ----
__global__ void paste_hpass_kernel(const float* __restrict__ logits, int g_h, int g_w,
                                   const int* __restrict__ jobs /*[n,6] img,x,y,w,h,rank*/,
                                   const int2* __restrict__ bounds, const int* __restrict__ coef,
                                   uint8_t* __restrict__ tmp, int out_cap, int kmax,
                                   const int* __restrict__ njobs_dev) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y, job = blockIdx.z;
    if (njobs_dev != nullptr && job >= __ldg(njobs_dev)) return;
    const int* jb = jobs + job * 6;
    const int w = jb[3], h = jb[4];
    if (w <= 0 || h <= 0 || xx >= w || xx >= out_cap) return;
    const size_t hb = ((size_t)job * 2 + 0) * out_cap + xx;
    const int2 bd = bounds[hb];
    const int* kk = coef + ((size_t)job * 2 + 0) * kmax * out_cap + xx;
    const float* row = logits + ((size_t)job * g_h + r) * g_w;
    int acc = 1 << (RS_PRECISION_BITS - 1);
    for (int k = 0; k < bd.y; ++k) acc += (row[bd.x + k] > 0x1.8p-24f ? 255 : 0) * __ldg(kk + (size_t)k * out_cap);
    tmp[((size_t)job * g_h + r) * out_cap + xx] = clip8(acc);
}

// ---- paste: vertical pass, written straight into the full-size mask.
// The reference pastes an image's boxes one after the other (loop_UCOD_DPL.py:332-351), each paste overwriting its
// whole rectangle: the final value of a pixel comes from the LAST box that covers it.  Job tables are image-major with
// ascending rank, so a job skips every pixel that a later job of the same image (entries job+1.. in `all_jobs`) covers;
// all jobs can then run in one launch, in any order, across chunks.  rank >= 0 restores the round-1 behaviour (one
// launch per rank, no look-ahead) for callers that pass unordered tables.
constexpr int PV_ROWS = 16;
__global__ void paste_vpass_kernel(const uint8_t* __restrict__ tmp, int g_h, const int* __restrict__ jobs,
                                   const int2* __restrict__ bounds, const int* __restrict__ coef,
                                   uint8_t* __restrict__ mask, int S_h, int S_w, int out_cap, int rank, int kmax,
                                   const int* __restrict__ njobs_dev, const int* __restrict__ all_jobs,
                                   int first_index, int n_all, const int* __restrict__ n_all_dev) {
    // CTA = 128 columns x PV_ROWS rows of one job's rectangle: the grid is sized for the largest admissible box, so the
    // finer one-row-per-CTA grid launched ~160 000 CTAs per chunk of which a few thousand had work (184 us per launch)
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int job = blockIdx.z;
    if (njobs_dev != nullptr && job >= __ldg(njobs_dev)) return;
    const int* jb = jobs + job * 6;
    if (rank >= 0 && jb[5] != rank) return;
    const int w = jb[3], h = jb[4];
    const int y_begin = blockIdx.y * PV_ROWS;
    if (w <= 0 || h <= 0 || xx >= w || y_begin >= h || xx >= out_cap) return;
    const int dx = jb[1] + xx;
    if (dx < 0 || dx >= S_w) return;  // Image.paste clips
    const int total = all_jobs == nullptr ? 0 : (n_all_dev != nullptr ? min(n_all, __ldg(n_all_dev)) : n_all);
    const uint8_t* src = tmp + (size_t)job * g_h * out_cap + xx;
    const int y_end = min(min(y_begin + PV_ROWS, h), out_cap);
    for (int yy = y_begin; yy < y_end; ++yy) {
        const int dy = jb[2] + yy;
        if (dy < 0 || dy >= S_h) continue;
        bool owned = false;
        for (int k = first_index + job + 1; k < total; ++k) {
            const int* o = all_jobs + (size_t)k * 6;
            if (o[0] != jb[0]) break;
            if (o[3] > 0 && o[4] > 0 && o[3] <= out_cap && o[4] <= out_cap && dx >= o[1] && dx < o[1] + o[3] &&
                dy >= o[2] && dy < o[2] + o[4]) {
                owned = true;  // a later paste of this image owns the pixel
                break;
            }
        }
        if (owned) continue;
        const size_t vb = ((size_t)job * 2 + 1) * out_cap + yy;
        const int2 bd = bounds[vb];
        const int* kk = coef + ((size_t)job * 2 + 1) * kmax * out_cap + yy;
        int acc = 1 << (RS_PRECISION_BITS - 1);
        for (int k = 0; k < bd.y; ++k) acc += (int)src[(size_t)(bd.x + k) * out_cap] * __ldg(kk + (size_t)k * out_cap);
        mask[((size_t)jb[0] * S_h + dy) * S_w + dx] = clip8(acc);
    }
}

// out = in ? mul : 0, 16 bytes per thread where the pointers allow
__global__ void mask_scale_vec_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n16, int mul) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    const uint4 v = in[i];
    const uint32_t m = (uint32_t)mul * 0x01010101u;
    auto sc = [&](uint32_t x) {  // per byte: non-zero -> mul
        uint32_t nz = (x | (x >> 1) | (x >> 2) | (x >> 3) | (x >> 4) | (x >> 5) | (x >> 6) | (x >> 7)) & 0x01010101u;
        return (nz * 0xffu) & m;
    };
    out[i] = make_uint4(sc(v.x), sc(v.y), sc(v.z), sc(v.w));
}
__global__ void mask_scale_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t n, int mul) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] ? (uint8_t)mul : 0;
}

__global__ void fill_crop_sizes_kernel(const int* jobs, int* sizes, int njobs, int out_w, int out_h,
                                       const int* __restrict__ njobs_dev) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs || (njobs_dev != nullptr && j >= __ldg(njobs_dev))) return;
    sizes[j * 2 + 0] = jobs[j * 5 + 3];
    sizes[j * 2 + 1] = jobs[j * 5 + 4];
    sizes[njobs * 2 + j * 2 + 0] = out_w;
    sizes[njobs * 2 + j * 2 + 1] = out_h;
}
__global__ void fill_paste_sizes_kernel(const int* jobs, int* sizes, int njobs, int g_w, int g_h, int out_cap,
                                        int* __restrict__ err, const int* __restrict__ njobs_dev) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs || (njobs_dev != nullptr && j >= __ldg(njobs_dev))) return;
    sizes[j * 2 + 0] = g_w;
    sizes[j * 2 + 1] = g_h;
    int w = jobs[j * 6 + 3], h = jobs[j * 6 + 4];
    if (w > out_cap || h > out_cap) {  // larger than the caller's bound: flagged, skipped
        atomicOr(err, 2);
        w = h = 0;
    }
    sizes[njobs * 2 + j * 2 + 0] = w;
    sizes[njobs * 2 + j * 2 + 1] = h;
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

// ================================================================================================
size_t lt_boxes_workspace_bytes(int B, int H, int W) {
    const size_t n = (size_t)B * H * W;
    return align256(n * 4) * 2 + align256((size_t)B * sizeof(ImgSummary)) + align256((size_t)B * LT_MAXBIG * 4) * 2 +
           align256((size_t)B * LT_MAXBIG * sizeof(BigStats)) + 1024;
}

int lt_boxes(const uint8_t* mask, int B, int H, int W, double look_twice_th, int dynamic, double const_scale,
             int* boxes, int* nbox, int* status, int* labels_out, void* workspace, size_t ws_bytes,
             cudaStream_t stream, int algorithm) {
    UCOD_REQUIRE(mask && boxes && nbox && status && workspace, "lt_boxes: null argument");
    UCOD_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "lt_boxes: bad geometry");
    UCOD_REQUIRE(ws_bytes >= lt_boxes_workspace_bytes(B, H, W), "lt_boxes: workspace too small");
    UCOD_REQUIRE(algorithm >= 0 && algorithm <= 2, "lt_boxes: algorithm must be 0 (auto), 1 (global) or 2 (shared)");
    const int n_img = H * W;
    const size_t total = (size_t)B * n_img;
    uint8_t* p = static_cast<uint8_t*>(workspace);
    int* L = reinterpret_cast<int*>(p);
    p += align256(total * 4);
    int* area = reinterpret_cast<int*>(p);
    p += align256(total * 4);
    ImgSummary* summ = reinterpret_cast<ImgSummary*>(p);
    p += align256((size_t)B * sizeof(ImgSummary));
    int* big_root = reinterpret_cast<int*>(p);
    p += align256((size_t)B * LT_MAXBIG * 4);
    int* big_area = reinterpret_cast<int*>(p);
    p += align256((size_t)B * LT_MAXBIG * 4);
    BigStats* st = reinterpret_cast<BigStats*>(p);

    const int rmax = ccl_runs_capacity(H, W);
    UCOD_REQUIRE(algorithm != 2 || rmax > 0, "lt_boxes: %d x %d masks do not fit the shared-memory labeller", H, W);
    if (algorithm != 1 && rmax > 0) {
        // run-based labelling in shared memory, one CTA per mask (no label image, no global atomics)
        const size_t smem = ccl_runs_smem_bytes(H, W, rmax);
        static size_t configured = 0;
        if (smem > configured) {
            UCOD_CHECK_CUDA(cudaFuncSetAttribute(ccl_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        {
            ProfScope ps(KC_CCL, stream, (double)total * 9);  // credited like the global path
            ccl_runs_kernel<<<B, 1024, smem, stream>>>(mask, H, W, rmax, summ, big_area, st, labels_out, 0.01);
        }
        {
            ProfScope ps(KC_CCL, stream, 0.0);
            lt_boxes_kernel<<<ceil_div(B, 32), 32, 0, stream>>>(summ, big_area, st, boxes, nbox, status, B, H, W,
                                                               look_twice_th, dynamic, const_scale);
        }
        UCOD_CHECK_CUDA(cudaGetLastError());
        return 0;
    }

    UCOD_CHECK_CUDA(cudaMemsetAsync(area, 0, total * 4, stream));
    UCOD_CHECK_CUDA(cudaMemsetAsync(summ, 0, (size_t)B * sizeof(ImgSummary), stream));
    const int T = 256;
    const unsigned lin = (unsigned)((total + T - 1) / T);
    dim3 g2(ceil_div(W, 128), H, B);
    // algorithmic bytes credited to the whole CC + box stage: the u8 mask read once plus one int32
    // label write and read per pixel = 9 B/pixel (2.4 MB per 518^2 mask); booked on the first kernel, the others add 0
    {
        ProfScope ps(KC_CCL, stream, (double)total * 9);
        ccl_init_kernel<<<lin, T, 0, stream>>>(mask, L, n_img, total);
    }
    {
        ProfScope ps(KC_CCL, stream, 0.0);
        ccl_merge_kernel<<<g2, 128, 0, stream>>>(L, H, W, B);
    }
    {
        ProfScope ps(KC_CCL, stream, 0.0);
        ccl_flatten_area_kernel<<<lin, T, 0, stream>>>(L, area, n_img, total);
    }
    lt_init_stats_kernel<<<ceil_div(B * LT_MAXBIG, T), T, 0, stream>>>(st, B * LT_MAXBIG, H, W);
    {
        ProfScope ps(KC_CCL, stream, 0.0);
        ccl_roots_kernel<<<lin, T, 0, stream>>>(L, area, summ, big_root, big_area, n_img, total, 0.01);
    }
    {
        ProfScope ps(KC_CCL, stream, 0.0);
        ccl_bbox_kernel<<<g2, 128, 0, stream>>>(L, area, st, H, W, B);
    }
    {
        ProfScope ps(KC_CCL, stream, 0.0);
        lt_boxes_kernel<<<ceil_div(B, 32), 32, 0, stream>>>(summ, big_area, st, boxes, nbox, status, B, H, W,
                                                           look_twice_th, dynamic, const_scale);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    if (labels_out) UCOD_CHECK_CUDA(cudaMemcpyAsync(labels_out, L, total * 4, cudaMemcpyDeviceToDevice, stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// tap capacity for a down-scale of in_max source samples onto out samples (Pillow: support = max(scale, 1))
static int taps_for(int in_max, int out) {
    const double scale = (double)in_max / (double)(out > 0 ? out : 1);
    const int k = (int)ceil(scale < 1.0 ? 1.0 : scale) * 2 + 1;
    return k < 5 ? 5 : k;
}

static size_t crop_ws_bytes(int njobs, int max_crop_h, int out_h, int out_w, int kmax) {
    const int cap = out_h > out_w ? out_h : out_w;
    return align256((size_t)njobs * 2 * cap * sizeof(int2)) + align256((size_t)njobs * 2 * cap * kmax * 4) +
           align256((size_t)njobs * 3 * max_crop_h * out_w) + align256((size_t)njobs * 4 * 4) + 1024;
}
size_t roi_crop_resize_workspace_bytes(int njobs, int max_crop_h, int out_h, int out_w) {
    return crop_ws_bytes(njobs, max_crop_h, out_h, out_w, RS_KMAX);
}
// crops may stick out of the image (a Look-Twice box grows by up to sqrt(2) and is then mapped to the original): the
// tap table covers crops of up to twice the source extent; larger ones set err bit 0
size_t roi_crop_resize_dyn_workspace_bytes(int capacity, int H0, int W0, int out_h, int out_w) {
    const int kh = taps_for(2 * H0, out_h), kw = taps_for(2 * W0, out_w);
    return crop_ws_bytes(capacity, H0, out_h, out_w, kh > kw ? kh : kw);
}

static int crop_resize_core(const uint8_t* images, int H0, int W0, long long img_stride, long long ch_stride,
                            long long row_stride, long long px_stride, const int* jobs, int njobs, const int* njobs_dev,
                            int max_crop_h, int kmax, uint8_t* out, int out_h, int out_w, void* workspace,
                            int* err_flag, cudaStream_t stream) {
    const int cap = out_h > out_w ? out_h : out_w;
    uint8_t* p = static_cast<uint8_t*>(workspace);
    int2* bounds = reinterpret_cast<int2*>(p);
    p += align256((size_t)njobs * 2 * cap * sizeof(int2));
    int* coef = reinterpret_cast<int*>(p);
    p += align256((size_t)njobs * 2 * cap * kmax * 4);
    uint8_t* tmp = p;
    p += align256((size_t)njobs * 3 * max_crop_h * out_w);
    int* sizes = reinterpret_cast<int*>(p);  // [njobs,2] in (w,h) then [njobs,2] out (w,h)

    // in/out size tables from the job list (device-side, no host round trip)
    fill_crop_sizes_kernel<<<ceil_div(njobs, 128), 128, 0, stream>>>(jobs, sizes, njobs, out_w, out_h, njobs_dev);
    {
        ProfScope ps(KC_RESAMPLE, stream, 0.0);  // coefficient tables are scratch, not algorithmic traffic
        dim3 g(ceil_div(cap, 128), 2, njobs);
        resample_coeffs_kernel<<<g, 128, 0, stream>>>(sizes, sizes + njobs * 2, njobs, cap, kmax, 0, bounds, coef,
                                                      err_flag, njobs_dev);
    }
    if (kmax <= CR_ROWS && (size_t)CR_ROWS * (out_w + 3) <= 200 * 1024) {
        // fused path: no intermediate image in HBM
        const size_t smem = (size_t)CR_ROWS * ((out_w + 3) & ~3);
        static size_t configured = 0;
        if (smem > configured && smem > 48 * 1024) {
            UCOD_CHECK_CUDA(cudaFuncSetAttribute(crop_resize_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem));
            configured = smem;
        }
        ProfScope ps(KC_RESAMPLE, stream, njobs_dev ? 0.0 : (double)njobs * 3 * ((double)max_crop_h * W0 + (double)out_h * out_w));
        dim3 g(ceil_div(out_h, CR_TY), 3, njobs);
        crop_resize_fused_kernel<<<g, 256, smem, stream>>>(images, H0, W0, img_stride, ch_stride, row_stride, px_stride,
                                                           jobs, bounds, coef, out, cap, out_w, out_h, kmax, njobs_dev);
        UCOD_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    {
        ProfScope ps(KC_RESAMPLE, stream, njobs_dev ? 0.0 : (double)njobs * 3 * max_crop_h * (out_w + W0));
        dim3 g(ceil_div(out_w, 128), 3 * max_crop_h, njobs);
        crop_hpass_kernel<<<g, 128, 0, stream>>>(images, H0, W0, img_stride, ch_stride, row_stride, px_stride, jobs,
                                                 bounds, coef, tmp, cap, out_w, out_h, max_crop_h, kmax, njobs_dev);
    }
    {
        ProfScope ps(KC_RESAMPLE, stream, njobs_dev ? 0.0 : (double)njobs * 3 * out_h * out_w * 3);
        dim3 g(ceil_div(out_w, 128), 3 * out_h, njobs);
        crop_vpass_kernel<<<g, 128, 0, stream>>>(tmp, jobs, bounds, coef, out, cap, out_w, out_h, max_crop_h, kmax, H0,
                                                 njobs_dev);
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int roi_crop_resize(const uint8_t* images, int n_img, int H0, int W0, long long img_stride, long long ch_stride,
                    long long row_stride, long long px_stride, const int* jobs, int njobs, int max_crop_h,
                    uint8_t* out, int out_h, int out_w, void* workspace, size_t ws_bytes, int* err_flag,
                    cudaStream_t stream) {
    UCOD_REQUIRE(images && jobs && out && workspace && err_flag, "roi_crop_resize: null argument");
    UCOD_REQUIRE(njobs > 0 && out_h > 0 && out_w > 0 && max_crop_h > 0 && n_img > 0, "roi_crop_resize: bad geometry");
    UCOD_REQUIRE(ws_bytes >= roi_crop_resize_workspace_bytes(njobs, max_crop_h, out_h, out_w),
                 "roi_crop_resize: workspace too small");
    return crop_resize_core(images, H0, W0, img_stride, ch_stride, row_stride, px_stride, jobs, njobs, nullptr,
                            max_crop_h, RS_KMAX, out, out_h, out_w, workspace, err_flag, stream);
}

// Device-count variant: `jobs` holds up to `capacity` entries of which the first *njobs_dev are valid; any crop of
// the H0 x W0 sources is supported (tap capacity and row scratch are sized for the whole source).
int roi_crop_resize_dyn(const uint8_t* images, int n_img, int H0, int W0, long long img_stride, long long ch_stride,
                        long long row_stride, long long px_stride, const int* jobs, int capacity, const int* njobs_dev,
                        uint8_t* out, int out_h, int out_w, void* workspace, size_t ws_bytes, int* err_flag,
                        cudaStream_t stream) {
    UCOD_REQUIRE(images && jobs && out && workspace && err_flag && njobs_dev, "roi_crop_resize_dyn: null argument");
    UCOD_REQUIRE(capacity > 0 && out_h > 0 && out_w > 0 && H0 > 0 && W0 > 0 && n_img > 0,
                 "roi_crop_resize_dyn: bad geometry");
    UCOD_REQUIRE(ws_bytes >= roi_crop_resize_dyn_workspace_bytes(capacity, H0, W0, out_h, out_w),
                 "roi_crop_resize_dyn: workspace too small");
    const int kh = taps_for(2 * H0, out_h), kw = taps_for(2 * W0, out_w);
    return crop_resize_core(images, H0, W0, img_stride, ch_stride, row_stride, px_stride, jobs, capacity, njobs_dev, H0,
                            kh > kw ? kh : kw, out, out_h, out_w, workspace, err_flag, stream);
}

// ------------------------------------------------------------------------------------------------
static size_t paste_ws_bytes(int njobs, int g_h, int out_cap, int kmax) {
    return align256((size_t)njobs * 2 * out_cap * sizeof(int2)) + align256((size_t)njobs * 2 * out_cap * kmax * 4) +
           align256((size_t)njobs * g_h * out_cap) + align256((size_t)njobs * 4 * 4) + 1024;
}
size_t paste_bicubic_workspace_bytes(int njobs, int g_h, int out_cap) {
    return paste_ws_bytes(njobs, g_h, out_cap, RS_KMAX);
}
// bicubic support is 2 * max(scale, 1): a g x g map pasted into >= 1 pixel
static int paste_taps(int g_h, int g_w) { return 2 * taps_for(g_h > g_w ? g_h : g_w, 1); }
size_t paste_bicubic_dyn_workspace_bytes(int capacity, int g_h, int g_w, int out_cap) {
    return paste_ws_bytes(capacity, g_h, out_cap, paste_taps(g_h, g_w));
}

static int paste_core(const float* logits, int njobs, const int* njobs_dev, int g_h, int g_w, const int* jobs,
                      int max_rank, const int* all_jobs, int first_index, int n_all, const int* n_all_dev,
                      uint8_t* mask, int S_h, int S_w, int out_cap, int kmax, void* workspace, int* err_flag,
                      cudaStream_t stream) {
    uint8_t* p = static_cast<uint8_t*>(workspace);
    int2* bounds = reinterpret_cast<int2*>(p);
    p += align256((size_t)njobs * 2 * out_cap * sizeof(int2));
    int* coef = reinterpret_cast<int*>(p);
    p += align256((size_t)njobs * 2 * out_cap * kmax * 4);
    uint8_t* tmp = p;
    p += align256((size_t)njobs * g_h * out_cap);
    int* sizes = reinterpret_cast<int*>(p);
    fill_paste_sizes_kernel<<<ceil_div(njobs, 128), 128, 0, stream>>>(jobs, sizes, njobs, g_w, g_h, out_cap, err_flag,
                                                                     njobs_dev);
    {
        ProfScope ps(KC_RESAMPLE, stream, 0.0);
        dim3 g(ceil_div(out_cap, 128), 2, njobs);
        resample_coeffs_kernel<<<g, 128, 0, stream>>>(sizes, sizes + njobs * 2, njobs, out_cap, kmax, 1, bounds, coef,
                                                      err_flag, njobs_dev);
    }
    {
        ProfScope ps(KC_RESAMPLE, stream, njobs_dev ? 0.0 : (double)njobs * g_h * (out_cap + g_w * 4));
        dim3 g(ceil_div(out_cap, 128), g_h, njobs);
        paste_hpass_kernel<<<g, 128, 0, stream>>>(logits, g_h, g_w, jobs, bounds, coef, tmp, out_cap, kmax, njobs_dev);
    }
    dim3 g(ceil_div(out_cap, 128), ceil_div(out_cap, PV_ROWS), njobs);
    if (all_jobs != nullptr) {
        ProfScope ps(KC_RESAMPLE, stream, 0.0);
        paste_vpass_kernel<<<g, 128, 0, stream>>>(tmp, g_h, jobs, bounds, coef, mask, S_h, S_w, out_cap, -1, kmax,
                                                  njobs_dev, all_jobs, first_index, n_all, n_all_dev);
    } else {
        for (int r = 0; r <= max_rank; ++r) {
            ProfScope ps(KC_RESAMPLE, stream, (double)njobs * out_cap * out_cap / (max_rank + 1));
            paste_vpass_kernel<<<g, 128, 0, stream>>>(tmp, g_h, jobs, bounds, coef, mask, S_h, S_w, out_cap, r, kmax,
                                                      njobs_dev, nullptr, 0, 0, nullptr);
        }
    }
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int paste_bicubic(const float* logits, int njobs, int g_h, int g_w, const int* jobs, int max_rank, uint8_t* mask,
                  int n_img, int S_h, int S_w, int out_cap, void* workspace, size_t ws_bytes, int* err_flag,
                  cudaStream_t stream) {
    UCOD_REQUIRE(logits && jobs && mask && workspace && err_flag, "paste_bicubic: null argument");
    UCOD_REQUIRE(njobs > 0 && g_h > 0 && g_w > 0 && out_cap > 0 && max_rank >= 0, "paste_bicubic: bad geometry");
    UCOD_REQUIRE(ws_bytes >= paste_bicubic_workspace_bytes(njobs, g_h, out_cap), "paste_bicubic: workspace too small");
    (void)n_img;
    return paste_core(logits, njobs, nullptr, g_h, g_w, jobs, max_rank, nullptr, 0, 0, nullptr, mask, S_h, S_w,
                      out_cap, RS_KMAX, workspace, err_flag, stream);
}

// Device-count variant for one chunk of an image-major, rank-ascending job table: `jobs` = all_jobs + first_index*6,
// *njobs_dev of its `capacity` entries are valid, the table holds min(n_all, *n_all_dev) entries in total.
int paste_bicubic_dyn(const float* logits, int capacity, const int* njobs_dev, int g_h, int g_w, const int* all_jobs,
                      int first_index, int n_all, const int* n_all_dev, uint8_t* mask, int n_img, int S_h, int S_w,
                      int out_cap, void* workspace, size_t ws_bytes, int* err_flag, cudaStream_t stream) {
    UCOD_REQUIRE(logits && all_jobs && mask && workspace && err_flag && njobs_dev && n_all_dev,
                 "paste_bicubic_dyn: null argument");
    UCOD_REQUIRE(capacity > 0 && g_h > 0 && g_w > 0 && out_cap > 0 && first_index >= 0 && n_all >= first_index,
                 "paste_bicubic_dyn: bad geometry");
    UCOD_REQUIRE(ws_bytes >= paste_bicubic_dyn_workspace_bytes(capacity, g_h, g_w, out_cap),
                 "paste_bicubic_dyn: workspace too small");
    (void)n_img;
    return paste_core(logits, capacity, njobs_dev, g_h, g_w, all_jobs + (size_t)first_index * 6, 0, all_jobs,
                      first_index, n_all, n_all_dev, mask, S_h, S_w, out_cap, paste_taps(g_h, g_w), workspace,
                      err_flag, stream);
}

// ------------------------------------------------------------------------------------------------
// Look-Twice job tables on the device (replaces the host loop of loop_UCOD_DPL.py:331-342 and `resize_bbox`,
// :387-397): image b contributes max(nbox[b], 0) jobs, image-major, rank = position in the sorted box list.
// crop job = box mapped to the original image with CPython's float arithmetic (`int(v * (new / old))`, fp64,
// truncation); paste job = the box itself in mask coordinates.
// counts[0] = jobs written (<= capacity), counts[1] = status bits: 1 an image's box maths raised ValueError
// (nbox == -2), 2 more jobs than `capacity` (extra ones dropped), 4 a box or mapped crop with w or h <= 0 (PIL raises
// in the reference; the job is kept with its size zeroed so that crop / paste skip it), counts[2] = total requested.
// chunk_counts[c] = number of valid jobs in [c*chunk, (c+1)*chunk).
__global__ void lt_build_jobs_kernel(const int* __restrict__ boxes, const int* __restrict__ nbox, int B, int S_h,
                                     int S_w, int src_h, int src_w, const int* __restrict__ orig_sizes,
                                     int* __restrict__ crop_jobs, int* __restrict__ paste_jobs, int capacity,
                                     int* __restrict__ counts, int chunk, int n_chunks, int* __restrict__ chunk_counts) {
    extern __shared__ int offs[];  // [B + 1] exclusive prefix of the per-image job counts
    __shared__ int status;
    if (threadIdx.x == 0) status = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int n = nbox[b];
        if (n == -2) atomicOr(&status, 1);
        offs[b + 1] = n > 0 ? (n > LT_MAXBIG ? LT_MAXBIG : n) : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        offs[0] = 0;
        for (int b = 0; b < B; ++b) offs[b + 1] += offs[b];
    }
    __syncthreads();
    const int total = offs[B];
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int n = offs[b + 1] - offs[b];
        const int H0 = orig_sizes ? orig_sizes[2 * b] : src_h, W0 = orig_sizes ? orig_sizes[2 * b + 1] : src_w;
        const double ws = __ddiv_rn((double)W0, (double)S_w), hs = __ddiv_rn((double)H0, (double)S_h);
        for (int i = 0; i < n; ++i) {
            const int slot = offs[b] + i;
            if (slot >= capacity) break;
            const int* bb = boxes + ((size_t)b * LT_MAXBIG + i) * 4;
            int x = (int)__dmul_rn((double)bb[0], ws), y = (int)__dmul_rn((double)bb[1], hs);
            int w = (int)__dmul_rn((double)bb[2], ws), h = (int)__dmul_rn((double)bb[3], hs);
            int pw = bb[2], ph = bb[3];
            if (w <= 0 || h <= 0 || pw <= 0 || ph <= 0) {
                atomicOr(&status, 4);
                w = h = pw = ph = 0;
            }
            int* cj = crop_jobs + (size_t)slot * 5;
            cj[0] = b, cj[1] = x, cj[2] = y, cj[3] = w, cj[4] = h;
            int* pj = paste_jobs + (size_t)slot * 6;
            pj[0] = b, pj[1] = bb[0], pj[2] = bb[1], pj[3] = pw, pj[4] = ph, pj[5] = i;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int kept = total < capacity ? total : capacity;
        counts[0] = kept;
        counts[1] = status | (total > capacity ? 2 : 0);
        counts[2] = total;
        for (int c = 0; c < n_chunks; ++c) {
            const int left = kept - c * chunk;
            chunk_counts[c] = left < 0 ? 0 : (left > chunk ? chunk : left);
        }
    }
}

int lt_build_jobs(const int* boxes, const int* nbox, int B, int S_h, int S_w, int src_h, int src_w,
                  const int* orig_sizes, int* crop_jobs, int* paste_jobs, int capacity, int* counts, int chunk,
                  int* chunk_counts, cudaStream_t stream) {
    UCOD_REQUIRE(boxes && nbox && crop_jobs && paste_jobs && counts && chunk_counts, "lt_build_jobs: null argument");
    UCOD_REQUIRE(B > 0 && B <= 8192 && S_h > 0 && S_w > 0 && capacity > 0 && chunk > 0 &&
                     (orig_sizes != nullptr || (src_h > 0 && src_w > 0)),
                 "lt_build_jobs: bad geometry");
    const int n_chunks = ceil_div(capacity, chunk);
    ProfScope ps(KC_CCL, stream, 0.0);
    lt_build_jobs_kernel<<<1, 256, (size_t)(B + 1) * sizeof(int), stream>>>(boxes, nbox, B, S_h, S_w, src_h, src_w,
                                                                            orig_sizes, crop_jobs, paste_jobs,
                                                                            capacity, counts, chunk, n_chunks,
                                                                            chunk_counts);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int mask_scale_u8(const uint8_t* in, uint8_t* out, size_t n, int mul, cudaStream_t stream) {
    UCOD_REQUIRE(in && out, "mask_scale_u8: null argument");
    if (n == 0) return 0;
    ProfScope ps(KC_RESAMPLE, stream, (double)n * 2);
    size_t head = 0;
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 && n >= 16) {
        const size_t n16 = n / 16;
        mask_scale_vec_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, stream>>>(
                reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), n16, mul);
        head = n16 * 16;
    }
    if (head < n)
        mask_scale_kernel<<<(unsigned)((n - head + 255) / 256), 256, 0, stream>>>(in + head, out + head, n - head, mul);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// torchvision ToTensor (u8 -> fp32 / 255) followed by Normalize ((x - mean) / std), each step one correctly
// rounded fp32 operation like the torch elementwise ops it replaces (data/datasets/transforms.py:14-18).
// Planar [planes, hw] with channel = plane % channels; channels == 0 skips the normalisation (label transform).
__global__ void to_tensor_normalize_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, size_t n, int hw,
                                           int channels, float m0, float m1, float m2, float s0, float s1, float s2) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    uint8_t v[4];
    const bool full = i + 4 <= n;
    if (full) {
        *reinterpret_cast<uint32_t*>(v) = *reinterpret_cast<const uint32_t*>(in + i);
    } else {
        for (int k = 0; k < 4; ++k) v[k] = i + k < n ? in[i + k] : 0;
    }
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float x = __fdiv_rn((float)v[k], 255.0f);
        if (channels > 0) {
            const int c = (int)(((i + k) / (size_t)hw) % (size_t)channels);
            const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
            x = __fdiv_rn(__fsub_rn(x, m), sd);
        }
        r[k] = x;
    }
    if (full) {
        *reinterpret_cast<float4*>(out + i) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
        for (int k = 0; k < 4 && i + k < n; ++k) out[i + k] = r[k];
    }
}

int to_tensor_normalize(const uint8_t* in, float* out, size_t planes, int hw, int channels, const float* mean,
                        const float* stddev, cudaStream_t stream) {
    UCOD_REQUIRE(in && out, "to_tensor_normalize: null argument");
    UCOD_REQUIRE(channels == 0 || (channels <= 3 && mean && stddev), "to_tensor_normalize: 0..3 channels with mean/std");
    UCOD_REQUIRE((reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "to_tensor_normalize: input must be 4-byte and output 16-byte aligned");
    const size_t n = planes * (size_t)hw;
    if (n == 0) return 0;
    float m[3] = {0, 0, 0}, s[3] = {1, 1, 1};
    for (int c = 0; c < channels; ++c) m[c] = mean[c], s[c] = stddev[c];
    ProfScope ps(KC_RESAMPLE, stream, (double)n * 5);
    to_tensor_normalize_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, stream>>>(in, out, n, hw, channels, m[0],
                                                                                       m[1], m[2], s[0], s[1], s[2]);
    UCOD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ucod
