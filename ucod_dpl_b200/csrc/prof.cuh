// Launch accounting + optional per-kernel-class CUDA-event timing (used by bench.py for the roofline line).
#pragma once
#include <cuda_runtime.h>

namespace ucod {

enum KernelClass : int {
    KC_GEMM = 0,       // tcgen05 GEMM                     (work = FLOPs)
    KC_ATTENTION = 1,  // tcgen05 fused attention          (work = FLOPs)
    KC_LAYERNORM = 2,  // LayerNorm                        (work = bytes)
    KC_EMBED = 3,      // im2col / CLS init / CLS-row attn (work = bytes)
    KC_DECODER = 4,    // decoder norm / gate / heads      (work = bytes)
    KC_RESAMPLE = 5,   // bilinear / PIL-exact resampling  (work = bytes)
    KC_PSEUDO = 6,     // pseudo-label scoring / cleanup   (work = bytes)
    KC_CCL = 7,        // connected components / boxes     (work = bytes)
    KC_OTHER = 8,
    KC_COUNT = 9
};

void prof_pre(int cls, cudaStream_t s, double work);
void prof_post(int cls, cudaStream_t s);

struct ProfScope {
    int cls;
    cudaStream_t s;
    ProfScope(int c, cudaStream_t st, double work) : cls(c), s(st) { prof_pre(c, st, work); }
    ~ProfScope() { prof_post(cls, s); }
};

void prof_enable(int on);
// Sums since the last call; synchronises the recorded events. Arrays of KC_COUNT entries.
int prof_collect(double* ms, double* work, long long* launches);
long long launch_count_total();

}  // namespace ucod
