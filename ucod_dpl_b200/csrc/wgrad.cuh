// Tensor-core weight gradient of the decoder's 1x1 conv (see wgrad.cu).
#pragma once
#include "common.cuh"

namespace ucod {

size_t wgrad_workspace_bytes(int T, int dim);
// Layout of the contraction's buffers inside `workspace` (256-byte aligned, >= wgrad_workspace_bytes).  The producer of
// dD (decoder.cu) fills dDt [128, t_pad] bf16 (zero beyond T) and bpart [n_kblocks, 128] (column sums of 64-row blocks).
struct WgradPlan {
    __nv_bfloat16* dDt;
    float* partial;
    float* bpart;
    int n_kblocks, t_pad, splits, kb_per_split;
};
int wgrad_plan(int T, int dim, void* workspace, size_t ws_bytes, WgradPlan* plan);
// keys_bf16 [T, dim] token-major.  Writes dW [128, dim] and db [128] (no pre-zeroing needed, bit-reproducible).
int wgrad_contract(const WgradPlan& plan, const void* keys_bf16, int T, int dim, float* dW, float* db,
                   cudaStream_t stream);

}  // namespace ucod
