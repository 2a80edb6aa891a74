// Tensor-core weight gradient of the decoder's 1x1 conv (see wgrad.cu).
#pragma once
#include "common.cuh"

namespace ucod {

size_t wgrad_workspace_bytes(int T, int dim);
// a_buf / e_buf [T,128], tsum / sumsq [B,128], emb [128]: the backward's intermediates (decoder.cu);
// keys_bf16 [T, dim] token-major.  Writes dW [128, dim] and db [128] (no pre-zeroing needed, bit-reproducible).
int wgrad_tensor_core(const float* a_buf, const float* e_buf, const float* tsum, const float* sumsq, const float* emb,
                      const void* keys_bf16, int rows_per_img, int T, int dim, float* dW, float* db, void* workspace,
                      size_t ws_bytes, cudaStream_t stream);

}  // namespace ucod
