// Fused attention forward (tcgen05), see attention.cu.
#pragma once
#include "common.cuh"

namespace ucod {

// q, k: [B*H, T, 64] bf16 ; vt: [B*H, 64, Tpad] bf16 (columns >= T must be finite, normally zero) ;
// ctx: [B, T, H*64] bf16.  scale = 1/sqrt(head_dim).
int launch_attention_d64(const void* q, const void* k, const void* vt, void* ctx, int B, int H, int T, int Tpad,
                         float scale, cudaStream_t stream);

}  // namespace ucod
