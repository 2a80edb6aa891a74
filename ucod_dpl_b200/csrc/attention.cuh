// Fused attention forward (tcgen05), see attention.cu.
#pragma once
#include "common.cuh"

namespace ucod {

// q: [batch, tokens_q, ld_q] bf16, k / v: [batch, tokens_kv, ld_kv] bf16 — head h occupies columns
// [h*head_dim, (h+1)*head_dim) of each pointer (so the three pointers may address the Q / K / V column blocks of one
// fused projection output).  ctx: [batch, tokens_q, ld_ctx] bf16, same head layout.  head_dim 64 or 128;
// head_dim_real (<= head_dim) only feeds the FLOP accounting when the operands are zero-padded.
struct AttentionArgs {
    const void* q = nullptr;
    const void* k = nullptr;
    const void* v = nullptr;
    void* ctx = nullptr;
    int batch = 0, heads = 0, tokens_q = 0, tokens_kv = 0;
    int head_dim = 64, head_dim_real = 64;
    int ld_q = 0, ld_kv = 0, ld_ctx = 0;
    float scale = 0.125f;
    int kv_batch = 0;                   // number of K/V batches when kv_batch_map is used (0: same as batch)
    const int* kv_batch_map = nullptr;  // optional [batch]: K/V batch index of each q batch (cross-attention sharing)
    const int* batch_dev = nullptr;     // optional device-side batch count (<= batch): CTAs of batches beyond it exit
};
int launch_attention(const AttentionArgs& args, cudaStream_t stream);

}  // namespace ucod
