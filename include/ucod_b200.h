/* ucod_b200 — C-ABI of the B200-native UCOD-DPL hot path.
 *
 * Plain C: raw device/host pointers, sizes and a `void* stream` (a cudaStream_t).  No torch types.
 * Every function returns 0 on success; non-zero means failure and `ucod_last_error()` holds the reason
 * (the Python host raises RuntimeError with it — the reference raises Python exceptions on its path,
 * engine/runner/runner.py:122-123,273-274).  Outputs are always caller-allocated; the library never
 * allocates result buffers (reference ownership model: all tensors belong to the caller's allocator).
 *
 * The reference (Heartfirey/UCOD-DPL) has no FFI layer: its boundary is the Python object API
 * (SURVEY.md §8b).  Each entry point below names the reference call it replaces (file:line relative to
 * the reference root); INTEGRATION.md shows the ctypes binding a maintainer would add.
 */
#ifndef UCOD_B200_H
#define UCOD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UCOD_B200_ABI_VERSION 1

/* Last error message of the calling thread ("" if none). */
const char* ucod_last_error(void);
int ucod_abi_version(void);

/* ---- dense building block -------------------------------------------------------------------
 * out[M,N] = epilogue(A[M,K] * W[N,K]^T) ; A, W bf16 row-major (K contiguous), fp32 accumulate (tcgen05).
 * epi_mode: 0 = bf16 out, +bias ; 1 = bf16 out, gelu(+bias) ; 2 = fp32 in-place residual
 *           out += scale*(acc+bias) ; 5 = fp32 out, +bias.   bias/scale may be NULL.
 * Replaces: torch.nn.Linear / 1x1 Conv2d library GEMMs on the path (HF modeling_dinov2.py:153-235,348-387;
 * models/modules/DBA.py:13,35). */
int ucod_gemm_bf16(const void* a, int lda, const void* w, int ldw, int m, int n, int k, int epi_mode,
                   const float* bias, const float* scale, void* out, int ld_out, void* stream);

/* Fused softmax(q k^T * scale) v, head_dim 64, non-causal (tcgen05 flash-attention forward).
 * q,k: [batch*heads, tokens, 64] bf16 ; vt: [batch*heads, 64, tokens_pad] bf16 (V transposed, pad columns zero,
 * tokens_pad % 8 == 0) ; ctx: [batch, tokens, heads*64] bf16.
 * Replaces: HF Dinov2SelfAttention/ViTSelfAttention (modeling_dinov2.py:153-235) under
 * data/utils/feature_extractor.py:49-59. */
int ucod_attention_d64(const void* q, const void* k, const void* vt, void* ctx, int batch, int heads, int tokens,
                       int tokens_pad, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UCOD_B200_H */
