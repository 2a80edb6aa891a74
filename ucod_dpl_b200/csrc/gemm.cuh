// Interface of the tcgen05 GEMM used by every dense contraction on the path
// (patch-embed, QKV, out-proj, MLP, last-layer key projection, decoder 1x1 conv, CORAL CSF).
#pragma once
#include "common.cuh"

namespace ucod {

enum GemmEpiMode : int {
    EPI_BIAS_BF16 = 0,       // out_bf16[m,n] = acc + bias[n]
    EPI_BIAS_GELU_BF16 = 1,  // out_bf16[m,n] = gelu_erf(acc + bias[n])
    EPI_RESID_F32 = 2,       // x_f32[m,n] += acc + bias[n]   (TMA reduce-add; LayerScale is folded into W, bias)
    EPI_PATCH = 4,           // x_f32[b*(P+1)+1+p, n] = acc + bias[n] + pos[1+p, n]   (m = b*P + p)
    EPI_BIAS_F32 = 5,        // out_f32[m,n] = acc + bias[n]
    EPI_KEYS = 6,            // last-layer key projection: drop `skip` leading tokens per image, write
                             //   out (fp32, optional) and out2 (bf16, optional) as [B*(T-skip), N]
};

struct GemmEpi {
    int mode = EPI_BIAS_BF16;
    const float* bias = nullptr;   // [N]
    const float* pos = nullptr;    // [T, N] (EPI_PATCH)
    void* out = nullptr;
    void* out2 = nullptr;
    int ld_out = 0;      // row pitch of out/out2 in elements
    int tokens = 0;      // T (EPI_KEYS) or P (EPI_PATCH)
    int skip = 0;        // EPI_KEYS
    // Device-side row count (second Look-Twice pass: the number of crops is only known on the device).  When
    // m_dev != nullptr the kernel processes min(M, *m_dev * m_per) rows; M stays the capacity the tensor maps and
    // buffers are sized for.  No host synchronisation is involved.
    const int* m_dev = nullptr;
    int m_per = 1;
};

// D[M,N] = A[M,K] * W[N,K]^T with fp32 accumulation in TMEM; A, W bf16 row-major (K contiguous).
// lda / ldw are row pitches in elements (multiples of 8). N must be a multiple of 128.
int launch_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpi& ep,
                     cudaStream_t stream);

}  // namespace ucod
