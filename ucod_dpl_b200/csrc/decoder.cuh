// DBA decoder forward + layout / upsample helpers (see decoder.cu).
#pragma once
#include "common.cuh"

namespace ucod {

struct DecoderWeights {
    int dim;              // input channels (768)
    const void* w_dec;    // bf16 [128, dim]      decoupling.weight
    const float* b_dec;   // [128]                decoupling.bias
    const float* emb;     // [2, 64]              learnable_embedding
    const float* w_fg;    // [64]                 conv_out_fg.weight
    const float* b_fg;    // [1]
    const float* w_bg;    // [64]                 conv_out_bg.weight
    const float* b_bg;    // [1]
};

size_t decoder_workspace_bytes(int B, int gin_h, int gin_w, int out_h, int out_w, int want_ortho);
int decoder_forward(const void* keys_bf16, int B, int gin_h, int gin_w, int out_h, int out_w, const DecoderWeights& w,
                    float* fg, float* bg, float* ortho, void* workspace, size_t ws_bytes, cudaStream_t stream,
                    const int* batch_dev = nullptr);
struct DecoderGrads {  // fp32 gradients, same shapes as the weights (w_dec as fp32 [128, dim])
    float* w_dec; float* b_dec; float* w_fg; float* b_fg; float* w_bg; float* b_bg;
};
size_t decoder_backward_workspace_bytes(int B, int gin_h, int gin_w, int out_h, int out_w);
// fwd_workspace: the workspace a decoder_forward(... ortho != NULL ...) call on the same inputs left behind.
// loss2: device float[2] = {BCEwL(fg, target), BCEwL(bg, 1 - target)} (means over B*npix).
int decoder_backward(const void* keys_bf16, int B, int gin_h, int gin_w, int out_h, int out_w, const DecoderWeights& w,
                     const float* fg, const float* bg, const float* target, const float* dfg, const float* dbg,
                     const float* dortho, void* fwd_workspace, size_t fwd_ws_bytes, const DecoderGrads& g, float* loss2,
                     void* workspace, size_t ws_bytes, cudaStream_t stream);
int train_loss(const float* loss2, const float* ortho, const float* dis_loss, float* out, cudaStream_t stream);
int adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, size_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step_t, float grad_scale, float ema_alpha,
                   cudaStream_t stream);
int features_to_tokens_bf16(const float* in, void* out, int B, int C, int P, long long sb, long long sc, long long sp,
                            cudaStream_t stream);
int upsample_bilinear(const float* in, void* out, int B, int in_h, int in_w, int out_h, int out_w, int binarize,
                      cudaStream_t stream);

}  // namespace ucod
