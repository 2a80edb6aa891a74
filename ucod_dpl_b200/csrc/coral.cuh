// CORAL second-stage refiner kernels (see coral.cu).
#pragma once
#include "common.cuh"

namespace ucod {

int coral_entropy_select(const float* preds, int B, int P, int window_size, float threshold, float* entropy,
                         float* scores, uint8_t* mask, int* flag_scratch, cudaStream_t stream, int per_image = 0);
int coral_window_head(const float* taps, int ld_taps, int n_windows, int g, float bias_const, float* out,
                      cudaStream_t stream);
int coral_scatter_windows(const float* window_preds, const int* slot_of_cell, int B, int window_size, int g, float* out,
                          cudaStream_t stream);
size_t coral_gated_ensemble_workspace_bytes(int B, int S);
int coral_gated_ensemble(const float* preds, int P, const float* h_preds, int B, int S, int max_per_image,
                         const float* w0, const float* b0, const float* w2, const float* b2, float* out, float* weight,
                         void* workspace, size_t ws_bytes, cudaStream_t stream);
int layernorm_rows_bf16(const float* x, const float* w, const float* b, void* y, int rows, int dim, float eps,
                        cudaStream_t stream);
int cast_f32_to_bf16(const float* in, void* out, size_t n, cudaStream_t stream);
int features_to_tokens_f32(const float* in, float* out, int B, int C, int P, long long sb, long long sc, long long sp,
                           cudaStream_t stream);
int resize_tokens_bilinear(const float* in, float* out_f32, void* out_bf16, int n, int gin_h, int gin_w, int gout_h,
                           int gout_w, int C, cudaStream_t stream);

}  // namespace ucod
