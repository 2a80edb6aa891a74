// C-ABI entry points: library info, error string, and the raw GEMM building block (used by tests/bench).
#include "../../include/ucod_b200.h"
#include "gemm.cuh"
#include "attention.cuh"
#include "coral.cuh"
#include "vit.cuh"
#include "decoder.cuh"
#include "prof.cuh"
#include "pseudo_label.cuh"
#include "looktwice.cuh"
#include "discriminator.cuh"
#include "metrics.cuh"

using namespace ucod;

extern "C" {

const char* ucod_last_error(void) { return get_last_error(); }

int ucod_abi_version(void) { return UCOD_B200_ABI_VERSION; }

void ucod_prof_enable(int on) { prof_enable(on); }
int ucod_prof_collect(double* ms, double* work, long long* launches) { return prof_collect(ms, work, launches); }
long long ucod_launch_count(void) { return launch_count_total(); }

int ucod_gemm_bf16(const void* a, int lda, const void* w, int ldw, int m, int n, int k, int epi_mode,
                   const float* bias, void* out, int ld_out, void* stream) {
    UCOD_REQUIRE(a && w && out, "ucod_gemm_bf16: null pointer");
    UCOD_REQUIRE(epi_mode == EPI_BIAS_BF16 || epi_mode == EPI_BIAS_GELU_BF16 || epi_mode == EPI_RESID_F32 ||
                     epi_mode == EPI_BIAS_F32,
                 "ucod_gemm_bf16: epilogue mode %d is internal to the ViT pipeline", epi_mode);
    GemmEpi ep;
    ep.mode = epi_mode;
    ep.bias = bias;
    ep.out = out;
    ep.ld_out = ld_out;
    return launch_gemm_bf16(a, lda, w, ldw, m, n, k, ep, reinterpret_cast<cudaStream_t>(stream));
}

int ucod_attention(const void* q, int ld_q, const void* k, const void* v, int ld_kv, void* ctx, int ld_ctx, int batch,
                   int heads, int head_dim, int tokens_q, int tokens_kv, float scale, void* stream) {
    return ucod_attention_shared_kv(q, ld_q, k, v, ld_kv, ctx, ld_ctx, batch, heads, head_dim, head_dim, tokens_q,
                                    tokens_kv, scale, nullptr, 0, stream);
}
int ucod_attention_shared_kv(const void* q, int ld_q, const void* k, const void* v, int ld_kv, void* ctx, int ld_ctx,
                             int batch, int heads, int head_dim, int head_dim_real, int tokens_q, int tokens_kv,
                             float scale, const int32_t* kv_batch_map, int kv_batch, void* stream) {
    AttentionArgs a;
    a.q = q, a.k = k, a.v = v, a.ctx = ctx;
    a.batch = batch, a.heads = heads, a.tokens_q = tokens_q, a.tokens_kv = tokens_kv;
    a.head_dim = head_dim, a.head_dim_real = head_dim_real;
    a.ld_q = ld_q, a.ld_kv = ld_kv, a.ld_ctx = ld_ctx;
    a.scale = scale;
    a.kv_batch_map = kv_batch_map;
    a.kv_batch = kv_batch;
    return launch_attention(a, reinterpret_cast<cudaStream_t>(stream));
}

int ucod_vit_create(void** handle, const ucod_vit_cfg* cfg, const void* patch_w, const float* patch_b,
                    const float* cls_token, const ucod_vit_layer* layers) {
    return vit_create(handle, cfg, patch_w, patch_b, cls_token, layers);
}
int ucod_vit_destroy(void* handle) { return vit_destroy(handle); }
int ucod_vit_workspace_bytes(void* handle, int batch, int img_h, int img_w, uint64_t* bytes) {
    size_t b = 0;
    int rc = vit_workspace_bytes(handle, batch, img_h, img_w, &b);
    if (rc == 0 && bytes) *bytes = (uint64_t)b;
    return rc;
}
int ucod_vit_keys(void* handle, const void* images, int image_dtype, int batch, int img_h, int img_w,
                  const float* pos_emb, void* workspace, uint64_t workspace_bytes, float* keys_f32, void* keys_bf16,
                  float* cls_attn, int keep_cls, void* stream) {
    return vit_keys(handle, images, image_dtype, batch, img_h, img_w, pos_emb, workspace, (size_t)workspace_bytes,
                    keys_f32, keys_bf16, cls_attn, keep_cls, reinterpret_cast<cudaStream_t>(stream));
}

int ucod_vit_keys_dyn(void* handle, const void* images, int image_dtype, int batch, const int32_t* batch_dev, int img_h,
                      int img_w, const float* pos_emb, void* workspace, uint64_t workspace_bytes, float* keys_f32,
                      void* keys_bf16, float* cls_attn, int keep_cls, void* stream) {
    return vit_keys(handle, images, image_dtype, batch, img_h, img_w, pos_emb, workspace, (size_t)workspace_bytes,
                    keys_f32, keys_bf16, cls_attn, keep_cls, reinterpret_cast<cudaStream_t>(stream), batch_dev);
}

uint64_t ucod_decoder_workspace_bytes(int batch, int gin_h, int gin_w, int out_h, int out_w, int want_ortho) {
    return (uint64_t)decoder_workspace_bytes(batch, gin_h, gin_w, out_h, out_w, want_ortho);
}
int ucod_decoder_fwd(const void* keys_bf16, int batch, int dim, int gin_h, int gin_w, int out_h, int out_w,
                     const void* w_dec, const float* b_dec, const float* emb, const float* w_fg, const float* b_fg,
                     const float* w_bg, const float* b_bg, float* fg, float* bg, float* ortho, void* workspace,
                     uint64_t workspace_bytes, void* stream) {
    UCOD_REQUIRE(w_dec && b_dec && emb && w_fg && b_fg && w_bg && b_bg, "ucod_decoder_fwd: null weight pointer");
    DecoderWeights w{dim, w_dec, b_dec, emb, w_fg, b_fg, w_bg, b_bg};
    return decoder_forward(keys_bf16, batch, gin_h, gin_w, out_h, out_w, w, fg, bg, ortho, workspace,
                           (size_t)workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_decoder_fwd_dyn(const void* keys_bf16, int batch, const int32_t* batch_dev, int dim, int gin_h, int gin_w,
                         int out_h, int out_w, const void* w_dec, const float* b_dec, const float* emb,
                         const float* w_fg, const float* b_fg, const float* w_bg, const float* b_bg, float* fg,
                         float* bg, void* workspace, uint64_t workspace_bytes, void* stream) {
    UCOD_REQUIRE(w_dec && b_dec && emb && w_fg && b_fg && w_bg && b_bg, "ucod_decoder_fwd_dyn: null weight pointer");
    DecoderWeights w{dim, w_dec, b_dec, emb, w_fg, b_fg, w_bg, b_bg};
    return decoder_forward(keys_bf16, batch, gin_h, gin_w, out_h, out_w, w, fg, bg, nullptr, workspace,
                           (size_t)workspace_bytes, reinterpret_cast<cudaStream_t>(stream), batch_dev);
}
int ucod_features_to_tokens_bf16(const float* in, void* out, int batch, int channels, int pixels, int64_t sb,
                                 int64_t sc, int64_t sp, void* stream) {
    return features_to_tokens_bf16(in, out, batch, channels, pixels, sb, sc, sp,
                                   reinterpret_cast<cudaStream_t>(stream));
}
int ucod_upsample_bilinear(const float* in, void* out, int batch, int in_h, int in_w, int out_h, int out_w,
                           int binarize, void* stream) {
    return upsample_bilinear(in, out, batch, in_h, in_w, out_h, out_w, binarize,
                             reinterpret_cast<cudaStream_t>(stream));
}

int ucod_pseudo_label_score(const float* attn_cls, const void* keys, int keys_bf16, int batch, int heads, int patches,
                            float th_bkg, float epsilon, float* cos, uint8_t* bkg, int32_t* ref_idx, float* sim,
                            void* scratch, void* stream) {
    return pseudo_label_score(attn_cls, keys, keys_bf16, batch, heads, patches, th_bkg, epsilon, cos, bkg, ref_idx,
                              sim, static_cast<int*>(scratch), reinterpret_cast<cudaStream_t>(stream));
}
uint64_t ucod_pseudo_label_scratch_bytes(int batch, int heads) {
    return (uint64_t)pseudo_label_scratch_bytes(batch, heads);
}
int ucod_pseudo_label_score_ex(const float* attn_cls, const void* keys, int keys_bf16, int batch, int heads, int patches,
                               float th_bkg, float epsilon, int apply_weights, float* cos, uint8_t* bkg,
                               int32_t* ref_idx, float* sim, void* scratch, uint64_t scratch_bytes, void* stream) {
    return pseudo_label_score(attn_cls, keys, keys_bf16, batch, heads, patches, th_bkg, epsilon, cos, bkg, ref_idx,
                              sim, static_cast<int*>(scratch), reinterpret_cast<cudaStream_t>(stream), apply_weights,
                              (size_t)scratch_bytes);
}
int ucod_refine_small_components(const uint8_t* mask_in, uint8_t* mask_out, int batch, int h, int w,
                                 int area_threshold, void* stream) {
    return refine_small_components(mask_in, mask_out, batch, h, w, area_threshold,
                                   reinterpret_cast<cudaStream_t>(stream));
}

uint64_t ucod_lt_boxes_workspace_bytes(int batch, int h, int w) {
    return (uint64_t)lt_boxes_workspace_bytes(batch, h, w);
}
int ucod_lt_boxes(const uint8_t* mask, int batch, int h, int w, double look_twice_th, int expand_dynamic,
                  double const_scale, int32_t* boxes, int32_t* nbox, int32_t* status, int32_t* labels,
                  void* workspace, uint64_t workspace_bytes, void* stream) {
    return lt_boxes(mask, batch, h, w, look_twice_th, expand_dynamic, const_scale, boxes, nbox, status, labels,
                    workspace, (size_t)workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_lt_boxes_ex(const uint8_t* mask, int batch, int h, int w, double look_twice_th, int expand_dynamic,
                     double const_scale, int32_t* boxes, int32_t* nbox, int32_t* status, int32_t* labels,
                     void* workspace, uint64_t workspace_bytes, int algorithm, void* stream) {
    return lt_boxes(mask, batch, h, w, look_twice_th, expand_dynamic, const_scale, boxes, nbox, status, labels,
                    workspace, (size_t)workspace_bytes, reinterpret_cast<cudaStream_t>(stream), algorithm);
}
uint64_t ucod_roi_crop_resize_workspace_bytes(int njobs, int max_crop_h, int out_h, int out_w) {
    return (uint64_t)roi_crop_resize_workspace_bytes(njobs, max_crop_h, out_h, out_w);
}
int ucod_roi_crop_resize(const uint8_t* images, int n_images, int src_h, int src_w, int64_t image_stride,
                         int64_t channel_stride, int64_t row_stride, int64_t pixel_stride, const int32_t* jobs,
                         int njobs, int max_crop_h, uint8_t* out, int out_h, int out_w, void* workspace,
                         uint64_t workspace_bytes, int32_t* err_flag, void* stream) {
    return roi_crop_resize(images, n_images, src_h, src_w, image_stride, channel_stride, row_stride, pixel_stride,
                           jobs, njobs, max_crop_h, out, out_h, out_w, workspace, (size_t)workspace_bytes, err_flag,
                           reinterpret_cast<cudaStream_t>(stream));
}
uint64_t ucod_paste_bicubic_workspace_bytes(int njobs, int g_h, int out_cap) {
    return (uint64_t)paste_bicubic_workspace_bytes(njobs, g_h, out_cap);
}
int ucod_paste_bicubic(const float* logits, int njobs, int g_h, int g_w, const int32_t* jobs, int max_rank,
                       uint8_t* mask, int n_images, int s_h, int s_w, int out_cap, void* workspace,
                       uint64_t workspace_bytes, int32_t* err_flag, void* stream) {
    return paste_bicubic(logits, njobs, g_h, g_w, jobs, max_rank, mask, n_images, s_h, s_w, out_cap, workspace,
                         (size_t)workspace_bytes, err_flag, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_lt_build_jobs(const int32_t* boxes, const int32_t* nbox, int batch, int s_h, int s_w, int src_h, int src_w,
                       const int32_t* orig_sizes, int32_t* crop_jobs, int32_t* paste_jobs, int capacity,
                       int32_t* counts, int chunk, int32_t* chunk_counts, void* stream) {
    return lt_build_jobs(boxes, nbox, batch, s_h, s_w, src_h, src_w, orig_sizes, crop_jobs, paste_jobs, capacity, counts,
                         chunk, chunk_counts, reinterpret_cast<cudaStream_t>(stream));
}
uint64_t ucod_roi_crop_resize_dyn_workspace_bytes(int capacity, int src_h, int src_w, int out_h, int out_w) {
    return (uint64_t)roi_crop_resize_dyn_workspace_bytes(capacity, src_h, src_w, out_h, out_w);
}
int ucod_roi_crop_resize_dyn(const uint8_t* images, int n_images, int src_h, int src_w, int64_t image_stride,
                             int64_t channel_stride, int64_t row_stride, int64_t pixel_stride, const int32_t* jobs,
                             int capacity, const int32_t* njobs_dev, uint8_t* out, int out_h, int out_w,
                             void* workspace, uint64_t workspace_bytes, int32_t* err_flag, void* stream) {
    return roi_crop_resize_dyn(images, n_images, src_h, src_w, image_stride, channel_stride, row_stride, pixel_stride,
                               jobs, capacity, njobs_dev, out, out_h, out_w, workspace, (size_t)workspace_bytes,
                               err_flag, reinterpret_cast<cudaStream_t>(stream));
}
uint64_t ucod_paste_bicubic_dyn_workspace_bytes(int capacity, int g_h, int g_w, int out_cap) {
    return (uint64_t)paste_bicubic_dyn_workspace_bytes(capacity, g_h, g_w, out_cap);
}
int ucod_paste_bicubic_dyn(const float* logits, int capacity, const int32_t* njobs_dev, int g_h, int g_w,
                           const int32_t* all_jobs, int first_index, int n_all, const int32_t* n_all_dev,
                           uint8_t* mask, int n_images, int s_h, int s_w, int out_cap, void* workspace,
                           uint64_t workspace_bytes, int32_t* err_flag, void* stream) {
    return paste_bicubic_dyn(logits, capacity, njobs_dev, g_h, g_w, all_jobs, first_index, n_all, n_all_dev, mask,
                             n_images, s_h, s_w, out_cap, workspace, (size_t)workspace_bytes, err_flag,
                             reinterpret_cast<cudaStream_t>(stream));
}
int ucod_mask_scale_u8(const uint8_t* in, uint8_t* out, uint64_t n, int mul, void* stream) {
    return mask_scale_u8(in, out, (size_t)n, mul, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_to_tensor_normalize(const uint8_t* in, float* out, uint64_t planes, int hw, int channels,
                             const float* mean, const float* stddev, void* stream) {
    return to_tensor_normalize(in, out, (size_t)planes, hw, channels, mean, stddev,
                               reinterpret_cast<cudaStream_t>(stream));
}

uint64_t ucod_discriminator_workspace_bytes(int batch, int fs) {
    return (uint64_t)discriminator_workspace_bytes(batch, fs);
}
int ucod_discriminator_fwd(const float* mask, int batch, int fs, const ucod_disc_weights* w, int bn_train,
                           int update_running, float* prob, void* workspace, uint64_t workspace_bytes, void* stream) {
    UCOD_REQUIRE(w != nullptr, "ucod_discriminator_fwd: null weights");
    DiscWeights d{w->conv1, w->bn1_w, w->bn1_b, w->bn1_mean, w->bn1_var, w->conv2, w->bn2_w, w->bn2_b, w->bn2_mean,
                  w->bn2_var, w->conv3, w->bn3_w, w->bn3_b, w->bn3_mean, w->bn3_var, w->lin_w, w->lin_b};
    return discriminator_forward(mask, batch, fs, d, bn_train, update_running, prob, workspace,
                                 (size_t)workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
uint64_t ucod_discriminator_workspace_bytes_calls(int batch, int fs, int calls) {
    return (uint64_t)discriminator_workspace_bytes_groups(batch, fs, calls);
}
int ucod_discriminator_fwd_calls(const float* masks, int batch, int calls, int fs, const ucod_disc_weights* w,
                                 int bn_train, int update_running, float* prob, void* workspace,
                                 uint64_t workspace_bytes, void* stream) {
    UCOD_REQUIRE(w != nullptr, "ucod_discriminator_fwd_calls: null weights");
    DiscWeights d{w->conv1, w->bn1_w, w->bn1_b, w->bn1_mean, w->bn1_var, w->conv2, w->bn2_w, w->bn2_b, w->bn2_mean,
                  w->bn2_var, w->conv3, w->bn3_w, w->bn3_b, w->bn3_mean, w->bn3_var, w->lin_w, w->lin_b};
    return discriminator_forward(masks, batch, fs, d, bn_train, update_running, prob, workspace,
                                 (size_t)workspace_bytes, reinterpret_cast<cudaStream_t>(stream), calls);
}
int ucod_apm_binarize(const float* student, const float* teacher, const float* pl, float* s_mask, float* t_mask,
                      float* p_mask, uint64_t n, void* stream) {
    return apm_binarize(student, teacher, pl, s_mask, t_mask, p_mask, (size_t)n,
                        reinterpret_cast<cudaStream_t>(stream));
}
int ucod_apm_merge(const float* pl, const float* t_mask, const float* p_s, const float* p_p, float epoch_term,
                   float* merged, float* weight, float* dis_loss, int batch, int pixels, void* stream) {
    return apm_merge(pl, t_mask, p_s, p_p, epoch_term, merged, weight, dis_loss, batch, pixels,
                     reinterpret_cast<cudaStream_t>(stream));
}

// ---- first-stage training ----
uint64_t ucod_decoder_bwd_workspace_bytes(int batch, int gin_h, int gin_w, int out_h, int out_w) {
    return (uint64_t)decoder_backward_workspace_bytes(batch, gin_h, gin_w, out_h, out_w);
}
int ucod_decoder_bwd(const void* keys_bf16, int batch, int dim, int gin_h, int gin_w, int out_h, int out_w,
                     const void* w_dec, const float* b_dec, const float* emb, const float* w_fg, const float* b_fg,
                     const float* w_bg, const float* b_bg, const float* fg, const float* bg, const float* target,
                     const float* dfg, const float* dbg, const float* dortho, void* fwd_workspace,
                     uint64_t fwd_workspace_bytes, float* g_w_dec, float* g_b_dec, float* g_w_fg,
                     float* g_b_fg, float* g_w_bg, float* g_b_bg, float* loss2, void* workspace,
                     uint64_t workspace_bytes, void* stream) {
    UCOD_REQUIRE(w_dec && b_dec && emb && w_fg && b_fg && w_bg && b_bg, "ucod_decoder_bwd: null weight pointer");
    DecoderWeights w{dim, w_dec, b_dec, emb, w_fg, b_fg, w_bg, b_bg};
    DecoderGrads g{g_w_dec, g_b_dec, g_w_fg, g_b_fg, g_w_bg, g_b_bg};
    return decoder_backward(keys_bf16, batch, gin_h, gin_w, out_h, out_w, w, fg, bg, target, dfg, dbg, dortho,
                            fwd_workspace, (size_t)fwd_workspace_bytes, g, loss2, workspace, (size_t)workspace_bytes,
                            reinterpret_cast<cudaStream_t>(stream));
}
int ucod_train_loss(const float* loss2, const float* ortho, const float* dis_loss, float* out, void* stream) {
    return train_loss(loss2, ortho, dis_loss, out, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_adamw_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema, uint64_t n,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                        float ema_alpha, void* stream) {
    return adamw_ema_step(params, grads, exp_avg, exp_avg_sq, ema, (size_t)n, lr, beta1, beta2, eps, weight_decay, step,
                          grad_scale, ema_alpha, reinterpret_cast<cudaStream_t>(stream));
}

uint64_t ucod_discriminator_bwd_workspace_bytes(int batch, int fs) {
    return (uint64_t)discriminator_backward_workspace_bytes(batch, fs);
}
int ucod_discriminator_bwd(const float* mask, int batch, int fs, const ucod_disc_weights* w, const float* prob,
                           float label, int n_total, const ucod_disc_grads* g, float* loss, void* fwd_workspace,
                           void* workspace, uint64_t workspace_bytes, void* stream) {
    UCOD_REQUIRE(w != nullptr && g != nullptr, "ucod_discriminator_bwd: null weights / grads");
    DiscWeights d{w->conv1, w->bn1_w, w->bn1_b, w->bn1_mean, w->bn1_var, w->conv2, w->bn2_w, w->bn2_b, w->bn2_mean,
                  w->bn2_var, w->conv3, w->bn3_w, w->bn3_b, w->bn3_mean, w->bn3_var, w->lin_w, w->lin_b};
    DiscGrads gg{g->conv1, g->bn1_w, g->bn1_b, g->conv2, g->bn2_w, g->bn2_b, g->conv3, g->bn3_w, g->bn3_b, g->lin_w,
                 g->lin_b};
    return discriminator_backward(mask, batch, fs, d, prob, label, n_total, gg, loss, fwd_workspace, workspace,
                                  (size_t)workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

// ---- CORAL second stage ----
int ucod_coral_entropy_select(const float* preds, int batch, int size, int window_size, float threshold,
                              float* entropy, float* scores, uint8_t* mask, void* scratch, void* stream) {
    return coral_entropy_select(preds, batch, size, window_size, threshold, entropy, scores, mask,
                                static_cast<int*>(scratch), reinterpret_cast<cudaStream_t>(stream));
}
int ucod_coral_entropy_select_ex(const float* preds, int batch, int size, int window_size, float threshold,
                                 float* entropy, float* scores, uint8_t* mask, void* scratch, int per_image,
                                 void* stream) {
    return coral_entropy_select(preds, batch, size, window_size, threshold, entropy, scores, mask,
                                static_cast<int*>(scratch), reinterpret_cast<cudaStream_t>(stream), per_image);
}
int ucod_coral_window_head(const float* taps, int ld_taps, int n_windows, int grid, float bias_const, float* out,
                           void* stream) {
    return coral_window_head(taps, ld_taps, n_windows, grid, bias_const, out, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_coral_scatter_windows(const float* window_preds, const int32_t* slot_of_cell, int batch, int window_size,
                               int grid, float* out, void* stream) {
    return coral_scatter_windows(window_preds, slot_of_cell, batch, window_size, grid, out,
                                 reinterpret_cast<cudaStream_t>(stream));
}
uint64_t ucod_coral_gated_ensemble_workspace_bytes(int batch, int size) {
    return (uint64_t)coral_gated_ensemble_workspace_bytes(batch, size);
}
int ucod_coral_gated_ensemble(const float* preds, int preds_size, const float* h_preds, int batch, int size,
                              int max_per_image, const float* w0, const float* b0, const float* w2, const float* b2,
                              float* out, float* weight, void* workspace, uint64_t workspace_bytes, void* stream) {
    return coral_gated_ensemble(preds, preds_size, h_preds, batch, size, max_per_image, w0, b0, w2, b2, out, weight,
                                workspace, (size_t)workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_layernorm_bf16(const float* x, const float* weight, const float* bias, void* y, int rows, int dim, float eps,
                        void* stream) {
    return layernorm_rows_bf16(x, weight, bias, y, rows, dim, eps, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_cast_f32_bf16(const float* in, void* out, uint64_t n, void* stream) {
    return cast_f32_to_bf16(in, out, (size_t)n, reinterpret_cast<cudaStream_t>(stream));
}
int ucod_features_to_tokens_f32(const float* in, float* out, int batch, int channels, int pixels, int64_t sb,
                                int64_t sc, int64_t sp, void* stream) {
    return features_to_tokens_f32(in, out, batch, channels, pixels, sb, sc, sp,
                                  reinterpret_cast<cudaStream_t>(stream));
}
int ucod_resize_tokens_bilinear(const float* in, float* out_f32, void* out_bf16, int n, int gin_h, int gin_w,
                                int gout_h, int gout_w, int channels, void* stream) {
    return resize_tokens_bilinear(in, out_f32, out_bf16, n, gin_h, gin_w, gout_h, gout_w, channels,
                                  reinterpret_cast<cudaStream_t>(stream));
}

// ---- COD metric suite ----
uint64_t ucod_cod_metrics_workspace_bytes(int batch, int h, int w) {
    return (uint64_t)cod_metrics_workspace_bytes(batch, h, w);
}
int ucod_cod_metrics(const float* gt, const float* pred, int batch, int h, int w, double* out, void* workspace,
                     uint64_t workspace_bytes, void* stream) {
    return cod_metrics(gt, pred, batch, h, w, out, workspace, (size_t)workspace_bytes,
                       reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
