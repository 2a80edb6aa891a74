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

#define UCOD_B200_ABI_VERSION 5

/* Last error message of the calling thread ("" if none). */
const char* ucod_last_error(void);
int ucod_abi_version(void);

/* Launch accounting (bench.py): total kernels launched by this library in the process, and optional
 * per-kernel-class CUDA-event timing.  Classes: 0 gemm, 1 attention, 2 layernorm, 3 embed, 4 decoder,
 * 5 resample, 6 pseudo-label, 7 ccl/boxes, 8 other (arrays of UCOD_KERNEL_CLASSES entries).
 * `work` is FLOPs for classes 0-1 and algorithmic bytes for the rest. */
#define UCOD_KERNEL_CLASSES 9
void ucod_prof_enable(int on);
int ucod_prof_collect(double* ms, double* work, long long* launches);
long long ucod_launch_count(void);

/* ---- dense building block -------------------------------------------------------------------
 * out[M,N] = epilogue(A[M,K] * W[N,K]^T) ; A, W bf16 row-major (K contiguous), fp32 accumulate (tcgen05).
 * epi_mode: 0 = bf16 out, +bias ; 1 = bf16 out, erf-gelu(+bias) ; 2 = fp32 in-place residual
 *           out += acc+bias (TMA reduce-add) ; 5 = fp32 out, +bias.   bias may be NULL.  ld_out in elements.
 * Replaces: torch.nn.Linear / 1x1 Conv2d library GEMMs on the path (HF modeling_dinov2.py:153-235,348-387;
 * models/modules/DBA.py:13,35). */
int ucod_gemm_bf16(const void* a, int lda, const void* w, int ldw, int m, int n, int k, int epi_mode,
                   const float* bias, void* out, int ld_out, void* stream);

/* Fused softmax(q k^T * scale) v, non-causal (tcgen05 flash-attention forward), head_dim 64 or 128.
 * q: [batch, tokens_q, ld_q] bf16 ; k, v: [batch, tokens_kv, ld_kv] bf16 ; head h lives in columns
 * [h*head_dim, (h+1)*head_dim) of each pointer, so q/k/v may point at the three column blocks of one fused QKV
 * projection output (nothing is transposed in memory).  ctx: [batch, tokens_q, ld_ctx] bf16, same head layout.
 * Row pitches are in elements (multiples of 8), pointers 16-byte aligned.
 * Replaces: HF Dinov2SelfAttention/ViTSelfAttention (modeling_dinov2.py:153-235) under
 * data/utils/feature_extractor.py:49-59, and nn.MultiheadAttention of models/modules/mlp.py:134-148
 * (head_dim 96 zero-padded to 128). */
int ucod_attention(const void* q, int ld_q, const void* k, const void* v, int ld_kv, void* ctx, int ld_ctx, int batch,
                   int heads, int head_dim, int tokens_q, int tokens_kv, float scale, void* stream);

/* Same kernel with K/V shared between q batches: q batch b attends to K/V batch kv_batch_map[b] (device int32
 * [batch], values < kv_batch).  head_dim_real <= head_dim documents zero-padded heads (FLOP accounting only).
 * Used by the CORAL CSF block, where every selected window of an image attends to that image's low-res tokens
 * (`repeat_interleave` of models/modules/ASR.py:24 without materialising the copies). */
int ucod_attention_shared_kv(const void* q, int ld_q, const void* k, const void* v, int ld_kv, void* ctx, int ld_ctx,
                             int batch, int heads, int head_dim, int head_dim_real, int tokens_q, int tokens_kv,
                             float scale, const int32_t* kv_batch_map, int kv_batch, void* stream);

/* ---- frozen ViT-B backbone: last-layer key tokens ---------------------------------------------
 * Replaces `backbone.__init__/forward` (data/utils/feature_extractor.py:31-59) and the hook + attentions of
 * generate_pseudo_label.py:24-27,76-81,111-112.  Weights are device pointers owned by the caller
 * (bf16 matrices row-major [out,in]; fp32 vectors); the handle only records them. */
typedef struct ucod_vit_cfg {
    int hidden;      /* 768 */
    int layers;      /* 12 */
    int heads;       /* 12 (head_dim 64) */
    int mlp_dim;     /* 3072 */
    int patch;       /* 14 (DINOv2) or 8 (DINO ViT-B/8) */
    int patch_kpad;  /* 3*patch*patch rounded up to a multiple of 64 (row pitch of patch_w) */
    float ln_eps;    /* 1e-6 (DINOv2) / 1e-12 (HF ViT) */
} ucod_vit_cfg;

typedef struct ucod_vit_layer {
    const float* ln1_w; const float* ln1_b;
    const void* w_qkv;  const float* b_qkv;   /* bf16 [3*hidden, hidden] = [Wq;Wk;Wv], fp32 [3*hidden] */
    const void* w_o;    const float* b_o;     /* bf16 [hidden, hidden]; LayerScale lambda1 folded in (rows, bias) */
    const float* ln2_w; const float* ln2_b;
    const void* w_fc1;  const float* b_fc1;   /* bf16 [mlp_dim, hidden] */
    const void* w_fc2;  const float* b_fc2;   /* bf16 [hidden, mlp_dim]; LayerScale lambda2 folded in */
} ucod_vit_layer;

/* patch_w: bf16 [hidden, patch_kpad] (conv weight flattened c-major, zero padded), patch_b fp32 [hidden],
 * cls_token fp32 [hidden], layers: host array of cfg->layers entries (copied). */
int ucod_vit_create(void** handle, const ucod_vit_cfg* cfg, const void* patch_w, const float* patch_b,
                    const float* cls_token, const ucod_vit_layer* layers);
int ucod_vit_destroy(void* handle);
int ucod_vit_workspace_bytes(void* handle, int batch, int img_h, int img_w, uint64_t* bytes);
/* images: [batch,3,img_h,img_w] NCHW; image_dtype 0 = fp32 already normalised (reference transform output),
 * 1 = uint8 raw RGB (ToTensor+Normalize fused into the loader, transforms.py:14-18).
 * pos_emb: fp32 [1+P, hidden] position embedding for this resolution (row 0 = CLS).
 * Outputs (any may be NULL, at least one required):
 *   keys_f32  [batch, P (+1 if keep_cls), hidden] fp32  — `self.key` of feature_extractor.py:46-58 (token-major)
 *   keys_bf16 same shape, bf16 (feeds the decoder GEMM without a conversion pass)
 *   cls_attn  [batch, heads, P] fp32 — attentions[-1][:, :, 0, 1:] (found_bkg_mask.py:24)
 * workspace: device scratch of >= ucod_vit_workspace_bytes(), 1 KiB aligned. */
int ucod_vit_keys(void* handle, const void* images, int image_dtype, int batch, int img_h, int img_w,
                  const float* pos_emb, void* workspace, uint64_t workspace_bytes, float* keys_f32, void* keys_bf16,
                  float* cls_attn, int keep_cls, void* stream);

/* Same call with a DEVICE-side image count: `batch` is the capacity the buffers are sized for, *batch_dev (device
 * int32, <= batch) the number of leading images that are processed; every kernel of the pass reads it on the device,
 * so a pass whose size was decided by an earlier kernel (the second Look-Twice pass over the crops,
 * engine/runner/loop_UCOD_DPL.py:331-346) needs no host synchronisation.  Rows of images >= *batch_dev are left
 * untouched or undefined. */
int ucod_vit_keys_dyn(void* handle, const void* images, int image_dtype, int batch, const int32_t* batch_dev, int img_h,
                      int img_w, const float* pos_emb, void* workspace, uint64_t workspace_bytes, float* keys_f32,
                      void* keys_bf16, float* cls_attn, int keep_cls, void* stream);

/* ---- Dual-Branch Adversarial decoder ------------------------------------------------------------
 * Replaces `RevDecoder.forward` (models/modules/DBA.py:31-59) incl. the preceding bilinear feature upsample
 * of engine/runner/loop_UCOD_DPL.py:153,236,305 (commuted past the 1x1 conv) and `calc_orthogonal_loss`
 * (DBA.py:25-29, via the Gram identity).
 * keys: bf16 token-major [batch, gin_h*gin_w, dim]; logits are produced on the (out_h, out_w) grid.
 * Weights are device pointers: w_dec bf16 [128,dim]; b_dec[128], emb[2*64], w_fg[64], b_fg[1], w_bg[64], b_bg[1] fp32.
 * fg [batch,out_h*out_w] (required); bg same shape or NULL; ortho: device scalar or NULL (student only). */
uint64_t ucod_decoder_workspace_bytes(int batch, int gin_h, int gin_w, int out_h, int out_w, int want_ortho);
int ucod_decoder_fwd(const void* keys_bf16, int batch, int dim, int gin_h, int gin_w, int out_h, int out_w,
                     const void* w_dec, const float* b_dec, const float* emb, const float* w_fg, const float* b_fg,
                     const float* w_bg, const float* b_bg, float* fg, float* bg, float* ortho, void* workspace,
                     uint64_t workspace_bytes, void* stream);

/* Eval-only variant with a device-side image count (see ucod_vit_keys_dyn); no orthogonality loss. */
int ucod_decoder_fwd_dyn(const void* keys_bf16, int batch, const int32_t* batch_dev, int dim, int gin_h, int gin_w,
                         int out_h, int out_w, const void* w_dec, const float* b_dec, const float* emb,
                         const float* w_fg, const float* b_fg, const float* w_bg, const float* b_bg, float* fg,
                         float* bg, void* workspace, uint64_t workspace_bytes, void* stream);

/* [batch, channels, pixels] fp32 with element strides (sb, sc, sp) -> token-major bf16 [batch, pixels, channels].
 * Lets `baseline.forward` accept the reference's NCHW feature tensors (models/uscod.py:16-22). */
int ucod_features_to_tokens_bf16(const float* in, void* out, int batch, int channels, int pixels, int64_t sb,
                                 int64_t sc, int64_t sp, void* stream);

/* F.interpolate(mode='bilinear', align_corners=False) of [batch,in_h,in_w] fp32.  binarize = 0: fp32 output;
 * binarize = 1: uint8 {0,1} mask of `sigmoid(up(x)) > 0.5` (engine/runner/loop_UCOD_DPL.py:356-361);
 * binarize = 2: mask of `up(sigmoid(x)) > 0.5`, binarize = 3: mask of `up(x) > 0.5` for inputs that already are
 * probabilities (the two branches of engine/runner/loop_CORAL.py:331-340). */
int ucod_upsample_bilinear(const float* in, void* out, int batch, int in_h, int in_w, int out_h, int out_w,
                           int binarize, void* stream);

/* ---- fixed-strategy pseudo-labels ----------------------------------------------------------------
 * `compute_img_bkg_seg` (data/utils/found_bkg_mask.py:4-85) for the CLS attention row and patch keys:
 * attn_cls fp32 [batch, heads, patches] (= attentions[-1][:, :, 0, 1:]), keys [batch, patches, heads*64]
 * (fp32, or bf16 when keys_bf16 != 0).  Outputs: cos fp32 [batch,patches] (cosine to the least-attended patch),
 * bkg uint8 [batch,patches] (cos > th_bkg), ref_idx int32 [batch]; sim fp32 [batch,patches] or NULL
 * ((1-cos)/max(1-cos) * (1-bkg), the max taken over the whole call as in the reference, :81-85);
 * scratch: 4 device bytes (required when sim != NULL). */
int ucod_pseudo_label_score(const float* attn_cls, const void* keys, int keys_bf16, int batch, int heads, int patches,
                            float th_bkg, float epsilon, float* cos, uint8_t* bkg, int32_t* ref_idx, float* sim,
                            void* scratch, void* stream);
/* Same with the reference's `apply_weights` switch (found_bkg_mask.py:44-47,64-65): 0 = neither the descriptors nor the
 * per-patch attention sum are weighted by the head sparsity weights beta.  A scratch of
 * ucod_pseudo_label_scratch_bytes(batch, heads) (256-byte aligned) selects the two-launch path: the weights /
 * reference-patch prologue for all images first, then a streaming kernel over (image, slab of patches); with the
 * 4-byte scratch the single-launch kernel runs.  Results are identical. */
uint64_t ucod_pseudo_label_scratch_bytes(int batch, int heads);
int ucod_pseudo_label_score_ex(const float* attn_cls, const void* keys, int keys_bf16, int batch, int heads, int patches,
                               float th_bkg, float epsilon, int apply_weights, float* cos, uint8_t* bkg,
                               int32_t* ref_idx, float* sim, void* scratch, uint64_t scratch_bytes, void* stream);
/* `refine_post_process` (generate_pseudo_label.py:30-67): flip 8-connected foreground components with
 * area < area_threshold whose 1-px ring is entirely the opposite label; OpenCV label order; h*w <= 1024.
 * mask_in/mask_out: uint8 {0,1} [batch, h, w] (may alias). */
int ucod_refine_small_components(const uint8_t* mask_in, uint8_t* mask_out, int batch, int h, int w,
                                 int area_threshold, void* stream);

/* ---- Look-Twice -----------------------------------------------------------------------------------
 * `ValLoop_Look_Twice.process_preds` box logic (engine/runner/loop_UCOD_DPL.py:366-384) + `expand_bbox`
 * (:399-417) for a batch of binarised masks [batch,h,w] uint8 (non-zero = foreground):
 * 8-connected components (cv2.connectedComponents order), area fractions, boundingRect, dynamic/const
 * expansion in exact fp64, stable sort by -w*h.
 * boxes: int32 [batch, UCOD_LT_MAX_BOXES, 4] (x,y,w,h); nbox: int32 [batch] — >= 0 number of boxes,
 * -1 = "None" (largest component fraction >= look_twice_th, no second look), -2 = the reference would raise
 * ValueError (sqrt of a negative scale), -3 = labeller capacity exceeded (see ucod_lt_boxes_ex);
 * status: int32 [batch] bit0 = sqrt domain, bit1 = box table overflow, bit2 = labeller capacity.
 * labels (optional, int32 [batch,h,w]): per-pixel component root (min raster index of the component, -1 = bg). */
#define UCOD_LT_MAX_BOXES 128
uint64_t ucod_lt_boxes_workspace_bytes(int batch, int h, int w);
int ucod_lt_boxes(const uint8_t* mask, int batch, int h, int w, double look_twice_th, int expand_dynamic,
                  double const_scale, int32_t* boxes, int32_t* nbox, int32_t* status, int32_t* labels,
                  void* workspace, uint64_t workspace_bytes, void* stream);

/* Same with the labelling algorithm chosen by the caller.  algorithm 0 = automatic (what ucod_lt_boxes does): masks of
 * up to 1024 rows are labelled by ONE CTA each in shared memory — bit image, runs of foreground pixels, union-find over
 * runs; no label image ever reaches HBM, the only sizeable traffic is the mask itself.  A mask with more runs or
 * components than shared memory holds (~23 000 runs at 518^2; masks up-sampled from a 68 x 68 logit map have at most
 * 18 130) gets nbox = -3 and status bit 2 (value 4); re-run such a batch with algorithm 1 = union-find over pixels in
 * global memory (any size).  algorithm 2 = shared-memory labeller or an error when the geometry does not fit. */
int ucod_lt_boxes_ex(const uint8_t* mask, int batch, int h, int w, double look_twice_th, int expand_dynamic,
                     double const_scale, int32_t* boxes, int32_t* nbox, int32_t* status, int32_t* labels,
                     void* workspace, uint64_t workspace_bytes, int algorithm, void* stream);

/* PIL `crop` + torchvision `Resize((out_h,out_w))` (Pillow antialiased BILINEAR, bit-exact fixed-point two-pass
 * resample) of ROIs of uint8 RGB images (loop_UCOD_DPL.py:335-342).  images: uint8, element strides given
 * (planar CHW or interleaved HWC both work); jobs: device int32 [njobs,5] = (image index, x, y, w, h) in source
 * pixels (out-of-image area reads as 0, like PIL); max_crop_h >= max job h; out: uint8 planar [njobs,3,out_h,out_w];
 * err_flag: device int32, bit0 set if a scale exceeds the supported tap count. */
uint64_t ucod_roi_crop_resize_workspace_bytes(int njobs, int max_crop_h, int out_h, int out_w);
int ucod_roi_crop_resize(const uint8_t* images, int n_images, int src_h, int src_w, int64_t image_stride,
                         int64_t channel_stride, int64_t row_stride, int64_t pixel_stride, const int32_t* jobs,
                         int njobs, int max_crop_h, uint8_t* out, int out_h, int out_w, void* workspace,
                         uint64_t workspace_bytes, int32_t* err_flag, void* stream);

/* Second-look paste (loop_UCOD_DPL.py:348-351): binarise logits [njobs,g_h,g_w] (sigmoid > 0.5), Pillow default
 * BICUBIC resize of the {0,255} map to (w,h), paste at (x,y) into mask uint8 [n_images,s_h,s_w] (values 0..255).
 * jobs: device int32 [njobs,6] = (image index, x, y, w, h, rank); jobs of one image are applied in rank order
 * (0..max_rank). out_cap >= max(w,h) over jobs. */
uint64_t ucod_paste_bicubic_workspace_bytes(int njobs, int g_h, int out_cap);
int ucod_paste_bicubic(const float* logits, int njobs, int g_h, int g_w, const int32_t* jobs, int max_rank,
                       uint8_t* mask, int n_images, int s_h, int s_w, int out_cap, void* workspace,
                       uint64_t workspace_bytes, int32_t* err_flag, void* stream);
/* Look-Twice job tables built ON THE DEVICE from ucod_lt_boxes output — the host loop of
 * engine/runner/loop_UCOD_DPL.py:331-342 and `resize_bbox` (:387-397, CPython float arithmetic + int() truncation in
 * exact fp64).  Image b contributes max(nbox[b], 0) jobs, image-major, rank = position in its sorted box list.
 * orig_sizes: device int32 [batch,2] (h, w) of every original image, or NULL when all are src_h x src_w.
 * crop_jobs int32 [capacity,5] (image, x, y, w, h in original pixels); paste_jobs int32 [capacity,6]
 * (image, x, y, w, h, rank in mask pixels); counts int32[4]: [0] jobs written (<= capacity), [1] status bits
 * (1: an image's box maths raised ValueError, 2: more than `capacity` jobs — the rest dropped, 4: a box or crop with
 * w or h <= 0, where PIL raises in the reference), [2] jobs requested; chunk_counts int32 [ceil(capacity / chunk)]:
 * valid jobs of every `chunk`-sized slice, the device-side counts of the *_dyn calls that process the slices. */
int ucod_lt_build_jobs(const int32_t* boxes, const int32_t* nbox, int batch, int s_h, int s_w, int src_h, int src_w,
                       const int32_t* orig_sizes, int32_t* crop_jobs, int32_t* paste_jobs, int capacity,
                       int32_t* counts, int chunk, int32_t* chunk_counts, void* stream);
/* ucod_roi_crop_resize for a job slice whose length lives on the device (*njobs_dev <= capacity).  Any crop of the
 * src_h x src_w sources is supported (no tap-count limit); err_flag as above. */
uint64_t ucod_roi_crop_resize_dyn_workspace_bytes(int capacity, int src_h, int src_w, int out_h, int out_w);
int ucod_roi_crop_resize_dyn(const uint8_t* images, int n_images, int src_h, int src_w, int64_t image_stride,
                             int64_t channel_stride, int64_t row_stride, int64_t pixel_stride, const int32_t* jobs,
                             int capacity, const int32_t* njobs_dev, uint8_t* out, int out_h, int out_w,
                             void* workspace, uint64_t workspace_bytes, int32_t* err_flag, void* stream);
/* ucod_paste_bicubic for the slice [first_index, first_index + *njobs_dev) of an image-major, rank-ascending paste
 * table of min(n_all, *n_all_dev) entries.  A pixel is written by the LAST job of its image that covers it (what the
 * reference's sequential pastes leave behind), so slices can be pasted in one launch each and in any order.
 * Jobs wider or taller than out_cap set err_flag bit 1 and are skipped. */
uint64_t ucod_paste_bicubic_dyn_workspace_bytes(int capacity, int g_h, int g_w, int out_cap);
int ucod_paste_bicubic_dyn(const float* logits, int capacity, const int32_t* njobs_dev, int g_h, int g_w,
                           const int32_t* all_jobs, int first_index, int n_all, const int32_t* n_all_dev,
                           uint8_t* mask, int n_images, int s_h, int s_w, int out_cap, void* workspace,
                           uint64_t workspace_bytes, int32_t* err_flag, void* stream);
/* out[i] = in[i] ? mul : 0  ({0,1} mask -> {0,255} canvas, loop_UCOD_DPL.py:330). */
int ucod_mask_scale_u8(const uint8_t* in, uint8_t* out, uint64_t n, int mul, void* stream);

/* torchvision `ToTensor` + `Normalize(mean, std)` of planar uint8 images (data/datasets/transforms.py:14-18,
 * 31-35, 40-43): out = ((in / 255) - mean[c]) / std[c], every step one correctly rounded fp32 operation as in
 * the torch ops it replaces (bit-exact).  in: uint8 [planes, hw] (4-byte aligned), plane p has channel p % channels;
 * mean / stddev: HOST float[channels]; channels == 0: ToTensor only (label transform, transforms.py:23-26). */
int ucod_to_tensor_normalize(const uint8_t* in, float* out, uint64_t planes, int hw, int channels,
                             const float* mean, const float* stddev, void* stream);

/* ---- APM: discriminator + pseudo-label fusion -----------------------------------------------------
 * `Discriminator.forward` (models/discriminator.py:86-95, dis_use_features = False): mask fp32 [batch,1,fs,fs]
 * -> prob fp32 [batch].  Weight pointers follow the reference state_dict: maskConv.layers.{0,1}, convs.{0,1}.layers.{0,1},
 * linear.  bn_train != 0: BatchNorm uses batch statistics (the reference never calls .eval() on it) and, if
 * update_running != 0, also updates running_mean/var (momentum 0.1, unbiased variance) like nn.BatchNorm2d. */
typedef struct ucod_disc_weights {
    const float* conv1; const float* bn1_w; const float* bn1_b; float* bn1_mean; float* bn1_var;
    const float* conv2; const float* bn2_w; const float* bn2_b; float* bn2_mean; float* bn2_var;
    const float* conv3; const float* bn3_w; const float* bn3_b; float* bn3_mean; float* bn3_var;
    const float* lin_w; const float* lin_b;
} ucod_disc_weights;
uint64_t ucod_discriminator_workspace_bytes(int batch, int fs);
int ucod_discriminator_fwd(const float* mask, int batch, int fs, const ucod_disc_weights* w, int bn_train,
                           int update_running, float* prob, void* workspace, uint64_t workspace_bytes, void* stream);
/* Several consecutive `Discriminator.forward` calls of the same batch size in one set of launches: masks
 * [calls*batch, 1, fs, fs] (call-major), prob [calls*batch].  Each call keeps its own BatchNorm batch statistics and the
 * running buffers are updated in call order, i.e. exactly what `merge_pseudo_label`'s two calls
 * (engine/runner/loop_UCOD_DPL.py:262-263) produce one after the other. */
uint64_t ucod_discriminator_workspace_bytes_calls(int batch, int fs, int calls);
int ucod_discriminator_fwd_calls(const float* masks, int batch, int calls, int fs, const ucod_disc_weights* w,
                                 int bn_train, int update_running, float* prob, void* workspace,
                                 uint64_t workspace_bytes, void* stream);
/* `TrainLoop.merge_pseudo_label` (engine/runner/loop_UCOD_DPL.py:257-272), split around the two discriminator calls:
 * binarize: s_mask = sigmoid(student) > 0.5, t_mask = sigmoid(teacher) > 0.5, p_mask = pl > 0.5 (fp32 {0,1}, n elements)
 * merge:    w_b = clamp(0.5*(1+cos(pi*|p_s-p_p|)) + epoch_term, 0, 1); merged = pl*(1-w_b) + t_mask*w_b;
 *           weight[b] = w_b; dis_loss (optional device scalar) = BCE(p_s, 0). */
int ucod_apm_binarize(const float* student, const float* teacher, const float* pl, float* s_mask, float* t_mask,
                      float* p_mask, uint64_t n, void* stream);
int ucod_apm_merge(const float* pl, const float* t_mask, const float* p_s, const float* p_p, float epoch_term,
                   float* merged, float* weight, float* dis_loss, int batch, int pixels, void* stream);

/* ---- first-stage training step ----------------------------------------------------------------------------
 * Backward of the student `RevDecoder` for `TrainLoop._process_batch` (engine/runner/loop_UCOD_DPL.py:148-184):
 * loss = BCEWithLogits(fg, target) + BCEWithLogits(bg, 1 - target) + ortho, target = APM-merged pseudo labels
 * [batch, out_h*out_w] (constant).  Call after ucod_decoder_fwd(... ortho != NULL ...) on the same keys, passing
 * that call's workspace; fg / bg are its outputs.  Gradients (fp32, zeroed here): g_w_dec [128,dim], g_b_dec [128],
 * g_w_fg [64], g_b_fg [1], g_w_bg [64], g_b_bg [1] (all written, none accumulated; every sum runs in a fixed order, so
 * the gradients are bit-reproducible); the gradient of learnable_embedding is identically zero
 * (F.normalize).  loss2: device float[2] = the two BCE means (add the forward's ortho for the total).
 * Autograd entry: pass target = NULL and the upstream gradients dfg, dbg [batch, out_h*out_w] and dortho (device
 * scalar) instead; then no loss is computed. */
uint64_t ucod_decoder_bwd_workspace_bytes(int batch, int gin_h, int gin_w, int out_h, int out_w);
int ucod_decoder_bwd(const void* keys_bf16, int batch, int dim, int gin_h, int gin_w, int out_h, int out_w,
                     const void* w_dec, const float* b_dec, const float* emb, const float* w_fg, const float* b_fg,
                     const float* w_bg, const float* b_bg, const float* fg, const float* bg, const float* target,
                     const float* dfg, const float* dbg, const float* dortho, void* fwd_workspace,
                     uint64_t fwd_workspace_bytes, float* g_w_dec, float* g_b_dec, float* g_w_fg,
                     float* g_b_fg, float* g_w_bg, float* g_b_bg, float* loss2, void* workspace,
                     uint64_t workspace_bytes, void* stream);
/* Total of `_process_batch` (loop_UCOD_DPL.py:176-180): out = loss2[0] + loss2[1] + ortho - dis_loss, all device scalars;
 * dis_loss may be NULL (finetune epochs). */
int ucod_train_loss(const float* loss2, const float* ortho, const float* dis_loss, float* out, void* stream);
/* torch.optim.AdamW step (decoupled weight decay, bias correction with step >= 1) on flat fp32 buffers, fused with
 * `update_ema_decoder` (loop_UCOD_DPL.py:186-191): ema = ema_alpha*ema + (1-ema_alpha)*param (ema may be NULL).
 * grads are multiplied by grad_scale first (1/world_size after the gradient all-reduce). */
int ucod_adamw_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema, uint64_t n,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                        float ema_alpha, void* stream);

/* Backward of one `ucod_discriminator_fwd(bn_train = 1)` call for `TrainLoop.Discriminator_epoch`
 * (engine/runner/loop_UCOD_DPL.py:230-255): loss = BCE(prob, label) averaged over n_total samples (the epoch's two
 * calls — pseudo labels with label 1, student masks with label 0 — use n_total = 2*batch).  fwd_workspace: the
 * workspace of that forward call.  Gradients (fp32, shapes of the weights) and `loss` (device scalar) are ACCUMULATED;
 * zero them before the first call of a step.  Apply them with ucod_adamw_ema_step(ema = NULL). */
typedef struct ucod_disc_grads {
    float* conv1; float* bn1_w; float* bn1_b; float* conv2; float* bn2_w; float* bn2_b;
    float* conv3; float* bn3_w; float* bn3_b; float* lin_w; float* lin_b;
} ucod_disc_grads;
uint64_t ucod_discriminator_bwd_workspace_bytes(int batch, int fs);
int ucod_discriminator_bwd(const float* mask, int batch, int fs, const ucod_disc_weights* w, const float* prob,
                           float label, int n_total, const ucod_disc_grads* g, float* loss, void* fwd_workspace,
                           void* workspace, uint64_t workspace_bytes, void* stream);

/* ---- CORAL second stage (SparseRefiner, eval) ----------------------------------------------------------
 * The dense part of the CSF block (models/modules/CSF.py:38-43, mlp.py:134-148) is assembled by the host from
 * ucod_layernorm_bf16 + ucod_gemm_bf16 + ucod_attention_shared_kv; the functions below are the remaining stages.
 *
 * `EntropySelector.forward` (models/modules/ASR.py:41-51): preds fp32 [batch,size,size] (logits, or probabilities
 * when every value of the call lies in [0,1]); entropy fp32 [batch,size,size]; scores fp32 and mask uint8
 * [batch, window_size^2] (adaptive average pool > threshold); scratch: 4 device bytes. */
int ucod_coral_entropy_select(const float* preds, int batch, int size, int window_size, float threshold,
                              float* entropy, float* scores, uint8_t* mask, void* scratch, void* stream);
/* Same with the probability-or-logit decision taken per image (per_image != 0; scratch: batch * 4 device bytes): the
 * reference evaluates at batch 1, so `torch.all((preds >= 0) & (preds <= 1))` (ASR.py:42) is a per-image test there. */
int ucod_coral_entropy_select_ex(const float* preds, int batch, int size, int window_size, float threshold,
                                 float* entropy, float* scores, uint8_t* mask, void* scratch, int per_image,
                                 void* stream);
/* CSF tail (CSF.py:41-42): depthwise 7x7 (pad 3) + 1x1 mask_dec folded into 49 taps per token.
 * taps fp32 [n_windows*grid*grid, ld_taps] (column ky*7+kx = sum_c mask_dec.w[c]*dw.w[c,ky,kx]*x[token,c]);
 * out fp32 [n_windows, grid, grid] = bias_const + zero-padded 7x7 gather-sum. */
int ucod_coral_window_head(const float* taps, int ld_taps, int n_windows, int grid, float bias_const, float* out,
                           void* stream);
/* `HRE.concate_windows` (models/modules/HRE.py:18-39): slot_of_cell int32 [batch, window_size^2] = index of the
 * cell's window in window_preds [n,grid,grid] or -1; out fp32 [batch, window_size*grid, window_size*grid]. */
int ucod_coral_scatter_windows(const float* window_preds, const int32_t* slot_of_cell, int batch, int window_size,
                               int grid, float* out, void* stream);
/* `GatedEnsembler.forward` (models/modules/GE_pix_level.py:16-26): preds fp32 [batch,preds_size,preds_size] coarse
 * logits, h_preds fp32 [batch,size,size]; fuser weights w0[64], b0[64], w2[64], b2[1];
 * out / weight fp32 [batch,size,size].  `en_local.max()` (GE_pix_level.py:23) is a maximum over the whole tensor:
 * max_per_image = 0 reproduces that for the call's batch; max_per_image = 1 takes it per image, which is what the
 * reference's batch-1 eval loop (loop_CORAL.py:260-311) computes for every image and makes a batched eval
 * independent of the batch composition. */
uint64_t ucod_coral_gated_ensemble_workspace_bytes(int batch, int size);
int ucod_coral_gated_ensemble(const float* preds, int preds_size, const float* h_preds, int batch, int size,
                              int max_per_image, const float* w0, const float* b0, const float* w2, const float* b2,
                              float* out, float* weight, void* workspace, uint64_t workspace_bytes, void* stream);
/* nn.LayerNorm over the last dim (dim % 128 == 0, <= 1024): x fp32 [rows,dim] -> y bf16 [rows,dim]. */
int ucod_layernorm_bf16(const float* x, const float* weight, const float* bias, void* y, int rows, int dim, float eps,
                        void* stream);
int ucod_cast_f32_bf16(const float* in, void* out, uint64_t n, void* stream);
/* NCHW-style [batch, channels, pixels] fp32 (element strides sb, sc, sp) -> token-major fp32 [batch, pixels, channels]. */
int ucod_features_to_tokens_f32(const float* in, float* out, int batch, int channels, int pixels, int64_t sb,
                                int64_t sc, int64_t sp, void* stream);
/* F.interpolate(mode='bilinear') of token-major maps (engine/runner/loop_CORAL.py:224-227, loop_UCOD_DPL.py:305):
 * in fp32 [n, gin_h*gin_w, channels] -> out_f32 and/or out_bf16 [n, gout_h*gout_w, channels] (either may be NULL). */
int ucod_resize_tokens_bilinear(const float* in, float* out_f32, void* out_bf16, int n, int gin_h, int gin_w,
                                int gout_h, int gout_w, int channels, void* stream);

/* ---- COD metric suite (next row after the model path) ------------------------------------------
 * `statistics.step` per image (engine/utils/metrics/metric.py:19-74,128-531) in fp64 on the device:
 * gt, pred fp32 [batch,h,w] (any value range: `_prepare_data` normalisation is applied);
 * out fp64 [batch, UCOD_METRICS_OUT] = acc, iou, mae, s-measure, adaptive E, adaptive F, weighted F,
 * E curve[256], F curve[256] (index t = threshold 255 - t, like the reference's flipped cumulative histograms).
 * Dataset results are the means over images (and max / mean of the mean curves), `statistics.get_result`. */
#define UCOD_METRICS_OUT 519
uint64_t ucod_cod_metrics_workspace_bytes(int batch, int h, int w);
int ucod_cod_metrics(const float* gt, const float* pred, int batch, int h, int w, double* out, void* workspace,
                     uint64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UCOD_B200_H */
