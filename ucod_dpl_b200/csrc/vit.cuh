// ViT key extractor internals (see vit.cu). The config / per-layer weight structs are part of the C-ABI.
#pragma once
#include "../../include/ucod_b200.h"
#include "common.cuh"

namespace ucod {

int vit_create(void** handle, const ucod_vit_cfg* cfg, const void* patch_w, const float* patch_b, const float* cls,
               const ucod_vit_layer* layers);
int vit_destroy(void* handle);
int vit_workspace_bytes(void* handle, int B, int img_h, int img_w, size_t* out);
int vit_keys(void* handle, const void* images, int image_dtype, int B, int img_h, int img_w, const float* pos_emb,
             void* workspace, size_t ws_bytes, float* keys_f32, void* keys_bf16, float* cls_attn, int keep_cls,
             cudaStream_t stream, const int* batch_dev = nullptr);

}  // namespace ucod
