// Pseudo-label scoring + small-component cleanup (see pseudo_label.cu).
#pragma once
#include "common.cuh"

namespace ucod {

int pseudo_label_score(const float* attn_cls, const void* keys, int keys_bf16, int B, int nh, int P, float th_bkg,
                       float epsilon, float* cos_out, uint8_t* bkg_out, int* ref_out, float* sim_out, int* scratch,
                       cudaStream_t stream, int apply_weights = 1, size_t scratch_bytes = 4);
// scratch of at least this size selects the two-launch path (prologue for all images, then a pure streaming kernel)
size_t pseudo_label_scratch_bytes(int B, int nh);
int refine_small_components(const uint8_t* mask_in, uint8_t* mask_out, int B, int H, int W, int area_threshold,
                            cudaStream_t stream);

}  // namespace ucod
