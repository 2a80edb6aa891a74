// Look-Twice stages: CC + boxes, PIL-exact ROI crop/resize, bicubic paste (see looktwice.cu).
#pragma once
#include "common.cuh"

namespace ucod {

constexpr int LT_MAX_BOXES = 128;

size_t lt_boxes_workspace_bytes(int B, int H, int W);
int lt_boxes(const uint8_t* mask, int B, int H, int W, double look_twice_th, int dynamic, double const_scale,
             int* boxes, int* nbox, int* status, int* labels_out, void* workspace, size_t ws_bytes,
             cudaStream_t stream, int algorithm = 0);
size_t roi_crop_resize_workspace_bytes(int njobs, int max_crop_h, int out_h, int out_w);
int roi_crop_resize(const uint8_t* images, int n_img, int H0, int W0, long long img_stride, long long ch_stride,
                    long long row_stride, long long px_stride, const int* jobs, int njobs, int max_crop_h,
                    uint8_t* out, int out_h, int out_w, void* workspace, size_t ws_bytes, int* err_flag,
                    cudaStream_t stream);
size_t paste_bicubic_workspace_bytes(int njobs, int g_h, int out_cap);
int paste_bicubic(const float* logits, int njobs, int g_h, int g_w, const int* jobs, int max_rank, uint8_t* mask,
                  int n_img, int S_h, int S_w, int out_cap, void* workspace, size_t ws_bytes, int* err_flag,
                  cudaStream_t stream);
size_t roi_crop_resize_dyn_workspace_bytes(int capacity, int H0, int W0, int out_h, int out_w);
int roi_crop_resize_dyn(const uint8_t* images, int n_img, int H0, int W0, long long img_stride, long long ch_stride,
                        long long row_stride, long long px_stride, const int* jobs, int capacity, const int* njobs_dev,
                        uint8_t* out, int out_h, int out_w, void* workspace, size_t ws_bytes, int* err_flag,
                        cudaStream_t stream);
size_t paste_bicubic_dyn_workspace_bytes(int capacity, int g_h, int g_w, int out_cap);
int paste_bicubic_dyn(const float* logits, int capacity, const int* njobs_dev, int g_h, int g_w, const int* all_jobs,
                      int first_index, int n_all, const int* n_all_dev, uint8_t* mask, int n_img, int S_h, int S_w,
                      int out_cap, void* workspace, size_t ws_bytes, int* err_flag, cudaStream_t stream);
int lt_build_jobs(const int* boxes, const int* nbox, int B, int S_h, int S_w, int src_h, int src_w,
                  const int* orig_sizes, int* crop_jobs, int* paste_jobs, int capacity, int* counts, int chunk,
                  int* chunk_counts, cudaStream_t stream);
int mask_scale_u8(const uint8_t* in, uint8_t* out, size_t n, int mul, cudaStream_t stream);
int to_tensor_normalize(const uint8_t* in, float* out, size_t planes, int hw, int channels, const float* mean,
                        const float* stddev, cudaStream_t stream);

}  // namespace ucod
