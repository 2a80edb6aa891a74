// COD metric suite on the device (see metrics.cu).
#pragma once
#include "common.cuh"

namespace ucod {

constexpr int METRICS_OUT = 7 + 512;  // acc, iou, mae, sm, em_adp, fm_adp, wfm, em_curve[256], fm_curve[256]

size_t cod_metrics_workspace_bytes(int B, int h, int w);
// gt, pred: fp32 [B,h,w] ; out: fp64 [B, METRICS_OUT]
int cod_metrics(const float* gt, const float* pred, int B, int h, int w, double* out, void* workspace, size_t ws_bytes,
                cudaStream_t stream);

}  // namespace ucod
