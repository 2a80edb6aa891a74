// Discriminator forward + APM fusion (see discriminator.cu).
#pragma once
#include "common.cuh"

namespace ucod {

struct DiscWeights {
    const float* conv1;  // [32,1,3,3]
    const float *bn1_w, *bn1_b;
    float *bn1_mean, *bn1_var;
    const float* conv2;  // [16,32,3,3]
    const float *bn2_w, *bn2_b;
    float *bn2_mean, *bn2_var;
    const float* conv3;  // [8,16,3,3]
    const float *bn3_w, *bn3_b;
    float *bn3_mean, *bn3_var;
    const float* lin_w;  // [1, 8*h3*h3]
    const float* lin_b;  // [1]
};

size_t discriminator_workspace_bytes(int B, int fs);
size_t discriminator_workspace_bytes_groups(int B, int fs, int groups);
// `groups` > 1: mask [groups*B, 1, fs, fs] holds that many independent forward calls (each with its own batch statistics,
// running buffers updated in call order), run by one set of launches; prob [groups*B]; workspace from the _groups query.
int discriminator_forward(const float* mask, int B, int fs, const DiscWeights& w, int bn_train, int update_running,
                          float* prob, void* workspace, size_t ws_bytes, cudaStream_t stream, int groups = 1);
struct DiscGrads {  // fp32, same shapes as the weights; accumulated into
    float *conv1, *bn1_w, *bn1_b, *conv2, *bn2_w, *bn2_b, *conv3, *bn3_w, *bn3_b, *lin_w, *lin_b;
};
size_t discriminator_backward_workspace_bytes(int B, int fs);
// Backward of one discriminator_forward(bn_train = 1) call whose workspace is `fwd_workspace`; the loss is
// BCE(prob, label) averaged over n_total samples (the epoch's two calls share n_total = 2B).  Gradients and
// `loss` are accumulated.
int discriminator_backward(const float* mask, int B, int fs, const DiscWeights& w, const float* prob, float label,
                           int n_total, const DiscGrads& g, float* loss, void* fwd_workspace, void* workspace,
                           size_t ws_bytes, cudaStream_t stream);
int apm_binarize(const float* student, const float* teacher, const float* pl, float* s_mask, float* t_mask,
                 float* p_mask, size_t n, cudaStream_t stream);
int apm_merge(const float* pl, const float* t_mask, const float* p_s, const float* p_p, float epoch_term,
              float* merged, float* weight, float* dis_loss, int B, int npix, cudaStream_t stream);

}  // namespace ucod
