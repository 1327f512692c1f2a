// Internal launch interface of the encoder kernels (enc_kernels.cu).
#pragma once
#include "common.cuh"

namespace sfb {

int64_t enc_workspace_bytes(int B, int T, int n_codes);
int launch_encode_cloud(const sfb200_enc_weights *W, const float *cloud, int B, int T, void *workspace, int64_t *raw_ind,
                        unsigned char *mask, float *grid_feat, cudaStream_t s);
int launch_dense_to_tokens(const int64_t *raw_ind, const unsigned char *mask, int B, int cells, int n_codes, int max_len, int64_t end0,
                           int64_t end1, void *workspace, int64_t *dense, int64_t *tokens, int32_t *lengths, int64_t *modes,
                           cudaStream_t s);

}  // namespace sfb
