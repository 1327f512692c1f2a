// Internal launch interface of the decoder-side kernels.
#pragma once
#include "common.cuh"

namespace sfb {
int launch_code_gather(const int64_t *ind, const float *cb, float *out, int B, int cells, int C, int n_codes, cudaStream_t s);
int launch_to_channels_last(const float *src, float *dst, int B, int C, int64_t S, cudaStream_t s);
int launch_tokens_to_dense(const int64_t *tokens, const int64_t *empty, int64_t *dense, int B, int T, int cells,
                           int64_t end_pos, int64_t end_val, cudaStream_t s);
int decoder_set_weights_ffma(const float *w, cudaStream_t s);
int launch_decoder_points_ffma(const float *grid, const float *xtg, int64_t xtg_bstride, float *logits, int B, int R,
                               int64_t N, int sigmoid, cudaStream_t s);
// tcgen05 path (decoder_tc.cu)
int decoder_set_weights_tc(const float *w, cudaStream_t s);
int launch_decoder_points_tc(const float *grid, const float *xtg, int64_t xtg_bstride, float *logits, int B, int R,
                             int64_t N, int sigmoid, cudaStream_t s);
}  // namespace sfb
