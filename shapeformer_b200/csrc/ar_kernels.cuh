// Internal launch interface of the AR kernels (ar_kernels.cu), used by the engine (ar_engine.cu) and the C-ABI.
#pragma once
#include "common.cuh"

namespace sfb {

enum { ST_STEPS = 0, ST_ENDED = 1, ST_LEN = 2, ST_LCOND = 3, ST_CHUNK = 4, ST_WORDS = 8 };

struct SampleLaunch {
    const float *logits;
    int64_t *tokens;
    float *hist;
    const float *noise_sample, *noise_best;
    int B, V, max_len, L, L_cond, tuple_i;
    int64_t end0, end1;
    sfb200_ar_sampling sp;
    const int32_t *st;
    int64_t hist_row_stride, noise_step_stride;
    int noise_row_stride;
};

int launch_state_init(int32_t *st, int L_cond, cudaStream_t s);
int launch_chunk_reset(int32_t *st, cudaStream_t s);
int launch_advance(int32_t *st, const int64_t *tokens, int B, int max_len, int64_t end0, int64_t end1, cudaStream_t s);
int launch_embed(const int64_t *tokens, const float *emb0, const float *emb1, const float *embx, const float *pos_emb,
                 const float *cond_pos_emb, float *x, int B, int d, int max_len, int t0, int T, int L_cond, int64_t end0,
                 const int32_t *st, cudaStream_t s, const int32_t *rowmap = nullptr);
int launch_add_target(const float *x_in, float *x_out, const int64_t *tokens, const float *emb0, int B, int d, int max_len,
                      int t0, int T, const int32_t *st, cudaStream_t s, int Tin, const int32_t *rowmap = nullptr);
int launch_take_last(const float *x, float *out, int B, int d, int T, cudaStream_t s, const int32_t *rowmap = nullptr);
int launch_prefix_copy(float *kv, float *x0, const int32_t *dst_rows, const int32_t *src_rows, int n_dup, int H, int max_len, int T,
                       int d, int n_blocks, int64_t per_block, cudaStream_t s);
int launch_layernorm(const float *x, const float *w, const float *b, float *y, int rows, int d, cudaStream_t s);
int launch_gemv(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                int act, cudaStream_t s);
int launch_linear(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                  int act, cudaStream_t stream);
int launch_linear_tc(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                     int act, cudaStream_t stream);
int64_t tc_pretiled_floats(int N, int K);
int set_ps_timeline(unsigned long long *buf);
int launch_tc_pretile(const float *W, float *Wt, int N, int K, cudaStream_t s);
int launch_linear_tc_ps(const float *x, const float *Wt, const float *bias, const float *residual, float *y, int M, int N, int K,
                        int act, cudaStream_t stream);
int launch_attn_decode(const float *qkv, float *kc, float *vc, float *out, float *part, int B, int H, int max_len, int pos,
                       const int32_t *st, int n_split, cudaStream_t s, int group = 1, int lcond = 0, int lcond_delta = 0);
int launch_attn_prefill(const float *qkv, float *kc, float *vc, float *out, int B, int H, int T, int max_len,
                        cudaStream_t s, const int32_t *rowmap = nullptr);
int launch_sample(const SampleLaunch &p, cudaStream_t s);

}  // namespace sfb
