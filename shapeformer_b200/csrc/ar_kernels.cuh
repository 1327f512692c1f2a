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
    float *logp = nullptr;          // (B, max_steps, 2) log-probability of every sampled token, or NULL
    int64_t logp_row_stride = 0;
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
int launch_layernorm(const float *x, const float *w, const float *b, float *y, int rows, int d, cudaStream_t s,
                     float *y_lo = nullptr);
int launch_gemv(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                int act, cudaStream_t s);
int launch_linear(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                  int act, cudaStream_t stream);
int launch_linear_tc(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                     int act, cudaStream_t stream);
// ---- large-M GEMM (tc_big.cu): TMA-fed SS-mode 3xTF32 with the lo parts of both operands kept in memory
constexpr int BG_MAX_UNITS = 320;     // (tile, split) units the split-K scratch is sized for
constexpr int BG_MAX_TILES = 1024;    // arrival counters
size_t big_partial_floats();
int launch_split_lo(const float *x, float *lo, size_t n, cudaStream_t s);
int launch_linear_big(const float *x, const float *x_lo, const float *W, const float *W_lo, const float *bias, const float *residual,
                      float *y, float *y_lo, int M, int N, int K, int act, float *partial, int *tile_cnt, cudaStream_t stream);
int64_t tc_pretiled_floats(int N, int K);
int set_ps_timeline(unsigned long long *buf);
int launch_tc_pretile(const float *W, float *Wt, int N, int K, cudaStream_t s);
int launch_linear_tc_ps(const float *x, const float *Wt, const float *bias, const float *residual, float *y, int M, int N, int K,
                        int act, cudaStream_t stream);
// ---- decode-step GEMM chain (ar_chain.cu): a list of nn.Linear phases run by ONE persistent kernel
constexpr int CH_MAXP = 5;            // phases per launch
constexpr int CH_MAX_BARRIERS = 15;   // grid barriers per launch; counters [0, CH_MAX_BARRIERS) + the exit counter
struct alignas(64) TensorMapBlob { unsigned char b[128]; };   // a CUtensorMap (cuda.h) without the driver header
struct ChainPhase {
    const float *x;          // (M, K) input activations (GEMM phases)
    const float *ln_g, *ln_b;   // LayerNorm weight / bias applied to x on load, or NULL
    const float *stats_in;   // (M, ceil(K/128), 2) per-row (mean, M2) pieces of x when ln_g != NULL
    const float *bias;       // (N) or NULL
    const float *residual;   // (M, N) or NULL (may alias y)
    float *y;                // (M, N) output, or NULL (statistics-only phase)
    float *stats_out;        // (M, ceil(N/128), 2) pieces of the OUTPUT rows (a LayerNorm follows), or NULL
    int N, K;
    int tiles, splits;       // tiles = ceil(N/128); tiles * splits <= grid.  tiles == 0: no GEMM, the reduction step only sees
                             // `residual` (used to produce stats_out for a row vector written by an earlier kernel)
    int act;                 // 0 none, 1 exact-erf GELU
    int wait_before;         // 1: x / stats_in are produced by an earlier phase of the same launch (grid barrier first)
};
struct ChainArgs {
    TensorMapBlob wmap[CH_MAXP];   // weight tensor map of each GEMM phase (chain_weight_map)
    ChainPhase ph[CH_MAXP];
    int n_phases, M;
    float *scratch;                // chain_scratch_floats(grid) floats: split-K partial tiles
    unsigned int *bar;             // CH_MAX_BARRIERS + 1 counters, zero between launches
};
int chain_weight_map(const float *W, int N, int K, void *map_out /* TensorMapBlob */);
int chain_grid_size();
void chain_plan(int N, int K, int grid, int *tiles, int *splits);
size_t chain_scratch_floats(int grid);
int launch_chain(const ChainArgs &args, cudaStream_t stream);
int set_chain_timeline(unsigned long long *buf);

int launch_attn_decode(const float *qkv, float *kc, float *vc, float *out, float *part, int B, int H, int max_len, int pos,
                       const int32_t *st, int n_split, cudaStream_t s, int group = 1, int lcond = 0, int lcond_delta = 0);
int launch_attn_grouped(const float *qkv, float *kc, float *vc, float *out, float *part, int *cnt, int B, int H, int max_len, int pos,
                        const int32_t *st, int group, int lcond, int lcond_delta, cudaStream_t s, float *out_lo = nullptr);
int launch_attn_prefill(const float *qkv, float *kc, float *vc, float *out, int B, int H, int T, int max_len,
                        cudaStream_t s, const int32_t *rowmap = nullptr);
int launch_attn_prefill_tc(const float *qkv, float *kc, float *vc, float *out, int B, int H, int T, int max_len, cudaStream_t s,
                           const int32_t *rowmap = nullptr);
int launch_sample(const SampleLaunch &p, cudaStream_t s);

}  // namespace sfb
