// extern "C" surface of libsfb200 (declared in include/sfb200.h); thin argument checks + dispatch to the launchers.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ar_kernels.cuh"
#include "conv_tc.cuh"
#include "decoder_kernels.cuh"
#include "enc_kernels.cuh"
#include "mesh_kernels.cuh"

#include <atomic>
namespace sfb {
static std::atomic<long long> g_launches{0};
void count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static thread_local char g_err[512] = "";
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SFB200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;   // on by default; SFB200_PDL=0 disables
    }
    return v == 1;
}
void set_cuda_error(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
}  // namespace sfb

using namespace sfb;

extern "C" {

int sfb200_version(void) { return SFB200_VERSION; }

const char *sfb200_error_string(int code) {
    switch (code) {
        case SFB200_OK: return "ok";
        case SFB200_E_ARG: return "invalid argument or unsupported shape";
        case SFB200_E_CUDA: return "CUDA error (see sfb200_last_cuda_error)";
        case SFB200_E_STATE: return "call order violated";
    }
    return "unknown error";
}
const char *sfb200_last_cuda_error(void) { return g_err; }
int64_t sfb200_launch_count(void) { return (int64_t)g_launches.load(); }

int sfb200_code_gather(const int64_t *code_ind, const float *codebook, float *out, int B, int cells, int C, int n_codes,
                       void *stream) {
    if (!code_ind || !codebook || !out) return SFB200_E_ARG;
    return launch_code_gather(code_ind, codebook, out, B, cells, C, n_codes, as_stream(stream));
}
int sfb200_grid_to_channels_last(const float *src, float *dst, int B, int C, int64_t S, void *stream) {
    if (!src || !dst) return SFB200_E_ARG;
    return launch_to_channels_last(src, dst, B, C, S, as_stream(stream));
}
int sfb200_decoder_set_weights(const float *mlp_weights, void *stream) {
    if (!mlp_weights) return SFB200_E_ARG;
    SFB_TRY(decoder_set_weights_ffma(mlp_weights, as_stream(stream)));
    return decoder_set_weights_tc(mlp_weights, as_stream(stream));
}
int sfb200_decoder_points(const float *grid, const float *xtg, int64_t xtg_batch_stride, float *logits, int B, int R,
                          int64_t N, int impl, int sigmoid, void *stream) {
    if (!grid || !xtg || !logits) return SFB200_E_ARG;
    if (impl == 1) return launch_decoder_points_ffma(grid, xtg, xtg_batch_stride, logits, B, R, N, sigmoid, as_stream(stream));
    if (impl == 0) return launch_decoder_points_tc(grid, xtg, xtg_batch_stride, logits, B, R, N, sigmoid, as_stream(stream));
    return SFB200_E_ARG;
}
int sfb200_tokens_to_dense(const int64_t *tokens, const int64_t *empty_index, int64_t *dense, int B, int T, int cells,
                           int64_t end_pos, int64_t end_val, void *stream) {
    if (!tokens || !empty_index || !dense) return SFB200_E_ARG;
    return launch_tokens_to_dense(tokens, empty_index, dense, B, T, cells, end_pos, end_val, as_stream(stream));
}

int64_t sfb200_encoder_workspace_bytes(int B, int T, int n_codes) { return enc_workspace_bytes(B, T, n_codes); }
int sfb200_encode_cloud(const sfb200_enc_weights *w, const float *cloud, int B, int T, void *workspace, int64_t *raw_ind,
                        unsigned char *mask, float *grid_feat, void *stream) {
    return launch_encode_cloud(w, cloud, B, T, workspace, raw_ind, mask, grid_feat, as_stream(stream));
}
int sfb200_dense_to_tokens(const int64_t *raw_ind, const unsigned char *mask, int B, int cells, int n_codes, int max_len,
                           int64_t end_pos, int64_t end_val, void *workspace, int64_t *dense, int64_t *tokens, int32_t *lengths,
                           int64_t *modes, void *stream) {
    return launch_dense_to_tokens(raw_ind, mask, B, cells, n_codes, max_len, end_pos, end_val, workspace, dense, tokens, lengths,
                                  modes, as_stream(stream));
}

int sfb200_linear(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                  int act, void *stream) {
    if (!x || !W || !y || (act != 0 && act != 1)) return SFB200_E_ARG;
    return launch_linear(x, W, bias, residual, y, M, N, K, act, as_stream(stream));
}
int sfb200_linear_tc(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                     int act, void *stream) {
    if (!x || !W || !y || (act != 0 && act != 1)) return SFB200_E_ARG;
    return launch_linear_tc(x, W, bias, residual, y, M, N, K, act, as_stream(stream));
}
int64_t sfb200_tc_pretiled_floats(int N, int K) { return (N > 0 && K > 0 && K % 32 == 0) ? tc_pretiled_floats(N, K) : -1; }
int sfb200_tc_pretile(const float *W, float *Wt, int N, int K, void *stream) {
    if (!W || !Wt) return SFB200_E_ARG;
    return launch_tc_pretile(W, Wt, N, K, as_stream(stream));
}
int sfb200_linear_tc_ps(const float *x, const float *Wt, const float *bias, const float *residual, float *y, int M, int N, int K,
                        int act, void *stream) {
    if (!x || !Wt || !y || (act != 0 && act != 1)) return SFB200_E_ARG;
    return launch_linear_tc_ps(x, Wt, bias, residual, y, M, N, K, act, as_stream(stream));
}
// workspace of sfb200_chain_linear: [barrier counters 256 B | row statistics 64 x 32 pieces x 2 floats | split-K partials]
static const int64_t kChainWsBar = 256, kChainWsStats = 64 * 32 * 2 * 4;
int64_t sfb200_chain_workspace_bytes(void) { return kChainWsBar + kChainWsStats + (int64_t)chain_scratch_floats(160) * 4; }
int sfb200_chain_linear(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                        int act, const float *ln_w, const float *ln_b, void *workspace, void *stream) {
    if (!x || !W || !y || !workspace || (act != 0 && act != 1) || M < 1 || M > 64 || N < 1 || K < 32 || K % 32 != 0 || K > 4096)
        return SFB200_E_ARG;
    if ((ln_w == nullptr) != (ln_b == nullptr)) return SFB200_E_ARG;
    int grid = chain_grid_size();
    if (grid <= 0) return SFB200_E_CUDA;
    if (grid > 160) grid = 160;
    char *ws = static_cast<char *>(workspace);
    float *stats = reinterpret_cast<float *>(ws + kChainWsBar);
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M;
    a.bar = reinterpret_cast<unsigned int *>(ws);
    a.scratch = reinterpret_cast<float *>(ws + kChainWsBar + kChainWsStats);
    if (ln_w) {   // statistics-only phase over the input rows, then the GEMM applies the LayerNorm on load
        ChainPhase &s0 = a.ph[a.n_phases++];
        s0.residual = x; s0.stats_out = stats; s0.N = K;
    }
    ChainPhase &ph = a.ph[a.n_phases];
    SFB_TRY(chain_weight_map(W, N, K, &a.wmap[a.n_phases]));
    a.n_phases++;
    ph.x = x; ph.ln_g = ln_w; ph.ln_b = ln_b; ph.stats_in = ln_w ? stats : nullptr; ph.bias = bias; ph.residual = residual;
    ph.y = y; ph.N = N; ph.K = K; ph.act = act; ph.wait_before = ln_w ? 1 : 0;
    chain_plan(N, K, grid, &ph.tiles, &ph.splits);
    return launch_chain(a, as_stream(stream));
}
int sfb200_debug_chain_timeline(void *buf128) { return set_chain_timeline(static_cast<unsigned long long *>(buf128)); }
int sfb200_debug_ps_timeline(void *buf16) { return set_ps_timeline(static_cast<unsigned long long *>(buf16)); }
int sfb200_layernorm(const float *x, const float *w, const float *b, float *y, int rows, int d, void *stream) {
    if (!x || !w || !b || !y) return SFB200_E_ARG;
    return launch_layernorm(x, w, b, y, rows, d, as_stream(stream));
}
int sfb200_attn_decode(const float *qkv, float *kcache, float *vcache, float *out, float *part, int B, int H, int max_len,
                       int pos, const int32_t *pos_dev, int n_split, void *stream) {
    if (!qkv || !kcache || !vcache || !out || B < 1 || H < 1 || pos < 0 || pos >= max_len) return SFB200_E_ARG;
    if (pos_dev) return SFB200_E_ARG;  // device-side position is an engine-internal mode (state words)
    return launch_attn_decode(qkv, kcache, vcache, out, part, B, H, max_len, pos, nullptr, n_split, as_stream(stream));
}
int sfb200_attn_decode_grouped(const float *qkv, float *kcache, float *vcache, float *out, float *part, int32_t *counters, int B,
                               int H, int max_len, int pos, int group, int shared_len, void *stream) {
    if (!qkv || !kcache || !vcache || !out || !part || !counters || B < 1 || H < 1 || pos < 0 || pos >= max_len) return SFB200_E_ARG;
    if (shared_len < 0) return SFB200_E_ARG;
    return launch_attn_grouped(qkv, kcache, vcache, out, part, counters, B, H, max_len, pos, nullptr, group, shared_len, 0,
                               as_stream(stream));
}
int sfb200_attn_prefill(const float *qkv, float *kcache, float *vcache, float *out, int B, int H, int T, int max_len,
                        void *stream) {
    if (!qkv || !kcache || !vcache || !out || B < 1 || H < 1 || T < 0 || T > max_len) return SFB200_E_ARG;
    return launch_attn_prefill(qkv, kcache, vcache, out, B, H, T, max_len, as_stream(stream));
}
int sfb200_ar_sample(const float *logits, int64_t *tokens, float *hist_out, const float *noise_sample,
                     const float *noise_best, int B, int V, int max_len, int L, int L_cond, int tuple_i,
                     const int64_t *end_tokens, const sfb200_ar_sampling *sp, void *stream) {
    if (!logits || !tokens || !noise_sample || !noise_best || !end_tokens || !sp) return SFB200_E_ARG;
    if (B < 1 || L < 1 || L >= max_len || L_cond < 1 || L_cond > L || (tuple_i != 0 && tuple_i != 1)) return SFB200_E_ARG;
    if (!(sp->temperature > 0.f)) return SFB200_E_ARG;
    SampleLaunch p;
    p.logits = logits; p.tokens = tokens; p.hist = hist_out; p.noise_sample = noise_sample; p.noise_best = noise_best;
    p.B = B; p.V = V; p.max_len = max_len; p.L = L; p.L_cond = L_cond; p.tuple_i = tuple_i;
    p.end0 = end_tokens[0]; p.end1 = end_tokens[1]; p.sp = *sp; p.st = nullptr;
    p.hist_row_stride = V; p.noise_step_stride = 0; p.noise_row_stride = V;
    return launch_sample(p, as_stream(stream));
}

int64_t sfb200_big_partial_floats(void) { return (int64_t)big_partial_floats(); }
int sfb200_split_lo(const float *x, float *lo, int64_t n, void *stream) {
    if (n < 0) return SFB200_E_ARG;
    return launch_split_lo(x, lo, (size_t)n, as_stream(stream));
}
int sfb200_linear_big(const float *x, const float *x_lo, const float *W, const float *W_lo, const float *bias,
                      const float *residual, float *y, float *y_lo, int M, int N, int K, int act, float *partial,
                      int32_t *counters, void *stream) {
    return launch_linear_big(x, x_lo, W, W_lo, bias, residual, y, y_lo, M, N, K, act, partial, counters, as_stream(stream));
}

int sfb200_conv3d_tc(const float *in, const float *in_lo, const float *w, const float *w_lo, const float *bias, float *out,
                     double *stats, int B, int Z, int Y, int X, int Cin, int Cout, int taps, int relu, void *stream) {
    return launch_conv3d_tc(in, in_lo, w, w_lo, bias, out, stats, B, Z, Y, X, Cin, Cout, taps, relu, as_stream(stream));
}
int sfb200_conv_prep(const float *src0, int C0, int sh0, const double *st0, double n0, const float *src1, int C1, int sh1,
                     const double *st1, double n1, const float *gamma, const float *beta, int groups, float *dst, float *dst_lo,
                     int B, int Z, int Y, int X, void *stream) {
    return launch_conv_prep(src0, C0, sh0, st0, n0, src1, C1, sh1, st1, n1, gamma, beta, groups, dst, dst_lo, B, Z, Y, X,
                            as_stream(stream));
}
int sfb200_pool_stats(const float *src, float *dst, double *stats, int B, int Zo, int Yo, int Xo, int C, int win, void *stream) {
    return launch_pool_stats(src, dst, stats, B, Zo, Yo, Xo, C, win, as_stream(stream));
}
int sfb200_gather_codes_cl(const int64_t *idx, const float *codebook, float *out, double *stats, int B, int cells, int C,
                           int n_codes, void *stream) {
    return launch_gather_codes_cl(idx, codebook, out, stats, B, cells, C, n_codes, as_stream(stream));
}

int sfb200_mesh_mark_edges(const float *grid, int R, float thresh, int32_t *flag, void *stream) {
    return launch_mesh_mark_edges(grid, R, thresh, flag, as_stream(stream));
}
int sfb200_mesh_emit_vertices(const float *grid, int R, float thresh, const int32_t *flag, const int32_t *vid, float *verts,
                              void *stream) {
    return launch_mesh_emit_vertices(grid, R, thresh, flag, vid, verts, as_stream(stream));
}
int sfb200_mesh_count_faces(const float *grid, int R, float thresh, int32_t *count, void *stream) {
    return launch_mesh_count_faces(grid, R, thresh, count, as_stream(stream));
}
int sfb200_mesh_emit_faces(const float *grid, int R, float thresh, const int32_t *vid, const int32_t *foff, const float *verts,
                           int32_t *faces, void *stream) {
    return launch_mesh_emit_faces(grid, R, thresh, vid, foff, verts, faces, as_stream(stream));
}

}  // extern "C"
