/*
 * libsfb200 — C-ABI of the B200-native ShapeFormer hot path (sm_100a).
 *
 * The reference (QhelDIV/ShapeFormer) is pure Python/PyTorch and has NO native interface; these entry points are what a
 * ctypes binding on the reference side would call to replace the functions listed beside each one (paths relative to
 * /root/reference/shapeformer/models/).  INTEGRATION.md shows the binding.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer borrowed from the caller (torch `data_ptr()`), kept alive by the caller;
 *   - the library never allocates device memory, never synchronises a stream, never throws; all work is enqueued on
 *     `stream` (a `cudaStream_t` passed as void*).  The only host-side resources it owns are those of an sfb200_ar handle
 *     (a captured step graph, CUDA events, and a few hundred bytes of pinned staging for the per-batch row lists;
 *     sfb200_ar_begin_shared waits on the EVENT of the handle's previous staging copy, never on the stream);
 *   - return value 0 = ok, negative = SFB200_E_* (see sfb200_error_string); CUDA launch errors are returned as
 *     SFB200_E_CUDA with the message available from sfb200_last_cuda_error();
 *   - indices are int64 (the reference's torch.long), activations / weights fp32, row-major contiguous.
 */
#ifndef SFB200_H
#define SFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB200_VERSION 100

#define SFB200_OK 0
#define SFB200_E_ARG (-1)      /* invalid argument / unsupported shape */
#define SFB200_E_CUDA (-2)     /* CUDA runtime error (see sfb200_last_cuda_error) */
#define SFB200_E_STATE (-3)    /* call order violated (e.g. step before prefill) */

int sfb200_version(void);
const char *sfb200_error_string(int code);
const char *sfb200_last_cuda_error(void);
/* Number of kernels this library has launched in this process (a graph replay counts every kernel node it contains). */
int64_t sfb200_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * VQDIF decoder side  (vqdif/vqdif.py:60-76, vqdif/dec.py:62-100, vqdif/quantizer.py:19-30)
 * ------------------------------------------------------------------------------------------------------------------ */

/* Quantizer.get_code (vqdif/quantizer.py:19-30): out[b][c][cell] = codebook[code_ind[b][cell]][c].
 * code_ind (B, cells) int64 in [0, n_codes), codebook (n_codes, C) fp32, out (B, C, cells) fp32 (= BCDHW). */
int sfb200_code_gather(const int64_t *code_ind, const float *codebook, float *out, int B, int cells, int C, int n_codes,
                       void *stream);

/* Layout change of the decoder's feature grid, (B, C, S) channel-first -> (B, S, C) channel-last (S = D*H*W), so that one
 * trilinear corner is one contiguous C*4-byte read in sfb200_decoder_points. */
int sfb200_grid_to_channels_last(const float *src, float *dst, int B, int C, int64_t S, void *stream);

/* Number of floats of the packed LocalDecoder MLP weights (hidden = c_dim = 32, n_blocks = 5): see
 * shapeformer_b200/decoder.py::pack_mlp_weights for the order. */
#define SFB200_DEC_HIDDEN 32
#define SFB200_DEC_BLOCKS 5
#define SFB200_DEC_MLP_FLOATS (32 * 3 + 32 + 5 * (3 * (32 * 32 + 32)) + 32 + 1)

/* Upload the packed MLP weights (device pointer, SFB200_DEC_MLP_FLOATS floats) into the library's constant bank for the
 * current device.  Must be called (on `stream`) before sfb200_decoder_points whenever the weights change. */
int sfb200_decoder_set_weights(const float *mlp_weights, void *stream);

/* LocalDecoder.forward minus the conv prologue (vqdif/dec.py:85-97; normalize_3d_coordinate vqdif/common.py:260-276;
 * ResnetBlockFC vqdif/layers.py:39-48), including VQDIF.decode's Xtg/2 (vqdif/vqdif.py:71):
 *   grid   (B, R, R, R, 32) fp32 channel-last feature grid (output of UNet3D+Upsampler),
 *   xtg    (B or 1, N, 3) fp32 query points in [-1,1]; xtg_batch_stride = N*3 or 0 when shared by all shapes,
 *   logits (B, N) fp32 occupancy logits (the reference's (B,N,1)).
 * impl: 0 = default (tcgen05 kernel, TF32 hi/lo operand split, 3 MMAs per product), 1 = fp32 FFMA kernel (cross-check).
 * sigmoid != 0 writes occupancy = 1/(1+exp(-logit)) instead (decode_sample_indices, shapeformer/shapeformer.py:388). */
int sfb200_decoder_points(const float *grid, const float *xtg, int64_t xtg_batch_stride, float *logits, int B, int R,
                          int64_t N, int impl, int sigmoid, void *stream);

/* filter_end_tokens + batch_sparse2dense (shapeformer/common.py:50-55,171-189; caller shapeformer/shapeformer.py:342-351):
 * dense[b][:] = empty_index[b]; for t in order: if pos,val are not end tokens: dense[b][pos] = val (later writes win).
 * tokens (B, T, 2) int64, empty_index (B) int64, dense (B, cells) int64. */
int sfb200_tokens_to_dense(const int64_t *tokens, const int64_t *empty_index, int64_t *dense, int B, int T, int cells,
                           int64_t end_pos, int64_t end_val, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * VQDIF encoder + quantiser: partial point cloud -> code grid -> (pos, val) conditioning tuples   
 *   LocalPoolPointnet.forward vqdif/enc.py:66-140, Downsampler vqdif/updown.py:98-113, Quantizer.forward vqdif/quantizer.py:31-53,
 *   VQDIF.quantize_cloud vqdif/vqdif.py:50-58, batch_dense2sparse shapeformer/common.py:84-122,152-169
 * ------------------------------------------------------------------------------------------------------------------ */

/* Device pointers to the encoder weights (shipped shapenet_res16 shapes: hidden = c_dim = 32, 5 ResNet-FC blocks, 64^3 scatter
 * grid, Downsampler 32 -> 64 -> 128 over two k2s2 + k1 'crg' pairs, 128-d codes).  Linear weights are (out, in) row-major like
 * the state_dict; ds_wT[i] = encoder.downsampler.blocks.i.conv.weight.permute(2, 3, 4, 1, 0) made contiguous = (k^3 * Cin, Cout). */
typedef struct sfb200_enc_weights {
    const float *fc_pos_w, *fc_pos_b;                  /* (64, 3), (64) */
    const float *fc0_w[5], *fc0_b[5];                  /* blocks.i.fc_0: (32, 64), (32) */
    const float *fc1_w[5], *fc1_b[5];                  /* blocks.i.fc_1: (32, 32), (32) */
    const float *sc_w[5];                              /* blocks.i.shortcut.weight: (32, 64) */
    const float *fcc_w, *fcc_b;                        /* fc_c: (32, 32), (32) */
    const float *ds_wT[4], *ds_gn_w[4], *ds_gn_b[4];   /* Downsampler convs (transposed, see above) + GroupNorm(8) affine */
    const float *codebook;                             /* quantizer.embedding.weight (n_codes, 128) */
    int n_codes;
} sfb200_enc_weights;

/* Bytes of device workspace for sfb200_encode_cloud with B clouds of T points (about 150 MB per cloud). */
int64_t sfb200_encoder_workspace_bytes(int B, int T, int n_codes);

/* VQDIF.encode + Quantizer.forward: cloud (B, T, 3) fp32 in [-1, 1] -> raw_ind (B, 16, 16, 16) int64 nearest-code indices of
 * every cell, mask (B, 16^3) uint8 = cells that contain a point (enc.py:84-91, layout [z][y][x]), and optionally grid_feat
 * (B, 128, 16^3) fp32, the encoder output the reference feeds to the quantiser. */
int sfb200_encode_cloud(const sfb200_enc_weights *w, const float *cloud, int B, int T, void *workspace, int64_t *raw_ind,
                        unsigned char *mask, float *grid_feat, void *stream);

/* VQDIF.quantize_cloud + batch_dense2sparse: modes[0] = most frequent raw index of the whole batch (smallest on ties), dense =
 * mask ? raw : modes[0]; modes[1] = mode of dense (the empty index); tokens (B, max_len, 2) = per row the cells != modes[1] in
 * raveled order as (pos, val), padded with the end tokens; lengths[b] = number of such cells (may exceed max_len: the caller
 * crops like unpack_sparse).  workspace: n_codes int32. */
int sfb200_dense_to_tokens(const int64_t *raw_ind, const unsigned char *mask, int B, int cells, int n_codes, int max_len,
                           int64_t end_pos, int64_t end_val, void *workspace, int64_t *dense, int64_t *tokens, int32_t *lengths,
                           int64_t *modes, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Autoregressive sampler  (shapeformer/shapeformer.py:54-123, transformer/mingpt.py:46-111,185-319,
 *                          shapeformer/representers.py:120-155,187-196,432-442, shapeformer/common.py:260-299)
 * ------------------------------------------------------------------------------------------------------------------ */

typedef struct sfb200_ar_config {
    /* model shape — CondTupleGPT.__init__ (transformer/mingpt.py:187-244) */
    int n_embd;          /* 1024; multiple of 64 */
    int n_head;          /* 16; head dim must be 64 */
    int n_layers[2];     /* {20, 4} */
    int block_size;      /* 812 */
    int vocab[2];        /* {4097, 4097}  (<= 8192) */
    int extra_vocab;     /* 4097 */
    int64_t end_tokens[2]; /* {4096, 4096} */
    /* run shape */
    int max_rows;        /* B: rows sampled together */
    int max_len;         /* KV-cache capacity in positions (>= L_cond + max_steps) */
    int max_steps;       /* capacity of the logits history */
    int prefill_rows;    /* rows pushed through prefill together (workspace sizing) */
    int max_cond;        /* largest L_cond accepted */
    int keep_history;    /* 1: store masked logits of every sub-step (B,steps,V) x2 like the reference */
} sfb200_ar_config;

/* weight tensor ids for sfb200_ar_weight_offset (names = the reference state_dict keys) */
enum {
    SFB200_W_POS_EMB = 0,        /* pos_emb (block_size, d) */
    SFB200_W_COND_POS_EMB,       /* cond_pos_emb (block_size, d) */
    SFB200_W_TOK_EMB0,           /* tok_embs.0.weight (V0, d) */
    SFB200_W_TOK_EMB1,           /* tok_embs.1.weight (V1, d) */
    SFB200_W_EXTRA_EMB,          /* extra_tok_embs.0.weight (Ve, d) */
    SFB200_W_HEAD_LN_W,          /* heads.g.0.weight (d)            [group] */
    SFB200_W_HEAD_LN_B,          /* heads.g.0.bias (d)              [group] */
    SFB200_W_HEAD_W,             /* heads.g.1.weight (V_g, d)       [group] */
    SFB200_W_LN1_W,              /* blocks.g.l.ln1.weight           [group, layer] */
    SFB200_W_LN1_B,
    SFB200_W_QKV_W,              /* cat(attn.query, attn.key, attn.value).weight (3d, d) */
    SFB200_W_QKV_B,              /* (3d) */
    SFB200_W_PROJ_W,             /* attn.proj.weight (d, d) */
    SFB200_W_PROJ_B,
    SFB200_W_LN2_W,
    SFB200_W_LN2_B,
    SFB200_W_FC1_W,              /* mlp.0.weight (4d, d) */
    SFB200_W_FC1_B,
    SFB200_W_FC2_W,              /* mlp.2.weight (d, 4d) */
    SFB200_W_FC2_B,
    SFB200_W_COUNT
};

/* sizes (bytes unless stated) the caller must allocate */
int64_t sfb200_ar_weight_floats(const sfb200_ar_config *cfg);
int64_t sfb200_ar_weight_offset(const sfb200_ar_config *cfg, int tensor_id, int group, int layer); /* in floats, <0 = bad id */
int64_t sfb200_ar_kv_bytes(const sfb200_ar_config *cfg);
int64_t sfb200_ar_workspace_bytes(const sfb200_ar_config *cfg);
int64_t sfb200_ar_history_floats(const sfb200_ar_config *cfg);  /* 0 when keep_history == 0 */

/* floats of the optional pre-split copy of the transformer's GEMM weights (see sfb200_ar_set_pretiled) */
int64_t sfb200_ar_pretiled_floats(const sfb200_ar_config *cfg);

typedef struct sfb200_ar sfb200_ar;  /* host-side handle (small, malloc'ed); owns no device memory */

/* Bind caller-owned device buffers.  weights: packed per sfb200_ar_weight_offset.  tokens: (max_rows, max_len, 2) int64
 * — the reference's `sampled` buffer (shapeformer/shapeformer.py:66-69).  history may be NULL iff keep_history == 0. */
int sfb200_ar_create(const sfb200_ar_config *cfg, const float *weights, void *kv_cache, void *workspace, int64_t *tokens,
                     float *history, sfb200_ar **out);
void sfb200_ar_destroy(sfb200_ar *h);

/* Optional: bind a caller-owned buffer of sfb200_ar_pretiled_floats(cfg) floats and fill it (on `stream`) with the pre-split
 * TF32 tiles of every block / head weight.  Decode steps with 9..64 rows then use sfb200_linear_tc_ps. */
int sfb200_ar_set_pretiled(sfb200_ar *h, float *pretiled, void *stream);

/* Optional: bind a caller-owned buffer of sfb200_ar_weight_floats(cfg) floats and fill it (on `stream`) with the LOW parts of
 * the two-term TF32 split of the weight blob (same layout).  Every nn.Linear over more than 64 rows (prefill, decode batches
 * of > 64 rows) then runs the TMA-fed tensor-core kernel of csrc/tc_big.cu, which reads both parts of both operands from
 * memory (Block.forward / heads, transformer/mingpt.py:74-111,222-231). */
int sfb200_ar_set_lo_weights(sfb200_ar *h, float *lo_blob, void *stream);

/* y = act(x W^T + bias) + residual with that kernel (any M; x_lo / W_lo = low parts as produced by sfb200_split_lo; y_lo
 * optional output = low part of y; partial: sfb200_big_partial_floats() floats, counters: 2048 zeroed int32, both may be NULL
 * = no split-K).  act: 0 none, 1 exact-erf GELU. */
int64_t sfb200_big_partial_floats(void);
int sfb200_split_lo(const float *x, float *lo, int64_t n, void *stream);
int sfb200_linear_big(const float *x, const float *x_lo, const float *W, const float *W_lo, const float *bias,
                      const float *residual, float *y, float *y_lo, int M, int N, int K, int act, float *partial,
                      int32_t *counters, void *stream);

typedef struct sfb200_ar_sampling {
    int top_k;                    /* <= 0 disables (common.py:265) */
    float top_p;                  /* <= 0 disables (common.py:271) */
    float temperature;
    int best_in_first;            /* row 0 takes the greedy (top_k=1, top_p=0.001) draw (shapeformer.py:96-101) */
    int mask_invalid;             /* representer attribute (representers.py:57,134) */
    int mask_invalid_completion;  /* representer attribute (representers.py:141) */
} sfb200_ar_sampling;

/* Start a batch: B rows, each with L_cond conditioning tuples already written to tokens[:, :L_cond].  Runs the prefill:
 * blocks[0] over positions [0, L_cond), blocks[1] over [0, L_cond-1), filling both KV caches, and leaves logits0 of the
 * last conditioning position ready for step 0.  Replaces the first (and, being cached, every later) full forward of
 * CondTupleGPT.sample_next_tuple (transformer/mingpt.py:297-310) and AR_N.get_extra_indices (representers.py:187-196). */
int sfb200_ar_begin(sfb200_ar *h, int B, int L_cond, const sfb200_ar_sampling *sp, void *stream);

/* Same, with conditioning sharing: row_src (HOST array of B ints, row_src[b] <= b) names an earlier row whose conditioning
 * tuples are identical to row b's (row_src[b] == b for the first row of each group) — the reference's sample_n expansion of one
 * shape (shapeformer/shapeformer.py:229).  Only the group leaders are pushed through the prefill; the other rows receive a
 * copy of the leader's prefix K/V.  Results are identical to sfb200_ar_begin. */
int sfb200_ar_begin_shared(sfb200_ar *h, int B, int L_cond, const sfb200_ar_sampling *sp, const int32_t *row_src, void *stream);

/* Run `n_steps` AR steps (each = pos sub-pass + val sub-pass = one (pos,val) tuple per row), enqueued back to back.
 * noise: (n_steps, 4, B, Vmax) fp32 Exp(1) draws in the reference's order per step: pos-sample, pos-best, val-sample,
 * val-best (what torch.multinomial draws internally, shapeformer/common.py:296; Vmax = max(vocab)).
 * use_graph != 0 captures one step into a CUDA graph and replays it.  Steps continue past "all rows ended"; the first
 * step index at which every row's newest tuple holds an end token is recorded (sfb200_ar_status). */
int sfb200_ar_steps(sfb200_ar *h, int n_steps, const float *noise, int use_graph, void *stream);

/* Device-side status words the caller may copy back after a sync: status[0] = steps done, status[1] = first step index
 * (0-based) at which all rows had ended, or -1.  Returns a device pointer to int32[4]. */
const int32_t *sfb200_ar_status_ptr(const sfb200_ar *h);

/* Device pointer (inside the workspace) to the log-probabilities of the sampled tokens, (max_rows, max_steps, 2) fp32:
 * entry [b][j][i] = log_softmax(masked logits of sub-step i of step j)[sampled token] — what compute_log_probs
 * (shapeformer/shapeformer.py:407-418) derives from the logits history, accumulated on the device so that ranking the sample_n
 * completions does not need the (B, steps, 4097) x 2 history at all.  Valid for steps executed since sfb200_ar_begin. */
const float *sfb200_ar_logprob_ptr(const sfb200_ar *h);

/* Measurement aid for bench.py: when enabled (eager stepping only), every attention launch of sfb200_ar_steps is bracketed
 * by CUDA events on `stream`.  sfb200_ar_profile_read must be called after the stream has been synchronised; it returns the
 * summed device time (ms), the number of launches and the ALGORITHMIC bytes of those launches
 * (per launch: key rows * 2*d*4 (K,V read) + B * [2*d*4 (append) + 2*d*4 (q in, out)]; key rows = B*pos, or
 * with conditioning-prefix sharing L_cond per group + (pos - L_cond) per row) and resets the counters. */
int sfb200_ar_profile(sfb200_ar *h, int enable);
int sfb200_ar_profile_read(sfb200_ar *h, double *attn_ms, int64_t *attn_launches, double *attn_bytes,
                           double *attn_bytes_per_row /* may be NULL: SURVEY per-row formula without prefix sharing */);

/* ---- individual AR operators (exposed for parity tests and profiling; the same kernels the calls above enqueue) ---- */

/* y = act(x @ W^T + bias) + residual.  x (M,K), W (N,K), bias (N) or NULL, residual (M,N) or NULL (may alias y),
 * act: 0 none, 1 exact-erf GELU (nn.GELU).  nn.Linear of mingpt.py:56-61,99-104,228. */
int sfb200_linear(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                  int act, void *stream);

/* Same contract as sfb200_linear on the tcgen05 tensor cores with fp32-level accuracy: operands split into two TF32 parts,
 * three MMAs per product, TMEM accumulator promoted to fp32 registers every two K chunks, cluster (DSMEM) split-K. */
int sfb200_linear_tc(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                     int act, void *stream);

/* Decode-step variant (M <= 64) reading PRE-SPLIT weight tiles: sfb200_tc_pretile rewrites W (N, K) as TF32 hi/lo tiles in the
 * tensor core's shared-memory image (sfb200_tc_pretiled_floats(N, K) floats); sfb200_linear_tc_ps then streams each 32 KB
 * (tile, K-chunk) with one TMA bulk copy and feeds both MMA operands from shared memory.  Same results as sfb200_linear_tc. */
int64_t sfb200_tc_pretiled_floats(int N, int K);
int sfb200_tc_pretile(const float *W, float *Wt, int N, int K, void *stream);
int sfb200_linear_tc_ps(const float *x, const float *Wt, const float *bias, const float *residual, float *y, int M, int N, int K,
                        int act, void *stream);

/* One nn.Linear (optionally preceded by a LayerNorm over the K input features) through the persistent GEMM-chain kernel the
 * sampler's decode step uses for <= 64 rows (csrc/ar_chain.cu): y = act(LN?(x) W^T + bias) + residual, 3xTF32 on tcgen05 from
 * the fp32 weights (TMA-streamed), split-K reduced through L2 in a fixed order.  Replaces ln + nn.Linear of Block.forward
 * (transformer/mingpt.py:108-111) and of the heads (:222-231).  M <= 64, K % 32 == 0, K <= 4096; `workspace` =
 * sfb200_chain_workspace_bytes() bytes of device memory whose first 256 bytes are ZERO before the first call (the kernel
 * leaves them zero); ln_w / ln_b both NULL for a plain linear layer.  W must be 16-byte aligned. */
int64_t sfb200_chain_workspace_bytes(void);
int sfb200_chain_linear(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                        int act, const float *ln_w, const float *ln_b, void *workspace, void *stream);

/* Development aid: register a device buffer of 128 uint64 (or NULL to stop).  CTA 0 of every GEMM-chain launch adds the
 * nanoseconds spent in each stage of each phase to slot [phase * 8 + stage] and the visit count to slot [64 + phase * 8 +
 * stage] (stages: 0 wait-before barrier, 1 row statistics, 2 operand prefetch issue, 3 chunk loop, 4 partial store, 5 post-GEMM
 * barrier, 6 reduction; slot 62 = entry -> dependency wait).  The caller zeroes and reads the buffer. */
int sfb200_debug_chain_timeline(void *buf128);

/* Development aid: register a device buffer of 16 uint64; CTA (0,0) of sfb200_linear_tc_ps stamps %globaltimer (ns) at its
 * phase boundaries (slots documented in tc_gemm_ps.cu).  NULL unregisters. */
int sfb200_debug_ps_timeline(void *buf16);

/* LayerNorm over the last dim, eps = 1e-5 (mingpt.py:97-98,224). */
int sfb200_layernorm(const float *x, const float *w, const float *b, float *y, int rows, int d, void *stream);

/* Single-position causal attention with KV-cache append (CausalSelfAttention.forward mingpt.py:74-91, one new query).
 * qkv (B, 3d): [q | k | v] of the new position; kcache / vcache (B, H, max_len, 64); the new k,v are written at index
 * `pos` and the query attends to positions [0, pos].  pos is read from *pos_dev when pos_dev != NULL (graph replay).
 * out (B, d).  part: workspace for split-KV partials (B*H*n_split*(64+2) floats) or NULL when n_split == 1. */
int sfb200_attn_decode(const float *qkv, float *kcache, float *vcache, float *out, float *part, int B, int H, int max_len,
                       int pos, const int32_t *pos_dev, int n_split, void *stream);

/* sfb200_attn_decode for batches made of contiguous groups of `group` rows (2, 4, 6 or 8) whose first `shared_len` cached
 * positions hold identical K/V (the reference expands one conditioning to sample_n rows, shapeformer/shapeformer.py:229): the
 * shared prefix is read from the group's FIRST row once per (group, head) and scored against all `group` queries; every row's
 * own positions [shared_len, pos) and the new position (appended) follow per row.  TMA bulk copies into shared memory.
 * part: B*H*3*66 floats of scratch; counters: B*H int32, ZERO before the first call (left zero by every call). */
int sfb200_attn_decode_grouped(const float *qkv, float *kcache, float *vcache, float *out, float *part, int32_t *counters, int B,
                               int H, int max_len, int pos, int group, int shared_len, void *stream);

/* Causal attention over T new positions (prefill) + KV-cache fill.  qkv (B, T, 3d); writes k,v to cache [0,T). out (B,T,d) */
int sfb200_attn_prefill(const float *qkv, float *kcache, float *vcache, float *out, int B, int H, int T, int max_len,
                        void *stream);

/* sampling_masker + filter_sampling_logits + sample_logits for one tuple element (representers.py:120-155,
 * common.py:260-299).  logits (B,V); tokens (B,max_len,2); writes tokens[b][L][tuple_i]; hist_out (B,V) masked logits
 * or NULL.  noise_sample / noise_best (B,V).  L = current length (position being sampled), step_j = L - L_cond. */
int sfb200_ar_sample(const float *logits, int64_t *tokens, float *hist_out, const float *noise_sample,
                     const float *noise_best, int B, int V, int max_len, int L, int L_cond, int tuple_i,
                     const int64_t *end_tokens, const sfb200_ar_sampling *sp, void *stream);

/* ---- conv prologue of the decoder (csrc/conv_tc.cu): UNet3D + Upsampler of LocalDecoder (vqdif/dec.py:75-83,
 * vqdif/unet3d.py:449-474, vqdif/updown.py:79-132) on channels-last (N, D, H, W, C) fp32 tensors ------------------------------- */

/* taps = 8: the 3x3x3 convolution that FOLLOWS a nearest-neighbour x2 upsampling (updown.py:119-132) in sub-pixel form: `in` is the
 * low-resolution tensor (B,Z,Y,X,Cin), out is (B,2Z,2Y,2X,Cout); w packed [phase (pz,py,px)][tap (tz,ty,tz)][Cout][Cin] with the
 * original taps that read the same input voxel summed (decoder.py::pack_subpixel_weights).
 * 3x3x3 (taps = 27, padding 1) or 1x1x1 (taps = 1) convolution, no stride: out (B,Z,Y,X,Cout) = conv(in (B,Z,Y,X,Cin)) (+ bias)
 * (ReLU when relu != 0), tcgen05 3xTF32.  in_lo = low part of the operand split of `in` (dst_lo of sfb200_conv_prep);
 * w / w_lo = weights packed [tap][Cout][Cin] (tap = (dz*3 + dy)*3 + dx) and their low parts (sfb200_split_lo).  stats (optional):
 * (B, Cout, 2) doubles, += per-channel sum / sum of squares of the stored output.  Cin, Cout multiples of 32 (Cout 32, 64 or a
 * multiple of 128); Z, Y, X in {4, 8, 16, 32, 64, ...} (boxes of 128 voxels must tile the volume). */
int sfb200_conv3d_tc(const float *in, const float *in_lo, const float *w, const float *w_lo, const float *bias, float *out,
                     double *stats, int B, int Z, int Y, int X, int Cin, int Cout, int taps, int relu, void *stream);

/* Elementwise pass between convolutions: dst (B,Z,Y,X,C0+C1) = concat(src0 (C0 channels, read at coordinate >> sh0), src1 (C1
 * channels, >> sh1; may be NULL with C1 = 0)); with groups > 0 GroupNorm(groups, eps 1e-5, gamma, beta) from the per-channel sums
 * st0 / st1 ((B, C, 2) doubles over n0 / n1 voxels per sample); dst_lo (optional) = low part of the operand split of dst. */
int sfb200_conv_prep(const float *src0, int C0, int sh0, const double *st0, double n0, const float *src1, int C1, int sh1,
                     const double *st1, double n1, const float *gamma, const float *beta, int groups, float *dst, float *dst_lo,
                     int B, int Z, int Y, int X, void *stream);

/* dst (B,Zo,Yo,Xo,C) = max-pool win^3 (win 1 or 2) of src (B,Zo*win,Yo*win,Xo*win,C); stats += per-channel sums of dst.
 * dst may be NULL (statistics only). */
int sfb200_pool_stats(const float *src, float *dst, double *stats, int B, int Zo, int Yo, int Xo, int C, int win, void *stream);

/* Quantizer.get_code (vqdif/quantizer.py:19-30) into the channels-last layout: out (B, cells, C) = codebook[idx]; stats += sums. */
int sfb200_gather_codes_cl(const int64_t *idx, const float *codebook, float *out, double *stats, int B, int cells, int C,
                           int n_codes, void *stream);

/* ---- occupancy grid -> triangle mesh (csrc/mesh_kernels.cu): the step after decode_sample_indices in the reference,
 * geoutil.array2mesh (xgutils/geoutil.py:175-233) -> PyMCubes marching_cubes.  PyMCubes is not under the reference tree; this is
 * marching TETRAHEDRA on the Kuhn subdivision of every cube (same level set, linear interpolation, watertight, another
 * triangulation).  grid: (R, R, R) fp32 indexed [i][j][k]; vertices come out in grid-index coordinates.
 *   1. sfb200_mesh_mark_edges    flag (R^3 * 7) int32 = 1 where the level set crosses edge (vertex, direction)
 *   2. caller: vid = exclusive scan of flag, V = total
 *   3. sfb200_mesh_emit_vertices verts (V, 3) fp32
 *   4. sfb200_mesh_count_faces   count ((R-1)^3) int32 triangles per cube;  caller: foff = exclusive scan, F = total
 *   5. sfb200_mesh_emit_faces    faces (F, 3) int32, normals pointing from grid > thresh to grid <= thresh */
int sfb200_mesh_mark_edges(const float *grid, int R, float thresh, int32_t *flag, void *stream);
int sfb200_mesh_emit_vertices(const float *grid, int R, float thresh, const int32_t *flag, const int32_t *vid, float *verts,
                              void *stream);
int sfb200_mesh_count_faces(const float *grid, int R, float thresh, int32_t *count, void *stream);
int sfb200_mesh_emit_faces(const float *grid, int R, float thresh, const int32_t *vid, const int32_t *foff, const float *verts,
                           int32_t *faces, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SFB200_H */
