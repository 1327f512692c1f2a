// Host-side engine of the AR sampler: packed-weight layout, workspace carving, prefill, and the per-step kernel sequence
// (optionally replayed as a CUDA graph).  Replaces, for the hot path, the Python loop of ShapeFormer.sample_indices
// (shapeformer/shapeformer.py:72-115) and the uncached CondTupleGPT.sample_next_tuple (transformer/mingpt.py:297-310).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ar_kernels.cuh"

namespace sfb {

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Layout {
    int d, H, V[2], Ve, bs, nl[2];
    int64_t off[SFB200_W_COUNT];   // base offset of the first instance
    int64_t per_layer;             // floats per transformer block
    int64_t layer_base[2];         // offset of block (g, 0)
    int64_t head_base[2];
    int64_t total;
};

static int make_layout(const sfb200_ar_config *c, Layout *L) {
    if (!c) return SFB200_E_ARG;
    if (c->n_embd <= 0 || c->n_embd % 64 != 0 || c->n_head <= 0 || c->n_embd != c->n_head * 64) return SFB200_E_ARG;
    if (c->n_layers[0] < 1 || c->n_layers[1] < 1 || c->block_size < 2) return SFB200_E_ARG;
    if (c->vocab[0] < 2 || c->vocab[1] < 2 || c->vocab[0] > 8192 || c->vocab[1] > 8192 || c->extra_vocab < 1)
        return SFB200_E_ARG;
    L->d = c->n_embd; L->H = c->n_head; L->V[0] = c->vocab[0]; L->V[1] = c->vocab[1]; L->Ve = c->extra_vocab;
    L->bs = c->block_size; L->nl[0] = c->n_layers[0]; L->nl[1] = c->n_layers[1];
    const int64_t d = L->d;
    int64_t o = 0;
    L->off[SFB200_W_POS_EMB] = o; o += (int64_t)L->bs * d;
    L->off[SFB200_W_COND_POS_EMB] = o; o += (int64_t)L->bs * d;
    L->off[SFB200_W_TOK_EMB0] = o; o += (int64_t)L->V[0] * d;
    L->off[SFB200_W_TOK_EMB1] = o; o += (int64_t)L->V[1] * d;
    L->off[SFB200_W_EXTRA_EMB] = o; o += (int64_t)L->Ve * d;
    for (int g = 0; g < 2; ++g) {
        L->head_base[g] = o;
        o += 2 * d + (int64_t)L->V[g] * d;
    }
    // within a block
    int64_t p = 0;
    L->off[SFB200_W_LN1_W] = p; p += d;
    L->off[SFB200_W_LN1_B] = p; p += d;
    L->off[SFB200_W_QKV_W] = p; p += 3 * d * d;
    L->off[SFB200_W_QKV_B] = p; p += 3 * d;
    L->off[SFB200_W_PROJ_W] = p; p += d * d;
    L->off[SFB200_W_PROJ_B] = p; p += d;
    L->off[SFB200_W_LN2_W] = p; p += d;
    L->off[SFB200_W_LN2_B] = p; p += d;
    L->off[SFB200_W_FC1_W] = p; p += 4 * d * d;
    L->off[SFB200_W_FC1_B] = p; p += 4 * d;
    L->off[SFB200_W_FC2_W] = p; p += 4 * d * d;
    L->off[SFB200_W_FC2_B] = p; p += d;
    L->per_layer = p;
    L->layer_base[0] = o; o += p * L->nl[0];
    L->layer_base[1] = o; o += p * L->nl[1];
    L->total = o;
    return SFB200_OK;
}

static int64_t weight_offset(const Layout *L, int id, int g, int l) {
    if (id < 0 || id >= SFB200_W_COUNT) return -1;
    if (id <= SFB200_W_EXTRA_EMB) return L->off[id];
    if (g < 0 || g > 1) return -1;
    if (id == SFB200_W_HEAD_LN_W) return L->head_base[g];
    if (id == SFB200_W_HEAD_LN_B) return L->head_base[g] + L->d;
    if (id == SFB200_W_HEAD_W) return L->head_base[g] + 2 * L->d;
    if (l < 0 || l >= L->nl[g]) return -1;
    return L->layer_base[g] + (int64_t)l * L->per_layer + L->off[id];
}

struct Buffers {   // byte offsets into the workspace
    int64_t st, rowmap, x0, x1, h, qkv, att, ff, logits[2], part;
    int64_t px, px1, ph, pqkv, pff;
    int64_t logp;                           // (max_rows, max_steps, 2) log-probabilities of the sampled tokens
    int64_t att_cnt;                        // grouped attention: arrival counters per (row, head)
    int64_t ch_bar, ch_stats, ch_scratch;   // GEMM-chain kernel: barrier counters, LayerNorm row statistics, split-K partials
    // large-M GEMM (tc_big.cu): lo parts of the GEMM input activations (decode: h, att, ff; prefill: ph, pff), split-K scratch
    int64_t h_lo, att_lo, ff_lo, ph_lo, pff_lo, bg_part, bg_cnt;
    int64_t total;
};

constexpr int CHAIN_MAX_GRID = 160;    // CTAs the chain scratch is sized for (one per SM; B200 has 148)

static int pick_nsplit(int B, int H) {
    int n = (2 * 148 + B * H - 1) / (B * H);
    if (n < 1) n = 1;
    if (n > 8) n = 8;
    return n;
}

static void carve(const sfb200_ar_config *c, Buffers *b) {
    const int64_t d = c->n_embd, B = c->max_rows, F = sizeof(float);
    const int64_t Vmax = c->vocab[0] > c->vocab[1] ? c->vocab[0] : c->vocab[1];
    const int64_t P = (int64_t)c->prefill_rows * c->max_cond;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align_up(o + bytes, 256); return r; };
    b->st = take(ST_WORDS * 4);
    b->rowmap = take(3 * B * 4);    // prefill leaders | duplicate rows | their sources
    b->x0 = take(B * d * F);
    b->x1 = take(B * d * F);
    b->h = take(B * d * F);
    b->qkv = take(B * 3 * d * F);
    b->att = take(B * d * F);
    b->ff = take(B * 4 * d * F);
    b->logits[0] = take(B * Vmax * F);
    b->logits[1] = take(B * Vmax * F);
    b->part = take(B * c->n_head * 8 * 66 * F);
    b->px = take(P * d * F);
    b->px1 = take(P * d * F);
    b->ph = take(P * d * F);
    b->pqkv = take(P * 3 * d * F);
    b->pff = take(P * 4 * d * F);
    b->logp = take(B * (int64_t)c->max_steps * 2 * F);
    b->att_cnt = take(B * c->n_head * 4);
    b->ch_bar = take((CH_MAX_BARRIERS + 1) * 4);
    b->ch_stats = take(64 * ((d + 127) / 128) * 2 * F);
    b->ch_scratch = take((int64_t)chain_scratch_floats(CHAIN_MAX_GRID) * F);
    b->h_lo = take(B * d * F);
    b->att_lo = take(B * d * F);
    b->ff_lo = take(B * 4 * d * F);
    b->ph_lo = take(P * d * F);
    b->pff_lo = take(P * 4 * d * F);
    b->bg_part = take((int64_t)big_partial_floats() * F);
    b->bg_cnt = take(2 * BG_MAX_TILES * 4);
    b->total = o;
}

}  // namespace sfb

using namespace sfb;

struct sfb200_ar {
    sfb200_ar_config cfg;
    Layout lay;
    Buffers buf;
    const float *w;
    float *wt;                 // optional pre-split GEMM weight tiles (sfb200_ar_set_pretiled)
    const float *wlo;          // optional lo parts of the weight blob, same layout (sfb200_ar_set_lo_weights): GEMMs over more
                               // than 64 rows (large decode batches, prefill) then run the TMA-fed tc_big kernel
    char *ws;
    float *kv;
    int64_t *tokens;
    float *hist;
    int Vmax;
    // per batch
    int B, L_cond, n_split;
    int attn_group;            // > 1: contiguous groups of this many rows share their conditioning prefix K/V
    bool begun;
    sfb200_ar_sampling sp;
    // graph
    cudaGraphExec_t gexec;
    cudaStream_t cap_stream;   // private stream used only to record the step graph (the legacy default stream cannot capture)
    const float *g_noise;
    int g_B, g_group;
    sfb200_ar_sampling g_sp;
    long long g_nodes;         // kernel nodes in the captured step
    bool capturing;
    // attention profiling (bench.py roofline): events bracketing every attention launch of eager steps
    bool prof;
    std::vector<cudaEvent_t> *ev;
    size_t ev_used;
    double prof_bytes;         // bytes that must move (conditioning prefix counted once per group when shared)
    double prof_bytes_per_row; // SURVEY §8d per-row formula (no sharing)
    int steps_host;            // steps enqueued since sfb200_ar_begin (host mirror of st[ST_STEPS])
    // pinned staging for the per-batch row lists (leaders | duplicate rows | their sources), owned by the handle so that
    // sfb200_ar_begin_shared never has to synchronise the stream; stage_ev marks the last copy out of it
    int32_t *h_stage;
    cudaEvent_t stage_ev;
    bool stage_busy;
    // GEMM-chain kernel (ar_chain.cu): one TMA tensor map per GEMM weight, built on first use
    TensorMapBlob *maps;       // [n_blocks][QKV, PROJ, FC1, FC2] then the two heads
    bool maps_ready;
    int use_chain;             // decode steps of <= 64 rows run the persistent chain kernel (SFB200_CHAIN=0 disables)
    int use_grouped_attn;      // contiguous groups of identical conditionings use attn_grouped.cu (SFB200_ATTN_GROUPED=0 disables)
};

static inline const float *W_(const sfb200_ar *h, int id, int g, int l) { return h->w + weight_offset(&h->lay, id, g, l); }
template <typename T>
static inline T *WS_(const sfb200_ar *h, int64_t off) { return reinterpret_cast<T *>(h->ws + off); }

// K cache of block (g, l): (max_rows, H, max_len, 64); V cache follows K.
static inline float *kcache(const sfb200_ar *h, int g, int l, int row0) {
    const int64_t per = (int64_t)h->cfg.max_rows * h->cfg.n_head * h->cfg.max_len * 64;
    const int64_t li = (g == 0 ? 0 : h->cfg.n_layers[0]) + l;
    return h->kv + li * 2 * per + (int64_t)row0 * h->cfg.n_head * h->cfg.max_len * 64;
}
static inline float *vcache(const sfb200_ar *h, int g, int l, int row0) {
    const int64_t per = (int64_t)h->cfg.max_rows * h->cfg.n_head * h->cfg.max_len * 64;
    return kcache(h, g, l, row0) + per;
}

extern "C" {

int64_t sfb200_ar_weight_floats(const sfb200_ar_config *cfg) {
    Layout L;
    if (make_layout(cfg, &L) != SFB200_OK) return -1;
    return L.total;
}
int64_t sfb200_ar_weight_offset(const sfb200_ar_config *cfg, int tensor_id, int group, int layer) {
    Layout L;
    if (make_layout(cfg, &L) != SFB200_OK) return -1;
    return weight_offset(&L, tensor_id, group, layer);
}
int64_t sfb200_ar_kv_bytes(const sfb200_ar_config *cfg) {
    if (!cfg) return -1;
    return (int64_t)(cfg->n_layers[0] + cfg->n_layers[1]) * 2 * cfg->max_rows * cfg->n_head * cfg->max_len * 64 * 4;
}
int64_t sfb200_ar_workspace_bytes(const sfb200_ar_config *cfg) {
    if (!cfg) return -1;
    Buffers b;
    carve(cfg, &b);
    return b.total;
}
int64_t sfb200_ar_history_floats(const sfb200_ar_config *cfg) {
    if (!cfg) return -1;
    if (!cfg->keep_history) return 0;
    return (int64_t)cfg->max_rows * cfg->max_steps * ((int64_t)cfg->vocab[0] + cfg->vocab[1]);
}

int sfb200_ar_create(const sfb200_ar_config *cfg, const float *weights, void *kv_cache, void *workspace, int64_t *tokens,
                     float *history, sfb200_ar **out) {
    if (!cfg || !weights || !kv_cache || !workspace || !tokens || !out) return SFB200_E_ARG;
    if (cfg->keep_history && !history) return SFB200_E_ARG;
    if (cfg->max_rows < 1 || cfg->max_len < 2 || cfg->max_steps < 1 || cfg->prefill_rows < 1 || cfg->max_cond < 1)
        return SFB200_E_ARG;
    if (cfg->max_len > cfg->block_size || cfg->max_cond >= cfg->max_len || cfg->prefill_rows > cfg->max_rows)
        return SFB200_E_ARG;
    sfb200_ar *h = static_cast<sfb200_ar *>(calloc(1, sizeof(sfb200_ar)));
    if (!h) return SFB200_E_ARG;
    h->cfg = *cfg;
    int r = make_layout(cfg, &h->lay);
    if (r != SFB200_OK) { free(h); return r; }
    carve(cfg, &h->buf);
    h->w = weights;
    h->ws = static_cast<char *>(workspace);
    h->kv = static_cast<float *>(kv_cache);
    h->tokens = tokens;
    h->hist = cfg->keep_history ? history : nullptr;
    h->Vmax = cfg->vocab[0] > cfg->vocab[1] ? cfg->vocab[0] : cfg->vocab[1];
    h->begun = false;
    h->gexec = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void **>(&h->h_stage), sizeof(int32_t) * 3 * cfg->max_rows, cudaHostAllocDefault) !=
            cudaSuccess ||
        cudaEventCreateWithFlags(&h->stage_ev, cudaEventDisableTiming) != cudaSuccess) {
        set_cuda_error(cudaGetLastError(), "sfb200_ar_create: pinned staging");
        if (h->h_stage) cudaFreeHost(h->h_stage);
        free(h);
        return SFB200_E_CUDA;
    }
    {
        const char *e = getenv("SFB200_CHAIN");
        h->use_chain = (e && e[0] == '0') ? 0 : 1;
        e = getenv("SFB200_ATTN_GROUPED");
        h->use_grouped_attn = (e && e[0] == '0') ? 0 : 1;
    }
    *out = h;
    return SFB200_OK;
}

void sfb200_ar_destroy(sfb200_ar *h) {
    if (!h) return;
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->maps) free(h->maps);
    if (h->stage_ev) cudaEventDestroy(h->stage_ev);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->ev) {
        for (cudaEvent_t e : *h->ev) cudaEventDestroy(e);
        delete h->ev;
    }
    free(h);
}

const int32_t *sfb200_ar_status_ptr(const sfb200_ar *h) { return h ? WS_<int32_t>(h, h->buf.st) : nullptr; }
const float *sfb200_ar_logprob_ptr(const sfb200_ar *h) { return h ? WS_<float>(h, h->buf.logp) : nullptr; }

}  // extern "C"

// Pre-split tile blob: per block [QKV | PROJ | FC1 | FC2], then the two heads.
static int64_t wt_block_floats(const sfb200_ar_config *c) {
    const int d = c->n_embd;
    return tc_pretiled_floats(3 * d, d) + tc_pretiled_floats(d, d) + tc_pretiled_floats(4 * d, d) + tc_pretiled_floats(d, 4 * d);
}
static int64_t wt_offset(const sfb200_ar_config *c, int id, int g, int l) {
    const int d = c->n_embd;
    const int64_t nblocks = c->n_layers[0] + c->n_layers[1];
    if (id == SFB200_W_HEAD_W) return nblocks * wt_block_floats(c) + (g == 1 ? tc_pretiled_floats(c->vocab[0], d) : 0);
    int64_t o = ((g == 0 ? 0 : c->n_layers[0]) + l) * wt_block_floats(c);
    if (id == SFB200_W_QKV_W) return o;
    o += tc_pretiled_floats(3 * d, d);
    if (id == SFB200_W_PROJ_W) return o;
    o += tc_pretiled_floats(d, d);
    if (id == SFB200_W_FC1_W) return o;
    o += tc_pretiled_floats(4 * d, d);
    if (id == SFB200_W_FC2_W) return o;
    return -1;
}

// nn.Linear dispatch: 9..64 rows -> tcgen05 3xTF32 from pre-split tiles (when bound), else tcgen05 with in-kernel split;
// <= 8 rows -> GEMV / FFMA kernels on the fp32 weights.
static inline bool big_path(const sfb200_ar *h, int M) { return h->wlo != nullptr && M > 64; }

static int linear(sfb200_ar *h, int wid, int g, int l, const float *x, const float *bias, const float *residual, float *y,
                  int M, int N, int K, int act, cudaStream_t s, const float *x_lo = nullptr, float *y_lo = nullptr) {
    if (big_path(h, M) && x_lo)
        return launch_linear_big(x, x_lo, W_(h, wid, g, l), h->wlo + weight_offset(&h->lay, wid, g, l), bias, residual, y, y_lo, M, N,
                                 K, act, WS_<float>(h, h->buf.bg_part), WS_<int>(h, h->buf.bg_cnt), s);
    if (M >= 9 && M <= 64 && h->wt)
        return launch_linear_tc_ps(x, h->wt + wt_offset(&h->cfg, wid, g, l), bias, residual, y, M, N, K, act, s);
    const float *W = W_(h, wid, g, l);
    if (M >= 9) return launch_linear_tc(x, W, bias, residual, y, M, N, K, act, s);
    return launch_linear(x, W, bias, residual, y, M, N, K, act, s);
}

// One transformer block over M = rows*T positions (prefill) — Block.forward, transformer/mingpt.py:108-111.
static int block_prefill(sfb200_ar *h, int g, int l, float *x, int row0, int rows, int T, cudaStream_t s,
                         const int32_t *rowmap) {
    const int d = h->cfg.n_embd, H = h->cfg.n_head, M = rows * T;
    float *ph = WS_<float>(h, h->buf.ph), *pqkv = WS_<float>(h, h->buf.pqkv), *pff = WS_<float>(h, h->buf.pff);
    // tc_big GEMMs read the lo parts of their input activations from memory: written by the producing kernel where that is
    // one of ours with an epilogue (LayerNorm, GEMM), by split_lo for the attention output
    const bool big = big_path(h, M);
    float *ph_lo = big ? WS_<float>(h, h->buf.ph_lo) : nullptr, *pff_lo = big ? WS_<float>(h, h->buf.pff_lo) : nullptr;
    SFB_TRY(launch_layernorm(x, W_(h, SFB200_W_LN1_W, g, l), W_(h, SFB200_W_LN1_B, g, l), ph, M, d, s, ph_lo));
    SFB_TRY(linear(h, SFB200_W_QKV_W, g, l, ph, W_(h, SFB200_W_QKV_B, g, l), nullptr, pqkv, M, 3 * d, d, 0, s, ph_lo));
    // with a rowmap the cache rows are rowmap[i] (cache base = row 0), otherwise rows row0 .. row0+rows-1
    SFB_TRY(launch_attn_prefill(pqkv, kcache(h, g, l, rowmap ? 0 : row0), vcache(h, g, l, rowmap ? 0 : row0), ph, rows, H, T,
                                h->cfg.max_len, s, rowmap));
    if (big) SFB_TRY(launch_split_lo(ph, ph_lo, (size_t)M * d, s));
    SFB_TRY(linear(h, SFB200_W_PROJ_W, g, l, ph, W_(h, SFB200_W_PROJ_B, g, l), x, x, M, d, d, 0, s, ph_lo));
    SFB_TRY(launch_layernorm(x, W_(h, SFB200_W_LN2_W, g, l), W_(h, SFB200_W_LN2_B, g, l), ph, M, d, s, ph_lo));
    SFB_TRY(linear(h, SFB200_W_FC1_W, g, l, ph, W_(h, SFB200_W_FC1_B, g, l), nullptr, pff, M, 4 * d, d, 1, s, ph_lo, pff_lo));
    SFB_TRY(linear(h, SFB200_W_FC2_W, g, l, pff, W_(h, SFB200_W_FC2_B, g, l), x, x, M, d, 4 * d, 0, s, pff_lo));
    return SFB200_OK;
}

// Decode attention of block (g, l) for the newest position (read from the device state), bracketed by events when the
// bench's attention profiling is on.
static int attn_step(sfb200_ar *h, int g, int l, cudaStream_t s, float *att_lo = nullptr) {
    const int d = h->cfg.n_embd, H = h->cfg.n_head, B = h->B;
    const int32_t *st = WS_<int32_t>(h, h->buf.st);
    float *qkv = WS_<float>(h, h->buf.qkv), *att = WS_<float>(h, h->buf.att), *part = WS_<float>(h, h->buf.part);
    const bool timed = h->prof && !h->capturing;
    if (timed) {
        if (!h->ev) h->ev = new std::vector<cudaEvent_t>();
        while (h->ev->size() < h->ev_used + 2) {
            cudaEvent_t e;
            SFB_CUDA_TRY(cudaEventCreate(&e));
            h->ev->push_back(e);
        }
        SFB_CUDA_TRY(cudaEventRecord((*h->ev)[h->ev_used], s));
    }
    // att_lo: low part of the attention output for a tc_big projection — written by the grouped kernel's combine, by a
    // split_lo pass after the per-row kernel
    if (h->attn_group > 1 && h->use_grouped_attn) {
        SFB_TRY(launch_attn_grouped(qkv, kcache(h, g, l, 0), vcache(h, g, l, 0), att, part, WS_<int>(h, h->buf.att_cnt), B, H,
                                    h->cfg.max_len, 0, st, h->attn_group, 0, g == 1 ? -1 : 0, s, att_lo));
    } else {
        SFB_TRY(launch_attn_decode(qkv, kcache(h, g, l, 0), vcache(h, g, l, 0), att, part, B, H, h->cfg.max_len, 0, st,
                                   h->n_split, s, h->attn_group, 0, g == 1 ? -1 : 0));
        if (att_lo) SFB_TRY(launch_split_lo(att, att_lo, (size_t)B * d, s));
    }
    if (timed) {
        SFB_CUDA_TRY(cudaEventRecord((*h->ev)[h->ev_used + 1], s));
        h->ev_used += 2;
        // position of this launch = L-1: blocks[1] runs before the step's advance (L = L_cond + j), blocks[0] after it
        // (L = L_cond + j + 1, and steps_host has already been incremented)
        const double pos = (double)(h->L_cond + h->steps_host) - 1.0;
        // key rows that must be read: per row `pos`, or with prefix sharing the shared prefix once per group + own keys
        const double lc = (double)h->L_cond - (g == 1 ? 1.0 : 0.0);
        const double shared = pos < lc ? pos : lc;
        const double key_rows = h->attn_group > 1 ? (double)(B / h->attn_group) * shared + (double)B * (pos - shared)
                                                  : (double)B * pos;
        h->prof_bytes += key_rows * 2.0 * d * 4.0 + (double)B * 4.0 * d * 4.0;
        h->prof_bytes_per_row += (double)B * (2.0 * pos * d * 4.0 + 4.0 * d * 4.0);
    }
    return SFB200_OK;
}

// One transformer block for the newest position of every row (decode) as separate kernels (rows > 64, or SFB200_CHAIN=0).
static int block_step(sfb200_ar *h, int g, int l, float *x, cudaStream_t s) {
    const int d = h->cfg.n_embd, B = h->B;
    float *hb = WS_<float>(h, h->buf.h), *qkv = WS_<float>(h, h->buf.qkv), *att = WS_<float>(h, h->buf.att);
    float *ff = WS_<float>(h, h->buf.ff);
    const bool big = big_path(h, B);
    float *h_lo = big ? WS_<float>(h, h->buf.h_lo) : nullptr, *att_lo = big ? WS_<float>(h, h->buf.att_lo) : nullptr;
    float *ff_lo = big ? WS_<float>(h, h->buf.ff_lo) : nullptr;
    SFB_TRY(launch_layernorm(x, W_(h, SFB200_W_LN1_W, g, l), W_(h, SFB200_W_LN1_B, g, l), hb, B, d, s, h_lo));
    SFB_TRY(linear(h, SFB200_W_QKV_W, g, l, hb, W_(h, SFB200_W_QKV_B, g, l), nullptr, qkv, B, 3 * d, d, 0, s, h_lo));
    SFB_TRY(attn_step(h, g, l, s, att_lo));
    SFB_TRY(linear(h, SFB200_W_PROJ_W, g, l, att, W_(h, SFB200_W_PROJ_B, g, l), x, x, B, d, d, 0, s, att_lo));
    SFB_TRY(launch_layernorm(x, W_(h, SFB200_W_LN2_W, g, l), W_(h, SFB200_W_LN2_B, g, l), hb, B, d, s, h_lo));
    SFB_TRY(linear(h, SFB200_W_FC1_W, g, l, hb, W_(h, SFB200_W_FC1_B, g, l), nullptr, ff, B, 4 * d, d, 1, s, h_lo, ff_lo));
    SFB_TRY(linear(h, SFB200_W_FC2_W, g, l, ff, W_(h, SFB200_W_FC2_B, g, l), x, x, B, d, 4 * d, 0, s, ff_lo));
    return SFB200_OK;
}

// ---- persistent GEMM-chain path -------------------------------------------------------------------------------------
static int chain_map_index(const sfb200_ar *h, int wid, int g, int l) {
    const int nb = h->cfg.n_layers[0] + h->cfg.n_layers[1];
    if (wid == SFB200_W_HEAD_W) return nb * 4 + g;
    const int li = (g == 0 ? 0 : h->cfg.n_layers[0]) + l;
    const int k = wid == SFB200_W_QKV_W ? 0 : wid == SFB200_W_PROJ_W ? 1 : wid == SFB200_W_FC1_W ? 2 : 3;
    return li * 4 + k;
}

static int chain_build_maps(sfb200_ar *h) {
    if (h->maps_ready) return SFB200_OK;
    const int d = h->cfg.n_embd, nb = h->cfg.n_layers[0] + h->cfg.n_layers[1];
    if (!h->maps) {
        h->maps = static_cast<TensorMapBlob *>(aligned_alloc(64, sizeof(TensorMapBlob) * (size_t)(nb * 4 + 2)));
        if (!h->maps) return SFB200_E_ARG;
    }
    for (int g = 0; g < 2; ++g) {
        for (int l = 0; l < h->cfg.n_layers[g]; ++l) {
            SFB_TRY(chain_weight_map(W_(h, SFB200_W_QKV_W, g, l), 3 * d, d, &h->maps[chain_map_index(h, SFB200_W_QKV_W, g, l)]));
            SFB_TRY(chain_weight_map(W_(h, SFB200_W_PROJ_W, g, l), d, d, &h->maps[chain_map_index(h, SFB200_W_PROJ_W, g, l)]));
            SFB_TRY(chain_weight_map(W_(h, SFB200_W_FC1_W, g, l), 4 * d, d, &h->maps[chain_map_index(h, SFB200_W_FC1_W, g, l)]));
            SFB_TRY(chain_weight_map(W_(h, SFB200_W_FC2_W, g, l), d, 4 * d, &h->maps[chain_map_index(h, SFB200_W_FC2_W, g, l)]));
        }
        SFB_TRY(chain_weight_map(W_(h, SFB200_W_HEAD_W, g, 0), h->cfg.vocab[g], d, &h->maps[chain_map_index(h, SFB200_W_HEAD_W, g, 0)]));
    }
    h->maps_ready = true;
    return SFB200_OK;
}

static bool chain_active(const sfb200_ar *h) { return h->use_chain && h->B <= 64 && chain_grid_size() > 0; }

static void chain_gemm(sfb200_ar *h, ChainArgs &a, int wid, int g, int l, const float *x, const float *ln_g, const float *ln_b,
                       const float *bias, const float *residual, float *y, bool stats_out, int N, int K, int act,
                       int wait_before) {
    const int p = a.n_phases++;
    int grid = chain_grid_size();
    if (grid > CHAIN_MAX_GRID) grid = CHAIN_MAX_GRID;
    ChainPhase &ph = a.ph[p];
    a.wmap[p] = h->maps[chain_map_index(h, wid, g, l)];
    ph.x = x; ph.ln_g = ln_g; ph.ln_b = ln_b;
    ph.stats_in = ln_g ? WS_<float>(h, h->buf.ch_stats) : nullptr;
    ph.bias = bias; ph.residual = residual; ph.y = y;
    ph.stats_out = stats_out ? WS_<float>(h, h->buf.ch_stats) : nullptr;
    ph.N = N; ph.K = K; ph.act = act; ph.wait_before = wait_before;
    chain_plan(N, K, grid, &ph.tiles, &ph.splits);
}

// All blocks of group g (+ its head) for the newest position: n + 1 chain launches interleaved with n attention launches.
static int group_step_chain(sfb200_ar *h, int g, float *x, float *logits, cudaStream_t s) {
    const int d = h->cfg.n_embd, nl = h->cfg.n_layers[g];
    float *qkv = WS_<float>(h, h->buf.qkv), *att = WS_<float>(h, h->buf.att), *ff = WS_<float>(h, h->buf.ff);
    ChainArgs a;
    auto reset = [&]() {
        memset(&a, 0, sizeof(a));
        a.M = h->B;
        a.scratch = WS_<float>(h, h->buf.ch_scratch);
        a.bar = WS_<unsigned int>(h, h->buf.ch_bar);
    };
    // launch 0: row statistics of x (written by the embedding kernels) -> LN1 -> QKV of the first block
    reset();
    {
        ChainPhase &ph = a.ph[a.n_phases++];
        ph.residual = x; ph.stats_out = WS_<float>(h, h->buf.ch_stats); ph.N = d; ph.tiles = 0; ph.splits = 0;
    }
    chain_gemm(h, a, SFB200_W_QKV_W, g, 0, x, W_(h, SFB200_W_LN1_W, g, 0), W_(h, SFB200_W_LN1_B, g, 0),
               W_(h, SFB200_W_QKV_B, g, 0), nullptr, qkv, false, 3 * d, d, 0, 1);
    SFB_TRY(launch_chain(a, s));
    for (int l = 0; l < nl; ++l) {
        SFB_TRY(attn_step(h, g, l, s));
        reset();
        chain_gemm(h, a, SFB200_W_PROJ_W, g, l, att, nullptr, nullptr, W_(h, SFB200_W_PROJ_B, g, l), x, x, true, d, d, 0, 0);
        chain_gemm(h, a, SFB200_W_FC1_W, g, l, x, W_(h, SFB200_W_LN2_W, g, l), W_(h, SFB200_W_LN2_B, g, l),
                   W_(h, SFB200_W_FC1_B, g, l), nullptr, ff, false, 4 * d, d, 1, 1);
        chain_gemm(h, a, SFB200_W_FC2_W, g, l, ff, nullptr, nullptr, W_(h, SFB200_W_FC2_B, g, l), x, x, true, d, 4 * d, 0, 1);
        if (l + 1 < nl)
            chain_gemm(h, a, SFB200_W_QKV_W, g, l + 1, x, W_(h, SFB200_W_LN1_W, g, l + 1), W_(h, SFB200_W_LN1_B, g, l + 1),
                       W_(h, SFB200_W_QKV_B, g, l + 1), nullptr, qkv, false, 3 * d, d, 0, 1);
        else
            chain_gemm(h, a, SFB200_W_HEAD_W, g, 0, x, W_(h, SFB200_W_HEAD_LN_W, g, 0), W_(h, SFB200_W_HEAD_LN_B, g, 0), nullptr,
                       nullptr, logits, false, h->cfg.vocab[g], d, 0, 1);
        SFB_TRY(launch_chain(a, s));
    }
    return SFB200_OK;
}

static int group_step(sfb200_ar *h, int g, float *x, float *logits, cudaStream_t s);

static int head(sfb200_ar *h, int g, const float *x, float *logits, int rows, cudaStream_t s) {
    const int d = h->cfg.n_embd;
    float *hb = WS_<float>(h, h->buf.h);
    float *h_lo = big_path(h, rows) ? WS_<float>(h, h->buf.h_lo) : nullptr;
    SFB_TRY(launch_layernorm(x, W_(h, SFB200_W_HEAD_LN_W, g, 0), W_(h, SFB200_W_HEAD_LN_B, g, 0), hb, rows, d, s, h_lo));
    SFB_TRY(linear(h, SFB200_W_HEAD_W, g, 0, hb, nullptr, nullptr, logits, rows, h->cfg.vocab[g], d, 0, s, h_lo));
    return SFB200_OK;
}

static int group_step(sfb200_ar *h, int g, float *x, float *logits, cudaStream_t s) {
    if (chain_active(h)) return group_step_chain(h, g, x, logits, s);
    for (int l = 0; l < h->cfg.n_layers[g]; ++l) SFB_TRY(block_step(h, g, l, x, s));
    return head(h, g, x, logits, h->B, s);
}

extern "C" int sfb200_ar_begin_shared(sfb200_ar *h, int B, int L_cond, const sfb200_ar_sampling *sp, const int32_t *row_src,
                                      void *stream) {
    if (!h || !sp) return SFB200_E_ARG;
    if (B < 1 || B > h->cfg.max_rows || L_cond < 1 || L_cond > h->cfg.max_cond) return SFB200_E_ARG;
    if (!(sp->temperature > 0.f)) return SFB200_E_ARG;
    cudaStream_t s = as_stream(stream);
    const int d = h->cfg.n_embd;
    h->B = B; h->L_cond = L_cond; h->sp = *sp;
    h->n_split = pick_nsplit(B, h->cfg.n_head);
    h->steps_host = 0;
    int32_t *st = WS_<int32_t>(h, h->buf.st);
    SFB_TRY(launch_state_init(st, L_cond, s));
    SFB_CUDA_TRY(cudaMemsetAsync(WS_<int>(h, h->buf.att_cnt), 0, (size_t)h->cfg.max_rows * h->cfg.n_head * 4, s));
    SFB_CUDA_TRY(cudaMemsetAsync(WS_<int>(h, h->buf.bg_cnt), 0, 2 * BG_MAX_TILES * 4, s));
    if (chain_active(h)) {
        SFB_TRY(chain_build_maps(h));
        SFB_CUDA_TRY(cudaMemsetAsync(WS_<unsigned int>(h, h->buf.ch_bar), 0, (CH_MAX_BARRIERS + 1) * 4, s));
    }
    // ---- group leaders / duplicates
    std::vector<int32_t> lead, dup_dst, dup_src;
    for (int b = 0; b < B; ++b) {
        const int src = row_src ? row_src[b] : b;
        if (src < 0 || src > b || (row_src && row_src[src] != src)) return SFB200_E_ARG;
        if (src == b) lead.push_back(b); else { dup_dst.push_back(b); dup_src.push_back(src); }
    }
    const int n_lead = (int)lead.size(), n_dup = (int)dup_dst.size();
    // contiguous groups of equal size (the reference's c_indices.expand(sample_n)) enable the grouped attention kernel
    h->attn_group = 1;
    for (int G = 8; G >= 2; G -= 2) {
        if (B % G != 0 || !row_src) continue;
        bool ok = true;
        for (int b = 0; b < B && ok; ++b) ok = row_src[b] == (b / G) * G;
        if (ok) { h->attn_group = G; break; }
    }
    int32_t *rm = WS_<int32_t>(h, h->buf.rowmap);
    const int32_t *rm_lead = nullptr, *rm_dst = rm + h->cfg.max_rows, *rm_src = rm + 2 * h->cfg.max_rows;
    if (n_dup > 0) {
        // the row lists go through the handle's pinned staging buffer: no stream synchronisation.  The only wait is for the
        // PREVIOUS batch's copy out of the same buffer, which finished long ago unless batches are begun back to back.
        if (h->stage_busy) SFB_CUDA_TRY(cudaEventSynchronize(h->stage_ev));
        const int mr = h->cfg.max_rows;
        memcpy(h->h_stage, lead.data(), sizeof(int32_t) * n_lead);
        memcpy(h->h_stage + mr, dup_dst.data(), sizeof(int32_t) * n_dup);
        memcpy(h->h_stage + 2 * mr, dup_src.data(), sizeof(int32_t) * n_dup);
        SFB_CUDA_TRY(cudaMemcpyAsync(rm, h->h_stage, sizeof(int32_t) * 3 * mr, cudaMemcpyHostToDevice, s));
        SFB_CUDA_TRY(cudaEventRecord(h->stage_ev, s));
        h->stage_busy = true;
        rm_lead = rm;
    }
    float *px = WS_<float>(h, h->buf.px), *px1 = WS_<float>(h, h->buf.px1), *x0 = WS_<float>(h, h->buf.x0);
    const int64_t end0 = h->cfg.end_tokens[0];
    for (int r0 = 0; r0 < n_lead; r0 += h->cfg.prefill_rows) {
        const int rows = (n_lead - r0 < h->cfg.prefill_rows) ? n_lead - r0 : h->cfg.prefill_rows;
        // without sharing the chunk is rows r0.. of every buffer; with sharing a row list selects tokens / cache / x0 rows
        const int32_t *map = rm_lead ? rm_lead + r0 : nullptr;
        const int64_t *tok = h->tokens + (map ? 0 : (int64_t)r0 * h->cfg.max_len * 2);
        SFB_TRY(launch_embed(tok, W_(h, SFB200_W_TOK_EMB0, 0, 0), W_(h, SFB200_W_TOK_EMB1, 0, 0),
                             W_(h, SFB200_W_EXTRA_EMB, 0, 0), W_(h, SFB200_W_POS_EMB, 0, 0),
                             W_(h, SFB200_W_COND_POS_EMB, 0, 0), px, rows, d, h->cfg.max_len, 0, L_cond, L_cond, end0,
                             nullptr, s, map));
        for (int l = 0; l < h->cfg.n_layers[0]; ++l) SFB_TRY(block_prefill(h, 0, l, px, r0, rows, L_cond, s, map));
        SFB_TRY(launch_take_last(px, map ? x0 : x0 + (int64_t)r0 * d, rows, d, L_cond, s, map));
        const int T1 = L_cond - 1;
        if (T1 > 0) {
            // blocks[1] input of position t is blocks[0] output + tok_embs[0](pos of tuple t+1)   (mingpt.py:309)
            SFB_TRY(launch_add_target(px, px1, tok, W_(h, SFB200_W_TOK_EMB0, 0, 0), rows, d, h->cfg.max_len, 0, T1, nullptr,
                                      s, L_cond, map));
            for (int l = 0; l < h->cfg.n_layers[1]; ++l) SFB_TRY(block_prefill(h, 1, l, px1, r0, rows, T1, s, map));
        }
    }
    if (n_dup > 0) {
        const int64_t per = (int64_t)h->cfg.max_rows * h->cfg.n_head * h->cfg.max_len * 64;
        SFB_TRY(launch_prefix_copy(h->kv, x0, rm_dst, rm_src, n_dup, h->cfg.n_head, h->cfg.max_len, L_cond, d,
                                   h->cfg.n_layers[0] + h->cfg.n_layers[1], per, s));
    }
    SFB_TRY(head(h, 0, x0, WS_<float>(h, h->buf.logits[0]), B, s));
    h->begun = true;
    return SFB200_OK;
}

extern "C" int sfb200_ar_begin(sfb200_ar *h, int B, int L_cond, const sfb200_ar_sampling *sp, void *stream) {
    return sfb200_ar_begin_shared(h, B, L_cond, sp, nullptr, stream);
}

// One AR step: sample pos -> blocks[1] + head[1] for position L-1 -> sample val -> L += 1 -> blocks[0] + head[0] for the
// new position L-1 (so that logits0 is ready for the next step).
static int enqueue_step(sfb200_ar *h, const float *noise, cudaStream_t s) {
    const int d = h->cfg.n_embd, B = h->B;
    int32_t *st = WS_<int32_t>(h, h->buf.st);
    float *x0 = WS_<float>(h, h->buf.x0), *x1 = WS_<float>(h, h->buf.x1);
    const int64_t draw = (int64_t)B * h->Vmax;
    SampleLaunch p;
    p.tokens = h->tokens; p.B = B; p.max_len = h->cfg.max_len; p.L = 0; p.L_cond = 0;
    p.end0 = h->cfg.end_tokens[0]; p.end1 = h->cfg.end_tokens[1]; p.sp = h->sp; p.st = st;
    p.noise_step_stride = 4 * draw; p.noise_row_stride = h->Vmax;
    p.logp = WS_<float>(h, h->buf.logp); p.logp_row_stride = 2 * (int64_t)h->cfg.max_steps;
    // --- position
    p.logits = WS_<float>(h, h->buf.logits[0]); p.V = h->cfg.vocab[0]; p.tuple_i = 0;
    p.hist = h->hist; p.hist_row_stride = (int64_t)h->cfg.max_steps * h->cfg.vocab[0];
    p.noise_sample = noise; p.noise_best = noise + draw;
    SFB_TRY(launch_sample(p, s));
    // --- value
    SFB_TRY(launch_add_target(x0, x1, h->tokens, W_(h, SFB200_W_TOK_EMB0, 0, 0), B, d, h->cfg.max_len, 0, 1, st, s, 1));
    SFB_TRY(group_step(h, 1, x1, WS_<float>(h, h->buf.logits[1]), s));
    p.logits = WS_<float>(h, h->buf.logits[1]); p.V = h->cfg.vocab[1]; p.tuple_i = 1;
    p.hist = h->hist ? h->hist + (int64_t)h->cfg.max_rows * h->cfg.max_steps * h->cfg.vocab[0] : nullptr;
    p.hist_row_stride = (int64_t)h->cfg.max_steps * h->cfg.vocab[1];
    p.noise_sample = noise + 2 * draw; p.noise_best = noise + 3 * draw;
    SFB_TRY(launch_sample(p, s));
    SFB_TRY(launch_advance(st, h->tokens, B, h->cfg.max_len, p.end0, p.end1, s));
    if (!h->capturing) h->steps_host += 1;
    // --- next position through blocks[0]
    SFB_TRY(launch_embed(h->tokens, W_(h, SFB200_W_TOK_EMB0, 0, 0), W_(h, SFB200_W_TOK_EMB1, 0, 0),
                         W_(h, SFB200_W_EXTRA_EMB, 0, 0), W_(h, SFB200_W_POS_EMB, 0, 0), W_(h, SFB200_W_COND_POS_EMB, 0, 0),
                         x0, B, d, h->cfg.max_len, 0, 1, 0, p.end0, st, s));
    SFB_TRY(group_step(h, 0, x0, WS_<float>(h, h->buf.logits[0]), s));
    return SFB200_OK;
}

static bool same_sp(const sfb200_ar_sampling &a, const sfb200_ar_sampling &b) { return memcmp(&a, &b, sizeof(a)) == 0; }

extern "C" int sfb200_ar_steps(sfb200_ar *h, int n_steps, const float *noise, int use_graph, void *stream) {
    if (!h || !noise || n_steps < 0) return SFB200_E_ARG;
    if (!h->begun) return SFB200_E_STATE;
    cudaStream_t s = as_stream(stream);
    int32_t *st = WS_<int32_t>(h, h->buf.st);
    SFB_TRY(launch_chunk_reset(st, s));
    int done = 0;
    if (use_graph) {
        const bool valid = h->gexec && h->g_noise == noise && h->g_B == h->B && same_sp(h->g_sp, h->sp) &&
                           h->g_group == h->attn_group;
        if (!valid) {
            if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
            if (n_steps > 0) {   // first step eagerly: loads modules and sets function attributes outside capture
                SFB_TRY(enqueue_step(h, noise, s));
                done = 1;
            }
            cudaGraph_t graph = nullptr;
            if (!h->cap_stream) SFB_CUDA_TRY(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
            SFB_CUDA_TRY(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
            h->capturing = true;
            const long long before = sfb200_launch_count();
            const int r = enqueue_step(h, noise, h->cap_stream);
            h->g_nodes = sfb200_launch_count() - before;
            count_launches(-h->g_nodes);   // recorded, not executed
            h->capturing = false;
            const cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
            if (r != SFB200_OK) { if (graph) cudaGraphDestroy(graph); return r; }
            SFB_CUDA_TRY(e);
            const cudaError_t ei = cudaGraphInstantiate(&h->gexec, graph, 0);
            cudaGraphDestroy(graph);
            SFB_CUDA_TRY(ei);
            h->g_noise = noise; h->g_B = h->B; h->g_sp = h->sp; h->g_group = h->attn_group;
        }
        for (; done < n_steps; ++done) {
            SFB_CUDA_TRY(cudaGraphLaunch(h->gexec, s));
            count_launches(h->g_nodes);
            h->steps_host += 1;
        }
    } else {
        for (; done < n_steps; ++done) SFB_TRY(enqueue_step(h, noise, s));
    }
    return SFB200_OK;
}

extern "C" int sfb200_ar_profile(sfb200_ar *h, int enable) {
    if (!h) return SFB200_E_ARG;
    h->prof = enable != 0;
    h->ev_used = 0;
    h->prof_bytes = 0.0;
    h->prof_bytes_per_row = 0.0;
    return SFB200_OK;
}

extern "C" int sfb200_ar_profile_read(sfb200_ar *h, double *attn_ms, int64_t *attn_launches, double *attn_bytes,
                                      double *attn_bytes_per_row) {
    if (!h || !attn_ms || !attn_launches || !attn_bytes) return SFB200_E_ARG;
    double ms = 0.0;
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
        float t = 0.f;
        SFB_CUDA_TRY(cudaEventElapsedTime(&t, (*h->ev)[i], (*h->ev)[i + 1]));
        ms += t;
    }
    *attn_ms = ms;
    *attn_launches = (int64_t)(h->ev_used / 2);
    *attn_bytes = h->prof_bytes;
    if (attn_bytes_per_row) *attn_bytes_per_row = h->prof_bytes_per_row;
    h->ev_used = 0;
    h->prof_bytes = 0.0;
    h->prof_bytes_per_row = 0.0;
    return SFB200_OK;
}

extern "C" int64_t sfb200_ar_pretiled_floats(const sfb200_ar_config *cfg) {
    Layout L;
    if (make_layout(cfg, &L) != SFB200_OK) return -1;
    const int d = cfg->n_embd;
    return (int64_t)(cfg->n_layers[0] + cfg->n_layers[1]) * wt_block_floats(cfg) + tc_pretiled_floats(cfg->vocab[0], d) +
           tc_pretiled_floats(cfg->vocab[1], d);
}

extern "C" int sfb200_ar_set_lo_weights(sfb200_ar *h, float *lo_blob, void *stream) {
    if (!h || !lo_blob) return SFB200_E_ARG;
    SFB_TRY(launch_split_lo(h->w, lo_blob, (size_t)h->lay.total, as_stream(stream)));
    h->wlo = lo_blob;
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }   // the captured step used the other kernels
    return SFB200_OK;
}

extern "C" int sfb200_ar_set_pretiled(sfb200_ar *h, float *pretiled, void *stream) {
    if (!h || !pretiled) return SFB200_E_ARG;
    cudaStream_t s = as_stream(stream);
    const int d = h->cfg.n_embd;
    h->wt = pretiled;
    for (int g = 0; g < 2; ++g) {
        for (int l = 0; l < h->cfg.n_layers[g]; ++l) {
            SFB_TRY(launch_tc_pretile(W_(h, SFB200_W_QKV_W, g, l), h->wt + wt_offset(&h->cfg, SFB200_W_QKV_W, g, l), 3 * d, d, s));
            SFB_TRY(launch_tc_pretile(W_(h, SFB200_W_PROJ_W, g, l), h->wt + wt_offset(&h->cfg, SFB200_W_PROJ_W, g, l), d, d, s));
            SFB_TRY(launch_tc_pretile(W_(h, SFB200_W_FC1_W, g, l), h->wt + wt_offset(&h->cfg, SFB200_W_FC1_W, g, l), 4 * d, d, s));
            SFB_TRY(launch_tc_pretile(W_(h, SFB200_W_FC2_W, g, l), h->wt + wt_offset(&h->cfg, SFB200_W_FC2_W, g, l), d, 4 * d, s));
        }
        SFB_TRY(launch_tc_pretile(W_(h, SFB200_W_HEAD_W, g, 0), h->wt + wt_offset(&h->cfg, SFB200_W_HEAD_W, g, 0), h->cfg.vocab[g], d, s));
    }
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }   // the captured step used the other kernels
    return SFB200_OK;
}
