// AR sampler kernels: embed (+ AR_N extra index), LayerNorm, skinny/tiled fp32 GEMM with cluster split-K, single-query
// attention with KV-cache append, causal prefill attention, and the mask -> filter -> sample kernel.
//
// Arithmetic is fp32 throughout with fp32 accumulation: sampled tokens must equal the reference's for a given noise
// tensor, which needs logits within ~1e-5 of the fp32 PyTorch path.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "ar_kernels.cuh"

namespace cg = cooperative_groups;

namespace sfb {

// ---------------------------------------------------------------------------------------------------------------------
// device state words (int32), shared by all step kernels so that a captured CUDA graph is step-invariant
// ---------------------------------------------------------------------------------------------------------------------
//   st[ST_STEPS]  steps completed so far in this batch (the reference's loop variable j)
//   st[ST_ENDED]  first step index at which every row's newest tuple held an end token, or -1
//   st[ST_LEN]    current sequence length L (the position being sampled); newest complete tuple is L-1
//   st[ST_LCOND]  L_cond
//   st[ST_CHUNK]  step index inside the current sfb200_ar_steps call (selects the noise slab)

__global__ void ar_state_init_kernel(int32_t *st, int L_cond) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) {
        st[ST_STEPS] = 0;
        st[ST_ENDED] = -1;
        st[ST_LEN] = L_cond;
        st[ST_LCOND] = L_cond;
        st[ST_CHUNK] = 0;
    }
}
__global__ void ar_chunk_reset_kernel(int32_t *st) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) st[ST_CHUNK] = 0;
}

// After the val sub-pass: end detection (shapeformer.py:110-115) and L += 1.
__global__ void ar_advance_kernel(int32_t *st, const int64_t *tokens, int B, int max_len, int64_t end0, int64_t end1) {
    pdl_trigger();
    pdl_wait();
    __shared__ int not_ended;
    if (threadIdx.x == 0) not_ended = 0;
    __syncthreads();
    const int L = st[ST_LEN];
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int64_t *t = tokens + ((size_t)b * max_len + L) * 2;
        if (t[0] != end0 && t[1] != end1) atomicAdd(&not_ended, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (not_ended == 0 && st[ST_ENDED] < 0) st[ST_ENDED] = st[ST_STEPS];
        st[ST_STEPS] += 1;
        st[ST_LEN] = L + 1;
        st[ST_CHUNK] += 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Embedding: tok_embs[0](pos) + tok_embs[1](val) + extra_tok_embs[0](extra) + positional row, with AR_N's extra index
// (transformer/mingpt.py:256-286, representers.py:187-196,432-442) computed in place.
// One CTA per (row, position).  T positions starting at t0 (t0 read from st[ST_LEN]-1 when st != NULL).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ar_embed_kernel(const int64_t *tokens, const float *__restrict__ emb0,
                                                       const float *__restrict__ emb1, const float *__restrict__ embx,
                                                       const float *__restrict__ pos_emb,
                                                       const float *__restrict__ cond_pos_emb, float *x,
                                                       int d, int max_len, int t0_arg, int T, int L_cond_arg,
                                                       int64_t end0, const int32_t *st, const int32_t *rowmap) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;                      // compact row (index into x)
    const int br = rowmap ? rowmap[b] : b;         // row of the token buffer
    const int t0 = st ? st[ST_LEN] - 1 : t0_arg;
    const int L_cond = st ? st[ST_LCOND] : L_cond_arg;
    const int t = t0 + blockIdx.x;
    const int64_t *row = tokens + (size_t)br * max_len * 2;
    const int64_t pos = row[2 * t], val = row[2 * t + 1];
    int64_t extra;
    if (t < L_cond) {
        extra = pos;
    } else if (pos == end0) {
        extra = end0;
    } else {
        // searchsorted(c_pos, pos, right=True): first index with c_pos[idx] > pos
        int lo = 0, hi = L_cond;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (row[2 * mid] > pos) hi = mid; else lo = mid + 1;
        }
        if (lo >= L_cond) lo = L_cond - 1;  // reference would raise in gather; conditioning always ends with end token
        extra = row[2 * lo];
    }
    const float *e0 = emb0 + (size_t)pos * d, *e1 = emb1 + (size_t)val * d, *ex = embx + (size_t)extra * d;
    const float *pe = (t < L_cond) ? cond_pos_emb + (size_t)t * d : pos_emb + (size_t)(t - L_cond) * d;
    float *o = x + ((size_t)b * T + blockIdx.x) * d;
    for (int i = threadIdx.x * 4; i < d; i += blockDim.x * 4) {
        float4 a = ld4(e0 + i), c = ld4(e1 + i), e = ld4(ex + i), p = ld4(pe + i);
        float4 r;
        r.x = ((a.x + c.x) + e.x) + p.x;
        r.y = ((a.y + c.y) + e.y) + p.y;
        r.z = ((a.z + c.z) + e.z) + p.z;
        r.w = ((a.w + c.w) + e.w) + p.w;
        st4(o + i, r);
    }
}

// x_out[b, i] = x_in[b, i] + emb0[tokens[b][t(b,i) + 1][0]]   ("x = x + tok_embs[0](target)", mingpt.py:309)
// x_out rows are laid out (b, i) with i < T, x_in rows (b, i) with row stride Tin; source position t = t0 + i.
__global__ void __launch_bounds__(256) ar_add_target_kernel(const float *x_in, float *x_out,
                                                            const int64_t *tokens,
                                                            const float *__restrict__ emb0, int d, int max_len, int t0_arg,
                                                            int T, int Tin, const int32_t *st, const int32_t *rowmap) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int br = rowmap ? rowmap[b] : b;
    const int t0 = st ? st[ST_LEN] - 1 : t0_arg;
    const int t = t0 + blockIdx.x;
    const int64_t tgt = tokens[((size_t)br * max_len + t + 1) * 2];
    const float *e = emb0 + (size_t)tgt * d;
    const float *xi = x_in + ((size_t)b * Tin + blockIdx.x) * d;
    float *xo = x_out + ((size_t)b * T + blockIdx.x) * d;
    for (int i = threadIdx.x * 4; i < d; i += blockDim.x * 4) {
        float4 a = ld4(xi + i), c = ld4(e + i);
        a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
        st4(xo + i, a);
    }
}

// out[b, :] = x[b, T-1, :]
__global__ void ar_take_last_kernel(const float *x, float *out, int d, int T, const int32_t *rowmap) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x;
    const float *s = x + ((size_t)b * T + (T - 1)) * d;
    float *o = out + (size_t)(rowmap ? rowmap[b] : b) * d;
    for (int i = threadIdx.x * 4; i < d; i += blockDim.x * 4) st4(o + i, ld4(s + i));
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, biased variance), one warp per row.  NV > 0: the row (NV float4 per lane, d = 128 NV) is read once
// and kept in registers for the mean, the variance and the normalisation; NV == 0: generic three-pass form.
// ---------------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float *x, const float *__restrict__ w,
                                                        const float *__restrict__ bvec, float *y, int rows,
                                                        int d, float *y_lo) {
    pdl_trigger();
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float *xr = x + (size_t)warp * d;
    float *yr = y + (size_t)warp * d;
    if (NV > 0) {
        float4 v[NV > 0 ? NV : 1];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            v[j] = ld4(xr + lane * 4 + 128 * j);
            s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        }
        const float mean = warp_sum(s) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
            q += (a * a + b * b) + (c * c + e * e);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane * 4 + 128 * j;
            const float4 g = ld4(w + i), bb = ld4(bvec + i);
            float4 r;
            r.x = (v[j].x - mean) * rstd * g.x + bb.x;
            r.y = (v[j].y - mean) * rstd * g.y + bb.y;
            r.z = (v[j].z - mean) * rstd * g.z + bb.z;
            r.w = (v[j].w - mean) * rstd * g.w + bb.w;
            st4(yr + i, r);
            // lo part of the TF32 split (hi = hardware truncation) for a tc_big GEMM consumer
            if (y_lo) st4(y_lo + (size_t)warp * d + i, make_float4(tf32_lo(r.x), tf32_lo(r.y), tf32_lo(r.z), tf32_lo(r.w)));
        }
        return;
    }
    float s = 0.f;
    for (int i = lane * 4; i < d; i += 128) {
        float4 v = ld4(xr + i);
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
    for (int i = lane * 4; i < d; i += 128) {
        float4 v = ld4(xr + i);
        float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
        q += (a * a + b * b) + (c * c + e * e);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + 1e-5f);
    for (int i = lane * 4; i < d; i += 128) {
        float4 v = ld4(xr + i), g = ld4(w + i), bb = ld4(bvec + i), r;
        r.x = (v.x - mean) * rstd * g.x + bb.x;
        r.y = (v.y - mean) * rstd * g.y + bb.y;
        r.z = (v.z - mean) * rstd * g.z + bb.z;
        r.w = (v.w - mean) * rstd * g.w + bb.w;
        st4(yr + i, r);
        if (y_lo) st4(y_lo + (size_t)warp * d + i, make_float4(tf32_lo(r.x), tf32_lo(r.y), tf32_lo(r.z), tf32_lo(r.w)));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Linear: y = act(x W^T + bias) + residual, fp32 FFMA.
// CTA tile (16*MI rows) x 64 cols, K streamed in 32-wide slabs through a 3-stage cp.async ring; both operands stay
// K-major in shared memory (row stride 36 floats) and every thread reads float4 along K:
//   thread (tn = tid & 15, tm = tid >> 4) owns outputs  m = tm + 16 i (i < MI),  n = tn + 16 j (j < 4).
// Split-K runs as a thread-block cluster along z: partial tiles are exchanged through distributed shared memory and
// summed in rank order (deterministic), then bias / GELU / residual are applied once.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LIN_NT = 64, LIN_BK = 32, LIN_S = LIN_BK + 4, LIN_STAGES = 3;

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

template <int MI, int KSPLIT>
__global__ void __launch_bounds__(256) linear_kernel(const float *x, const float *__restrict__ W,
                                                     const float *__restrict__ bias, const float *residual, float *y,
                                                     int M, int N, int K, int act) {
    pdl_trigger();
    pdl_wait();
    constexpr int MT = 16 * MI;
    extern __shared__ __align__(16) float lin_smem[];
    float *xs = lin_smem;                             // [STAGES][MT][S]
    float *ws = lin_smem + LIN_STAGES * MT * LIN_S;   // [STAGES][NT][S]

    const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
    const int n0 = blockIdx.x * LIN_NT, m0 = blockIdx.y * MT;
    const int kslice = K / KSPLIT;
    const int kbeg = (KSPLIT > 1 ? (int)blockIdx.z : 0) * kslice;
    const int nk = kslice / LIN_BK;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kbeg + kt * LIN_BK;
        for (int i = tid; i < LIN_NT * 8; i += 256) {
            const int r = i >> 3, c = i & 7, n = n0 + r;
            const float *src = W + (size_t)(n < N ? n : N - 1) * K + k0 + c * 4;
            cp_async16(ws + (stage * LIN_NT + r) * LIN_S + c * 4, src, n < N ? 16 : 0);
        }
        for (int i = tid; i < MT * 8; i += 256) {
            const int r = i >> 3, c = i & 7, m = m0 + r;
            const float *src = x + (size_t)(m < M ? m : M - 1) * K + k0 + c * 4;
            cp_async16(xs + (stage * MT + r) * LIN_S + c * 4, src, m < M ? 16 : 0);
        }
    };

    float acc[MI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < LIN_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<LIN_STAGES - 2>();
        __syncthreads();
        if (kt + LIN_STAGES - 1 < nk) load_stage((kt + LIN_STAGES - 1) % LIN_STAGES, kt + LIN_STAGES - 1);
        cp_async_commit();
        const int stage = kt % LIN_STAGES;
        const float *wst = ws + (stage * LIN_NT + tn) * LIN_S;
        const float *xst = xs + (stage * MT + tm) * LIN_S;
#pragma unroll
        for (int k4 = 0; k4 < LIN_BK / 4; ++k4) {
            float4 wv[4], xv[MI];
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = ld4(wst + 16 * j * LIN_S + k4 * 4);
#pragma unroll
            for (int i = 0; i < MI; ++i) xv[i] = ld4(xst + 16 * i * LIN_S + k4 * 4);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = acc[i][j];
                    a = fmaf(xv[i].x, wv[j].x, a);
                    a = fmaf(xv[i].y, wv[j].y, a);
                    a = fmaf(xv[i].z, wv[j].z, a);
                    a = fmaf(xv[i].w, wv[j].w, a);
                    acc[i][j] = a;
                }
        }
    }
    cp_async_wait<0>();

    if (KSPLIT == 1) {
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int m = m0 + tm + 16 * i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tn + 16 * j;
                if (n >= N) continue;
                float v = acc[i][j];
                if (bias) v += bias[n];
                if (act == 1) v = gelu_erf(v);
                if (residual) v += residual[(size_t)m * N + n];
                y[(size_t)m * N + n] = v;
            }
        }
    } else {
        cg::cluster_group cluster = cg::this_cluster();
        __syncthreads();  // everyone is done with the operand ring; reuse it for the partial tile
        float *red = lin_smem;  // [MT][64]
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(tm + 16 * i) * LIN_NT + tn + 16 * j] = acc[i][j];
        cluster.sync();
        const unsigned rank = cluster.block_rank();
        constexpr int CHUNK = MT * LIN_NT / KSPLIT;
        const float *remote[KSPLIT];
#pragma unroll
        for (int r = 0; r < KSPLIT; ++r) remote[r] = cluster.map_shared_rank(red, r);
        for (int e = tid; e < CHUNK; e += 256) {
            const int idx = rank * CHUNK + e;
            float v = 0.f;
#pragma unroll
            for (int r = 0; r < KSPLIT; ++r) v += remote[r][idx];
            const int m = m0 + idx / LIN_NT, n = n0 + idx % LIN_NT;
            if (m < M && n < N) {
                if (bias) v += bias[n];
                if (act == 1) v = gelu_erf(v);
                if (residual) v += residual[(size_t)m * N + n];
                y[(size_t)m * N + n] = v;
            }
        }
        cluster.sync();  // keep shared memory alive until every peer has read it
    }
}

template <int MI, int KSPLIT>
static int launch_linear_t(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N,
                           int K, int act, cudaStream_t stream) {
    constexpr int MT = 16 * MI;
    const size_t smem = (size_t)LIN_STAGES * (MT + LIN_NT) * LIN_S * sizeof(float);
    static unsigned long long attr_done = 0;   // bit per device
    if (first_use_on_device(attr_done)) {
        SFB_CUDA_TRY(cudaFuncSetAttribute(linear_kernel<MI, KSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    return launch_ex("linear", linear_kernel<MI, KSPLIT>, dim3((N + LIN_NT - 1) / LIN_NT, (M + MT - 1) / MT, KSPLIT), dim3(256), smem,
                     stream, dim3(1, 1, KSPLIT), x, W, bias, residual, y, M, N, K, act);
}

template <int MI>
static int launch_linear_m(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N,
                           int K, int act, int ksplit, cudaStream_t stream) {
    switch (ksplit) {
        case 1: return launch_linear_t<MI, 1>(x, W, bias, residual, y, M, N, K, act, stream);
        case 2: return launch_linear_t<MI, 2>(x, W, bias, residual, y, M, N, K, act, stream);
        case 4: return launch_linear_t<MI, 4>(x, W, bias, residual, y, M, N, K, act, stream);
        case 8: return launch_linear_t<MI, 8>(x, W, bias, residual, y, M, N, K, act, stream);
    }
    return SFB200_E_ARG;
}

int pick_ksplit(int M, int N, int K, int MT) {
    const int tiles = ((N + LIN_NT - 1) / LIN_NT) * ((M + MT - 1) / MT);
    int ks = 1;
    // grow the split while the grid is under ~2 CTAs per SM and every slice keeps >= 4 K-slabs
    while (ks < 8 && tiles * ks < 2 * 148 && K % (ks * 2 * LIN_BK) == 0 && K / (ks * 2) >= 4 * LIN_BK) ks *= 2;
    return ks;
}

int launch_linear(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                  int act, cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0 || K % LIN_BK != 0) return SFB200_E_ARG;
    if (M <= 4) return launch_gemv(x, W, bias, residual, y, M, N, K, act, stream);
    if (M <= 16) return launch_linear_m<1>(x, W, bias, residual, y, M, N, K, act, pick_ksplit(M, N, K, 16), stream);
    if (M <= 32) return launch_linear_m<2>(x, W, bias, residual, y, M, N, K, act, pick_ksplit(M, N, K, 32), stream);
    return launch_linear_m<4>(x, W, bias, residual, y, M, N, K, act, pick_ksplit(M, N, K, 64), stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// GEMV-style linear for M <= 4 activation rows (single-shape latency runs): pure weight streaming.  One warp per output
// feature: lanes stride the K dimension with 128-bit loads (all of a row's loads are issued before use), the M activation
// rows are read through L1, and a shuffle reduction finishes each dot product.  fp32 FFMA, deterministic order.
// ---------------------------------------------------------------------------------------------------------------------
template <int MR>
__global__ void __launch_bounds__(256) gemv_kernel(const float *x, const float *__restrict__ W,
                                                   const float *__restrict__ bias, const float *residual, float *y, int N,
                                                   int K, int act) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    const float *wr = W + (size_t)n * K;
    float acc[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[m] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 1024) {
        float4 wv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + i * 128 + lane * 4;
            wv[i] = k < K ? ld4_stream(wr + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + i * 128 + lane * 4;
            if (k < K) {
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const float4 xv = ld4(x + (size_t)m * K + k);
                    float a = acc[m];
                    a = fmaf(wv[i].x, xv.x, a); a = fmaf(wv[i].y, xv.y, a);
                    a = fmaf(wv[i].z, xv.z, a); a = fmaf(wv[i].w, xv.w, a);
                    acc[m] = a;
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        float v = warp_sum(acc[m]);
        if (lane == 0) {
            if (bias) v += bias[n];
            if (act == 1) v = gelu_erf(v);
            if (residual) v += residual[(size_t)m * N + n];
            y[(size_t)m * N + n] = v;
        }
    }
}

int launch_gemv(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                int act, cudaStream_t s) {
    if (M < 1 || M > 4 || K % 4 != 0) return SFB200_E_ARG;
    const dim3 grid((N + 7) / 8);
    switch (M) {
        case 1: return launch_ex("gemv", gemv_kernel<1>, grid, dim3(256), 0, s, dim3(1, 1, 1), x, W, bias, residual, y, N, K, act);
        case 2: return launch_ex("gemv", gemv_kernel<2>, grid, dim3(256), 0, s, dim3(1, 1, 1), x, W, bias, residual, y, N, K, act);
        case 3: return launch_ex("gemv", gemv_kernel<3>, grid, dim3(256), 0, s, dim3(1, 1, 1), x, W, bias, residual, y, N, K, act);
        default: return launch_ex("gemv", gemv_kernel<4>, grid, dim3(256), 0, s, dim3(1, 1, 1), x, W, bias, residual, y, N, K, act);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention, one new query per (row, head), with KV-cache append.
// CTA = 4 warps for one (b, h, split).  A half-warp owns one key at a time: lane c = lane & 15 holds dims 4c..4c+3 of
// q / k / v, so one key row (256 B) is one coalesced 16-lane float4 load; each half-warp keeps its own online-softmax
// state (m, l, acc[4]) and the 8 states are merged through shared memory at the end.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ATT_U = 6;  // keys in flight per half-warp

struct SoftState {
    float m, l;
    float4 acc;
};

__device__ __forceinline__ float half_warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Streams keys t = t_beg + g, t_beg + g + G, ... (< t_end) for key-group g of G; key t lives at kbase + t*kstride.
// STREAM: evict-first loads for data touched once per launch; otherwise default caching (lines shared by sibling CTAs stay
// in L2 long enough to be hit).
template <bool STREAM = true>
__device__ __forceinline__ void attn_stream_keys(SoftState &st, const float4 q4, const float *kbase,
                                                 const float *vbase, size_t kstride, int t_beg, int t_end,
                                                 int g, int G, int c) {
    for (int r0 = t_beg; r0 < t_end; r0 += G * ATT_U) {
        float4 kk[ATT_U], vv[ATT_U];
        bool ok[ATT_U];
#pragma unroll
        for (int u = 0; u < ATT_U; ++u) {
            const int t = r0 + u * G + g;
            ok[u] = t < t_end;
            const size_t off = (size_t)(ok[u] ? t : t_beg) * kstride + c * 4;
            kk[u] = STREAM ? ld4_stream(kbase + off) : ld4(kbase + off);
            vv[u] = STREAM ? ld4_stream(vbase + off) : ld4(vbase + off);
        }
        float s[ATT_U];
        float mx = st.m;
#pragma unroll
        for (int u = 0; u < ATT_U; ++u) {
            float p = q4.x * kk[u].x;
            p = fmaf(q4.y, kk[u].y, p);
            p = fmaf(q4.z, kk[u].z, p);
            p = fmaf(q4.w, kk[u].w, p);
            p = half_warp_sum(p);
            s[u] = ok[u] ? p : -INFINITY;
            mx = fmaxf(mx, s[u]);
        }
        if (mx == -INFINITY) continue;
        const float corr = expf(st.m - mx);  // st.m == -inf -> 0
        st.l *= corr;
        st.acc.x *= corr; st.acc.y *= corr; st.acc.z *= corr; st.acc.w *= corr;
#pragma unroll
        for (int u = 0; u < ATT_U; ++u) {
            const float p = expf(s[u] - mx);  // -inf -> 0
            st.l += p;
            st.acc.x = fmaf(p, vv[u].x, st.acc.x);
            st.acc.y = fmaf(p, vv[u].y, st.acc.y);
            st.acc.z = fmaf(p, vv[u].z, st.acc.z);
            st.acc.w = fmaf(p, vv[u].w, st.acc.w);
        }
        st.m = mx;
    }
}

// Merge the per-half-warp states of one CTA (NG groups).  Result (unnormalised acc, M, Lsum) valid in warp 0, where
// lane handles dims 2*lane, 2*lane+1.  Group stride is 68 floats so that the float4 stores stay 16-byte aligned.
constexpr int MRG = 68;
template <int NG>
__device__ __forceinline__ void attn_merge(const SoftState &st, int grp, int c, float *sm /* NG*MRG */, float &M,
                                           float &Ls, float &o0, float &o1) {
    float *mine = sm + grp * MRG;
    if (c == 0) { mine[64] = st.m; mine[65] = st.l; }
    st4(mine + c * 4, st.acc);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        M = -INFINITY;
#pragma unroll
        for (int g = 0; g < NG; ++g) M = fmaxf(M, sm[g * MRG + 64]);
        Ls = 0.f; o0 = 0.f; o1 = 0.f;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float mg = sm[g * MRG + 64];
            const float w = (mg == -INFINITY) ? 0.f : expf(mg - M);
            Ls = fmaf(sm[g * MRG + 65], w, Ls);
            o0 = fmaf(sm[g * MRG + 2 * lane], w, o0);
            o1 = fmaf(sm[g * MRG + 2 * lane + 1], w, o1);
        }
    }
}

__global__ void __launch_bounds__(128) attn_decode_kernel(const float *qkv, float *kcache, float *vcache,
                                                          float *out, float *part, int H,
                                                          int max_len, int pos_arg, const int32_t *st_dev,
                                                          int n_split, int group, int lcond_arg, int lcond_delta) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float sm[8 * MRG];
    const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z;
    const int pos = st_dev ? st_dev[ST_LEN] - 1 : pos_arg;
    const int d = H * 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, c = lane & 15;
    const int grp = warp * 2 + half;
    const float *qrow = qkv + (size_t)b * 3 * d + h * 64 + c * 4;
    float4 q4 = ld4(qrow);
    q4.x *= 0.125f; q4.y *= 0.125f; q4.z *= 0.125f; q4.w *= 0.125f;  // 1/sqrt(64), exact
    const size_t base = ((size_t)b * H + h) * (size_t)max_len * 64;

    SoftState st;
    st.m = -INFINITY; st.l = 0.f; st.acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sp == 0 && warp == 0) {
        // the new position: append to the cache and seed group 0's state with it (no read-after-write).
        // Both halves of warp 0 run the shuffle reduction (full-mask shuffles need all 32 lanes); only half 0 keeps it.
        const float4 kn = ld4(qrow + d), vn = ld4(qrow + 2 * d);
        float p = q4.x * kn.x;
        p = fmaf(q4.y, kn.y, p); p = fmaf(q4.z, kn.z, p); p = fmaf(q4.w, kn.w, p);
        p = half_warp_sum(p);
        if (half == 0) {
            st4(kcache + base + (size_t)pos * 64 + c * 4, kn);
            st4(vcache + base + (size_t)pos * 64 + c * 4, vn);
            st.m = p; st.l = 1.f; st.acc = vn;
        }
    }
    // cached keys [0, pos) are divided between the splits in multiples of 8
    int per = (pos + n_split - 1) / n_split;
    per = (per + 7) & ~7;
    const int t_beg = min(sp * per, pos), t_end = min(t_beg + per, pos);
    // Rows of one shape (`group` consecutive rows, the reference's sample_n expansion) hold bit-identical K/V for the
    // conditioning prefix: all of them read the LEADER's copy with default caching, so the sibling CTAs (adjacent block
    // indices, scheduled together) hit each other's lines in L2 and the prefix crosses HBM once per shape.  blocks[1]'s
    // position L_cond-1 already carries each row's own first sampled pos (lcond_delta = -1).
    int shared_end = 0;
    if (group > 1) shared_end = min((st_dev ? st_dev[ST_LCOND] : lcond_arg) + lcond_delta, pos);
    const int a_end = min(t_end, max(shared_end, t_beg));
    if (a_end > t_beg) {
        const size_t lbase = ((size_t)((b / group) * group) * H + h) * (size_t)max_len * 64;
        attn_stream_keys<false>(st, q4, kcache + lbase, vcache + lbase, 64, t_beg, a_end, grp, 8, c);
    }
    attn_stream_keys<true>(st, q4, kcache + base, vcache + base, 64, a_end, t_end, grp, 8, c);

    float M, Ls, o0, o1;
    attn_merge<8>(st, grp, c, sm, M, Ls, o0, o1);
    if (threadIdx.x < 32) {
        if (n_split == 1) {
            const float inv = 1.0f / Ls;
            float2 r = make_float2(o0 * inv, o1 * inv);
            *reinterpret_cast<float2 *>(out + (size_t)b * d + h * 64 + 2 * lane) = r;
        } else {
            float *p = part + (((size_t)b * H + h) * n_split + sp) * 66;
            p[2 * lane] = o0; p[2 * lane + 1] = o1;
            if (lane == 0) { p[64] = M; p[65] = Ls; }
        }
    }
}

__global__ void __launch_bounds__(64) attn_combine_kernel(const float *part, float *out, int H,
                                                          int n_split) {
    pdl_trigger();
    pdl_wait();
    const int h = blockIdx.x, b = blockIdx.y, i = threadIdx.x;
    const float *p = part + (((size_t)b * H + h) * n_split) * 66;
    float M = -INFINITY;
    for (int s = 0; s < n_split; ++s) M = fmaxf(M, p[s * 66 + 64]);
    float Ls = 0.f, o = 0.f;
    for (int s = 0; s < n_split; ++s) {
        const float mg = p[s * 66 + 64];
        const float w = (mg == -INFINITY) ? 0.f : expf(mg - M);
        Ls = fmaf(p[s * 66 + 65], w, Ls);
        o = fmaf(p[s * 66 + i], w, o);
    }
    out[(size_t)b * H * 64 + h * 64 + i] = o / Ls;
}

// Causal prefill: qkv (B, T, 3d).  CTA = 8 warps = 8 consecutive queries of one (b, h); each warp streams keys
// [0, tq] straight from the qkv buffer (the 8 warps share them through L1) and scatters its own k, v into the cache.
__global__ void __launch_bounds__(256) attn_prefill_kernel(const float *qkv, float *kcache, float *vcache,
                                                           float *out, int H, int T, int max_len, const int32_t *rowmap) {
    pdl_trigger();
    pdl_wait();
    const int h = blockIdx.x, b = blockIdx.y;
    const int br = rowmap ? rowmap[b] : b;         // KV-cache row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tq = blockIdx.z * 8 + warp;
    if (tq >= T) return;
    const int half = lane >> 4, c = lane & 15;
    const int d = H * 64;
    const size_t rs = (size_t)3 * d;  // stride between positions
    const float *rowb = qkv + (size_t)b * T * rs + h * 64;
    float4 q4 = ld4(rowb + (size_t)tq * rs + c * 4);
    q4.x *= 0.125f; q4.y *= 0.125f; q4.z *= 0.125f; q4.w *= 0.125f;
    if (half == 0) {
        const size_t base = (((size_t)br * H + h) * (size_t)max_len + tq) * 64 + c * 4;
        st4(kcache + base, ld4(rowb + (size_t)tq * rs + d + c * 4));
        st4(vcache + base, ld4(rowb + (size_t)tq * rs + 2 * d + c * 4));
    }
    SoftState st;
    st.m = -INFINITY; st.l = 0.f; st.acc = make_float4(0.f, 0.f, 0.f, 0.f);
    attn_stream_keys(st, q4, rowb + d, rowb + 2 * d, rs, 0, tq + 1, half, 2, c);
    // merge the two halves of this warp
    const float m2 = __shfl_xor_sync(0xffffffffu, st.m, 16), l2 = __shfl_xor_sync(0xffffffffu, st.l, 16);
    float4 a2;
    a2.x = __shfl_xor_sync(0xffffffffu, st.acc.x, 16);
    a2.y = __shfl_xor_sync(0xffffffffu, st.acc.y, 16);
    a2.z = __shfl_xor_sync(0xffffffffu, st.acc.z, 16);
    a2.w = __shfl_xor_sync(0xffffffffu, st.acc.w, 16);
    const float M = fmaxf(st.m, m2);
    const float w1 = (st.m == -INFINITY) ? 0.f : expf(st.m - M), w2 = (m2 == -INFINITY) ? 0.f : expf(m2 - M);
    // combine in a fixed (half 0, half 1) order on both halves so the two copies agree bitwise
    const float wa = half ? w2 : w1, wb = half ? w1 : w2;
    const float la = half ? l2 : st.l, lb = half ? st.l : l2;
    const float4 xa = half ? a2 : st.acc, xb = half ? st.acc : a2;
    const float Ls = fmaf(lb, wb, la * wa);
    if (half == 0) {
        const float inv = 1.0f / Ls;
        float4 r;
        r.x = fmaf(xb.x, wb, xa.x * wa) * inv;
        r.y = fmaf(xb.y, wb, xa.y * wa) * inv;
        r.z = fmaf(xb.z, wb, xa.z * wa) * inv;
        r.w = fmaf(xb.w, wb, xa.w * wa) * inv;
        st4(out + ((size_t)b * T + tq) * d + h * 64 + c * 4, r);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// mask -> temperature -> top-k -> top-p -> softmax -> argmax(p / Exp(1) noise)      (one CTA of 1024 threads per row)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SMP_THREADS = 1024;

__device__ __forceinline__ uint32_t float_order(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct SampleArgs {
    const float *logits;        // (B, V)
    int64_t *tokens;            // (B, max_len, 2)
    float *hist;                // masked logits out: row b at hist + b*hist_row_stride (+ j*V when st given) or NULL
    const float *noise_sample;  // (B, V)
    const float *noise_best;    // (B, V)
    int V, npad, max_len;
    int L, L_cond, tuple_i;     // used when st == NULL
    int64_t end0, end1;
    int top_k;
    float top_p, temperature;
    int best_in_first, mask_invalid, mask_invalid_completion;
    const int32_t *st;          // device state or NULL
    int64_t hist_row_stride;    // floats between rows of the history (= max_steps * V)
    int64_t noise_step_stride;  // floats between steps in the noise slab (= 4 * B * Vmax)
    int noise_row_stride;       // floats between rows of one noise draw (= Vmax)
    float *logp;                // log-softmax(masked logits)[sampled token]: row b at logp + b*logp_row_stride + 2*step + tuple_i
    int64_t logp_row_stride;    // (= 2 * max_steps), or logp == NULL
};

__global__ void __launch_bounds__(SMP_THREADS) ar_sample_kernel(SampleArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smp_smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smp_smem);   // [npad]
    float *ev = reinterpret_cast<float *>(smp_smem + (size_t)a.npad * 8);           // [npad] exp(l_i - l_0), sorted order
    __shared__ double warp_tot[32];
    __shared__ float red_f[32];
    __shared__ int red_i[32];
    __shared__ int sh_nkeep, sh_n2;
    __shared__ float sh_sum;

    const int b = blockIdx.x, tid = threadIdx.x, V = a.V, npad = a.npad;
    const int L = a.st ? a.st[ST_LEN] : a.L;
    const int L_cond = a.st ? a.st[ST_LCOND] : a.L_cond;
    const int step_j = L - L_cond;
    const int64_t *row = a.tokens + (size_t)b * a.max_len * 2;
    const float *lg = a.logits + (size_t)b * V;
    const bool greedy = a.best_in_first && b == 0;
    const int top_k = greedy ? 1 : a.top_k;
    const float top_p = greedy ? 0.001f : a.top_p;
    size_t noise_off = (size_t)b * a.noise_row_stride;
    if (a.st) noise_off += (size_t)a.st[ST_CHUNK] * a.noise_step_stride;
    const float *q = (greedy ? a.noise_best : a.noise_sample) + noise_off;
    float *hist = a.hist ? a.hist + (size_t)b * a.hist_row_stride + (a.st ? (size_t)a.st[ST_STEPS] * V : 0) : nullptr;

    // ---- masker (representers.py:120-155)
    const int64_t last = row[2 * (L - 1)];
    int64_t nxt = 0;
    bool val_forced = false;
    if (a.tuple_i == 1) {
        val_forced = row[2 * L] == a.end0;
    } else if (a.mask_invalid_completion) {
        // cond_poses = cat(cond_pos, [end0 + 1]); searchsorted(right=True) -> first entry > last
        int lo = 0, hi = L_cond;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (row[2 * mid] > last) hi = mid; else lo = mid + 1;
        }
        nxt = lo < L_cond ? row[2 * lo] : a.end0 + 1;
    }
    auto masked = [&](int v) -> float {     // the masker's output for vocabulary entry v
        float x = lg[v];
        if (a.tuple_i == 1) {
            if (val_forced) x = (v == a.end1) ? 1.0f : -INFINITY;
        } else {
            if (a.mask_invalid && step_j > 0 && v <= last && v != a.end0) x = -INFINITY;
            if (a.mask_invalid_completion && v > nxt) x = -INFINITY;
        }
        return x;
    };
    float xmax = -INFINITY;
    for (int v = tid; v < npad; v += SMP_THREADS) {
        unsigned long long key = 0ull;
        if (v < V) {
            const float x = masked(v);
            xmax = fmaxf(xmax, x);
            if (hist) hist[v] = x;
            const float l = x / a.temperature;
            key = ((unsigned long long)float_order(l) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)v);
        }
        keys[v] = key;
    }
    // ---- log-sum-exp of the masked logits (for the sampled token's log-probability: compute_log_probs, shapeformer.py:407-418)
    float lse = 0.f;
    if (a.logp) {
        xmax = warp_max(xmax);
        if ((tid & 31) == 0) red_f[tid >> 5] = xmax;
        __syncthreads();
        xmax = warp_max(red_f[tid & 31]);
        __syncthreads();
        float se = 0.f;
        for (int v = tid; v < V; v += SMP_THREADS) se += expf(masked(v) - xmax);   // -inf -> 0
        se = warp_sum(se);
        if ((tid & 31) == 0) red_f[tid >> 5] = se;
        __syncthreads();
        se = warp_sum(red_f[tid & 31]);
        lse = xmax + logf(se);
    }
    __syncthreads();

    // ---- order the candidates, descending by (value, then lower index first), and apply top-k (ties at the k-th value are
    //      kept — common.py:265-269).  Fast path for 0 < top_k <= 1024: an 8-bit radix select finds the k-th largest value,
    //      the survivors are compacted and only they are sorted (1024-wide bitonic network instead of 8192-wide).
    __shared__ int rsel[256];
    __shared__ int sh_bin, sh_k, sh_cnt;
    const int kk = min(top_k, V);
    int np = npad;            // width of the sorted prefix array used below
    int nkeep = V;
    bool sorted = false;
    if (kk > 0 && kk <= 1024) {
        uint32_t prefix = 0, mask = 0;
        int krem = kk;
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (tid < 256) rsel[tid] = 0;
            __syncthreads();
            for (int v = tid; v < V; v += SMP_THREADS) {
                const uint32_t o = (uint32_t)(keys[v] >> 32);
                if ((o & mask) == prefix) atomicAdd(&rsel[(o >> shift) & 255u], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int c = 0, b = 255;
                for (; b > 0; --b) {
                    if (c + rsel[b] >= krem) break;
                    c += rsel[b];
                }
                sh_bin = b; sh_k = krem - c;
            }
            __syncthreads();
            prefix |= (uint32_t)sh_bin << shift;
            mask |= 0xFFu << shift;
            krem = sh_k;
        }
        const uint32_t kth = prefix;                       // order key of the k-th largest value
        unsigned long long *lst = reinterpret_cast<unsigned long long *>(ev);   // 1024 x u64 fits in the ev region
        if (tid == 0) sh_cnt = 0;
        lst[tid] = 0ull;                                   // SMP_THREADS == 1024
        __syncthreads();
        for (int v = tid; v < V; v += SMP_THREADS) {
            const unsigned long long key = keys[v];
            if ((uint32_t)(key >> 32) >= kth) {
                const int pos = atomicAdd(&sh_cnt, 1);
                if (pos < 1024) lst[pos] = key;
            }
        }
        __syncthreads();
        if (sh_cnt <= 1024) {
            nkeep = sh_cnt;
            for (int k = 2; k <= 1024; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    if (tid < 512) {
                        const int lo = tid & (j - 1);
                        const int ia = ((tid - lo) << 1) + lo, ib = ia + j;
                        const bool desc = (ia & k) == 0;
                        const unsigned long long ka = lst[ia], kb = lst[ib];
                        if ((ka < kb) == desc) { lst[ia] = kb; lst[ib] = ka; }
                    }
                    __syncthreads();
                }
            }
            keys[tid] = lst[tid];
            __syncthreads();
            np = 1024;
            sorted = true;
        }
    }
    if (!sorted) {
        for (int k = 2; k <= npad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < (npad >> 1); i += SMP_THREADS) {
                    const int lo = i & (j - 1);
                    const int ia = ((i - lo) << 1) + lo, ib = ia + j;
                    const bool desc = (ia & k) == 0;
                    const unsigned long long ka = keys[ia], kb = keys[ib];
                    if ((ka < kb) == desc) { keys[ia] = kb; keys[ib] = ka; }
                }
                __syncthreads();
            }
        }
        if (tid == 0) sh_nkeep = V;
        __syncthreads();
        if (kk > 0) {
            const uint32_t kth = (uint32_t)(keys[kk - 1] >> 32);
            for (int i = tid; i < V; i += SMP_THREADS) {
                const uint32_t oi = (uint32_t)(keys[i] >> 32);
                const uint32_t on = (i + 1 < V) ? (uint32_t)(keys[i + 1] >> 32) : 0u;
                if (oi >= kth && (i + 1 == V || on < kth)) sh_nkeep = i + 1;
            }
        }
        __syncthreads();
        nkeep = sh_nkeep;
    }

    // ---- e_i = exp(l_i - l_0) for the kept prefix (softmax numerators in sorted order)
    auto sorted_val = [&](int i) -> float {
        const uint32_t o = (uint32_t)(keys[i] >> 32);
        const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
        return __uint_as_float(u);
    };
    const float l0 = sorted_val(0);
    float part = 0.f;
    for (int i = tid; i < np; i += SMP_THREADS) {
        float e = 0.f;
        if (i < nkeep) e = expf(sorted_val(i) - l0);  // -inf -> 0
        ev[i] = e;
        part += e;
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) red_f[tid >> 5] = part;
    __syncthreads();
    if (tid < 32) {
        float t = red_f[tid];
        t = warp_sum(t);
        if (tid == 0) sh_sum = t;
    }
    __syncthreads();

    // ---- top-p: keep i == 0 or cumsum(p)[i-1] <= top_p (common.py:271-284); cumsum accumulated in fp64 like ATen's
    //      CPU cumsum (acc_type<float> = double) and rounded to fp32 before the comparison
    int n2 = nkeep;
    if (top_p > 0.0f) {
        const float sum1 = sh_sum;
        const int per = np / SMP_THREADS > 0 ? np / SMP_THREADS : 1;
        const int i0 = tid * per;
        double loc = 0.0;
        for (int u = 0; u < per; ++u) {
            const int i = i0 + u;
            if (i < np) loc += (double)(ev[i] / sum1);
        }
        // block exclusive scan of `loc`
        double inc = loc;
        const int lane = tid & 31, w = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        if (w == 0) {
            double t = warp_tot[lane];
            double ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += up;
            }
            warp_tot[lane] = ti - t;  // exclusive
        }
        __syncthreads();
        double run = warp_tot[w] + (inc - loc);
        int cnt = 0;
        for (int u = 0; u < per; ++u) {
            const int i = i0 + u;
            if (i < nkeep) {
                run += (double)(ev[i] / sum1);
                if (!((float)run > top_p)) ++cnt;
            }
        }
        // block sum of cnt
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) red_i[w] = cnt;
        __syncthreads();
        if (tid < 32) {
            int t = red_i[tid];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (tid == 0) sh_n2 = min(nkeep, 1 + t);
        }
        __syncthreads();
        n2 = sh_n2;
    }

    // ---- softmax over the survivors and the multinomial draw as argmax(p_i / q_i) (common.py:293-297)
    float s2 = 0.f;
    for (int i = tid; i < n2; i += SMP_THREADS) s2 += ev[i];
    s2 = warp_sum(s2);
    __syncthreads();
    if ((tid & 31) == 0) red_f[tid >> 5] = s2;
    __syncthreads();
    if (tid < 32) {
        float t = warp_sum(red_f[tid]);
        if (tid == 0) sh_sum = t;
    }
    __syncthreads();
    const float sum2 = sh_sum;
    float best = -1.0f;
    int best_v = 0x7fffffff;
    for (int i = tid; i < n2; i += SMP_THREADS) {
        const int v = (int)(0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull));
        const float sc = (ev[i] / sum2) / q[v];
        if (sc > best || (sc == best && v < best_v)) { best = sc; best_v = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int ov = __shfl_xor_sync(0xffffffffu, best_v, o);
        if (ob > best || (ob == best && ov < best_v)) { best = ob; best_v = ov; }
    }
    if ((tid & 31) == 0) { red_f[tid >> 5] = best; red_i[tid >> 5] = best_v; }
    __syncthreads();
    if (tid < 32) {
        best = red_f[tid]; best_v = red_i[tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            if (ob > best || (ob == best && ov < best_v)) { best = ob; best_v = ov; }
        }
        if (tid == 0) {
            a.tokens[((size_t)b * a.max_len + L) * 2 + a.tuple_i] = (int64_t)best_v;
            if (a.logp) {
                const int step = a.st ? a.st[ST_STEPS] : 0;
                a.logp[(size_t)b * a.logp_row_stride + 2 * step + a.tuple_i] = masked(best_v) - lse;
            }
        }
    }
}

// Rows that share their conditioning with an earlier row (the reference expands one shape to sample_n identical rows,
// shapeformer.py:229): copy the prefix K/V (T positions) of every block, and the blocks[0] output of the last position.
// grid (H, n_dup, n_blocks * 2); per CTA one contiguous T*64-float run.
__global__ void __launch_bounds__(256) kv_prefix_copy_kernel(float *kv, const int32_t *dst_rows, const int32_t *src_rows, int H,
                                                             int max_len, int T, int64_t per_block /* floats per K (or V) */) {
    pdl_trigger();
    pdl_wait();
    const int h = blockIdx.x, dst = dst_rows[blockIdx.y], src = src_rows[blockIdx.y];
    float *base = kv + (size_t)blockIdx.z * per_block;
    const float4 *s4 = reinterpret_cast<const float4 *>(base + ((size_t)src * H + h) * (size_t)max_len * 64);
    float4 *d4 = reinterpret_cast<float4 *>(base + ((size_t)dst * H + h) * (size_t)max_len * 64);
    for (int i = threadIdx.x; i < T * 16; i += 256) d4[i] = s4[i];
}
__global__ void __launch_bounds__(256) row_copy_kernel(float *x, const int32_t *dst_rows, const int32_t *src_rows, int d) {
    pdl_trigger();
    pdl_wait();
    const float *s = x + (size_t)src_rows[blockIdx.x] * d;
    float *o = x + (size_t)dst_rows[blockIdx.x] * d;
    for (int i = threadIdx.x * 4; i < d; i += 1024) st4(o + i, ld4(s + i));
}

// ---------------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------------
int launch_state_init(int32_t *st, int L_cond, cudaStream_t s) {
    return launch_ex("ar_state_init", ar_state_init_kernel, dim3(1), dim3(32), 0, s, dim3(1, 1, 1), st, L_cond);
}
int launch_chunk_reset(int32_t *st, cudaStream_t s) {
    return launch_ex("ar_chunk_reset", ar_chunk_reset_kernel, dim3(1), dim3(32), 0, s, dim3(1, 1, 1), st);
}
int launch_advance(int32_t *st, const int64_t *tokens, int B, int max_len, int64_t end0, int64_t end1, cudaStream_t s) {
    return launch_ex("ar_advance", ar_advance_kernel, dim3(1), dim3(128), 0, s, dim3(1, 1, 1), st, tokens, B, max_len, end0, end1);
}
int launch_embed(const int64_t *tokens, const float *emb0, const float *emb1, const float *embx, const float *pos_emb,
                 const float *cond_pos_emb, float *x, int B, int d, int max_len, int t0, int T, int L_cond, int64_t end0,
                 const int32_t *st, cudaStream_t s, const int32_t *rowmap) {
    if (T <= 0) return SFB200_OK;
    return launch_ex("ar_embed", ar_embed_kernel, dim3(T, B), dim3(256), 0, s, dim3(1, 1, 1), tokens, emb0, emb1, embx, pos_emb,
                     cond_pos_emb, x, d, max_len, t0, T, L_cond, end0, st, rowmap);
}
int launch_add_target(const float *x_in, float *x_out, const int64_t *tokens, const float *emb0, int B, int d, int max_len,
                      int t0, int T, const int32_t *st, cudaStream_t s, int Tin, const int32_t *rowmap) {
    if (T <= 0) return SFB200_OK;
    return launch_ex("ar_add_target", ar_add_target_kernel, dim3(T, B), dim3(256), 0, s, dim3(1, 1, 1), x_in, x_out, tokens, emb0, d,
                     max_len, t0, T, Tin, st, rowmap);
}
int launch_take_last(const float *x, float *out, int B, int d, int T, cudaStream_t s, const int32_t *rowmap) {
    return launch_ex("ar_take_last", ar_take_last_kernel, dim3(B), dim3(256), 0, s, dim3(1, 1, 1), x, out, d, T, rowmap);
}
int launch_layernorm(const float *x, const float *w, const float *b, float *y, int rows, int d, cudaStream_t s, float *y_lo) {
    if (rows <= 0) return SFB200_OK;
    if (d % 4 != 0) return SFB200_E_ARG;
    const dim3 grid((rows + 7) / 8);
    if (d == 1024) return launch_ex("layernorm", layernorm_kernel<8>, grid, dim3(256), 0, s, dim3(1, 1, 1), x, w, b, y, rows, d, y_lo);
    if (d == 128) return launch_ex("layernorm", layernorm_kernel<1>, grid, dim3(256), 0, s, dim3(1, 1, 1), x, w, b, y, rows, d, y_lo);
    return launch_ex("layernorm", layernorm_kernel<0>, grid, dim3(256), 0, s, dim3(1, 1, 1), x, w, b, y, rows, d, y_lo);
}
int launch_attn_decode(const float *qkv, float *kc, float *vc, float *out, float *part, int B, int H, int max_len, int pos,
                       const int32_t *st, int n_split, cudaStream_t s, int group, int lcond, int lcond_delta) {
    if (n_split < 1 || (n_split > 1 && !part) || group < 1 || B % group != 0) return SFB200_E_ARG;
    SFB_TRY(launch_ex("attn_decode", attn_decode_kernel, dim3(H, B, n_split), dim3(128), 0, s, dim3(1, 1, 1), qkv, kc, vc, out, part,
                      H, max_len, pos, st, n_split, group, lcond, lcond_delta));
    if (n_split > 1)
        SFB_TRY(launch_ex("attn_combine", attn_combine_kernel, dim3(H, B), dim3(64), 0, s, dim3(1, 1, 1), part, out, H, n_split));
    return SFB200_OK;
}
int launch_attn_prefill(const float *qkv, float *kc, float *vc, float *out, int B, int H, int T, int max_len,
                        cudaStream_t s, const int32_t *rowmap) {
    if (T <= 0) return SFB200_OK;
    // >= 32 positions: the tcgen05 kernel (attn_prefill_tc.cu: QK^T and PV on the tensor cores); SFB200_PREFILL_TC=0 keeps the FFMA one
    static int use_tc = -1;
    if (use_tc < 0) { const char *e = getenv("SFB200_PREFILL_TC"); use_tc = (e && e[0] == '0') ? 0 : 1; }
    if (use_tc && T >= 32) return launch_attn_prefill_tc(qkv, kc, vc, out, B, H, T, max_len, s, rowmap);
    return launch_ex("attn_prefill", attn_prefill_kernel, dim3(H, B, (T + 7) / 8), dim3(256), 0, s, dim3(1, 1, 1), qkv, kc, vc, out, H,
                     T, max_len, rowmap);
}

int launch_prefix_copy(float *kv, float *x0, const int32_t *dst_rows, const int32_t *src_rows, int n_dup, int H, int max_len, int T,
                       int d, int n_blocks, int64_t per_block, cudaStream_t s) {
    if (n_dup <= 0) return SFB200_OK;
    SFB_TRY(launch_ex("kv_prefix_copy", kv_prefix_copy_kernel, dim3(H, n_dup, n_blocks * 2), dim3(256), 0, s, dim3(1, 1, 1), kv,
                      dst_rows, src_rows, H, max_len, T, per_block));
    return launch_ex("row_copy", row_copy_kernel, dim3(n_dup), dim3(256), 0, s, dim3(1, 1, 1), x0, dst_rows, src_rows, d);
}

int launch_sample(const SampleLaunch &p, cudaStream_t s) {
    if (p.V < 1 || p.V > 8192) return SFB200_E_ARG;
    int npad = 1024;  // >= SMP_THREADS so that every thread owns >= 1 scan slot
    while (npad < p.V) npad <<= 1;
    SampleArgs a;
    a.logits = p.logits; a.tokens = p.tokens; a.hist = p.hist; a.noise_sample = p.noise_sample; a.noise_best = p.noise_best;
    a.V = p.V; a.npad = npad; a.max_len = p.max_len; a.L = p.L; a.L_cond = p.L_cond; a.tuple_i = p.tuple_i;
    a.end0 = p.end0; a.end1 = p.end1; a.top_k = p.sp.top_k; a.top_p = p.sp.top_p; a.temperature = p.sp.temperature;
    a.best_in_first = p.sp.best_in_first; a.mask_invalid = p.sp.mask_invalid;
    a.mask_invalid_completion = p.sp.mask_invalid_completion;
    a.st = p.st; a.hist_row_stride = p.hist_row_stride; a.noise_step_stride = p.noise_step_stride;
    a.noise_row_stride = p.noise_row_stride > 0 ? p.noise_row_stride : p.V;
    a.logp = p.logp; a.logp_row_stride = p.logp_row_stride;
    // keys (8 B each) + the ev / compaction-list region (>= 1024 x 8 B)
    const size_t smem = (size_t)npad * 8 + ((size_t)npad * 4 > 8192 ? (size_t)npad * 4 : 8192);
    static unsigned long long attr_done = 0;   // bit per device
    if (first_use_on_device(attr_done)) {
        SFB_CUDA_TRY(cudaFuncSetAttribute(ar_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
    }
    return launch_ex("ar_sample", ar_sample_kernel, dim3(p.B), dim3(SMP_THREADS), smem, s, dim3(1, 1, 1), a);
}

}  // namespace sfb
