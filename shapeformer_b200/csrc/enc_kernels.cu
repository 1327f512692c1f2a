// VQDIF encoder + quantiser + token packing on the GPU: partial point cloud -> (pos, val) conditioning tuples.
//   LocalPoolPointnet.forward / generate_grid_features   vqdif/enc.py:66-140   (hidden = c_dim = 32, 64^3 grid, 'max' pooling)
//   Downsampler ('crg' x 4: k2s2, k1, k2s2, k1)          vqdif/updown.py:98-113
//   Quantizer.forward (eval: nearest code)               vqdif/quantizer.py:31-53
//   VQDIF.quantize_cloud, batch_dense2sparse             vqdif/vqdif.py:36-58, shapeformer/common.py:84-122,152-169
// Replaces the third-party torch_scatter calls of the reference: scatter-max = integer atomicMax on an order-preserving key,
// scatter-mean = 64-bit FIXED-POINT atomic sums (order independent, hence deterministic) + counts.  Layouts are channel-last
// (cells x channels).  These kernels are latency / atomic bound, not roofline candidates: ~2 GFLOP per cloud.
#include <limits.h>

#include "enc_kernels.cuh"

namespace sfb {

constexpr int ENC_H = 32;            // hidden = c_dim
constexpr int ENC_R = 64;            // scatter grid resolution
constexpr double ENC_FIX = 1073741824.0;   // 2^30 fixed-point scale of the scatter-mean sums

__device__ __forceinline__ int f2key(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7FFFFFFF; }
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

__device__ __forceinline__ float enc_norm(float p) {     // normalize_3d_coordinate, vqdif/common.py:260-276
    float pn = p / 1.101f + 0.5f;
    pn = (pn >= 1.0f) ? 0.999f : pn;
    return (pn < 0.0f) ? 0.0f : pn;
}

__global__ void enc_fill_kernel(int *p, size_t n, int v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- per-point ResNet-FC stage.  STAGE 0: fc_pos + blocks[0]; 1..3: pooled gather + blocks[i]; 4: ... + fc_c + scatter-mean sums.
struct EncStageArgs {
    const float *cloud;       // (B, T, 3) in [-1, 1]
    float *net;               // (B, T, 32) in/out
    int *cell;                // (B, T) index in the 64^3 grid (written by stage 0)
    const int *pool_in;       // (B, 64^3, 32) max keys of the previous stage's output
    int *pool_out;            // same, for this stage's output (NULL for the last stage)
    unsigned long long *sum;  // (B, 64^3, 32) fixed-point sums of c (last stage)
    int *count;               // (B, 64^3)
    unsigned char *mask;      // (B, 16^3) occupancy of the code grid (stage 0)
    const float *fc_pos_w, *fc_pos_b, *fc0_w, *fc0_b, *fc1_w, *fc1_b, *sc_w, *fcc_w, *fcc_b;
    int T, stage, last;
};

__global__ void __launch_bounds__(128) enc_stage_kernel(EncStageArgs a) {
    __shared__ float s_fc0[ENC_H * 65], s_sc[ENC_H * 65], s_fc1[ENC_H * 33], s_fcc[ENC_H * 33];
    __shared__ float s_pos[64 * 3 + 64], s_b0[ENC_H], s_b1[ENC_H], s_bc[ENC_H];
    const int tid = threadIdx.x, b = blockIdx.y;
    for (int i = tid; i < ENC_H * 64; i += 128) {
        s_fc0[(i >> 6) * 65 + (i & 63)] = a.fc0_w[i];
        s_sc[(i >> 6) * 65 + (i & 63)] = a.sc_w[i];
    }
    for (int i = tid; i < ENC_H * ENC_H; i += 128) {
        s_fc1[(i >> 5) * 33 + (i & 31)] = a.fc1_w[i];
        if (a.last) s_fcc[(i >> 5) * 33 + (i & 31)] = a.fcc_w[i];
    }
    if (tid < ENC_H) { s_b0[tid] = a.fc0_b[tid]; s_b1[tid] = a.fc1_b[tid]; if (a.last) s_bc[tid] = a.fcc_b[tid]; }
    if (a.stage == 0)
        for (int i = tid; i < 64 * 3 + 64; i += 128) s_pos[i] = i < 192 ? a.fc_pos_w[i] : a.fc_pos_b[i - 192];
    __syncthreads();
    const int t = blockIdx.x * 128 + tid;
    if (t >= a.T) return;
    const size_t pt = (size_t)b * a.T + t;
    float x[64];
    int cell;
    if (a.stage == 0) {
        const float px = a.cloud[pt * 3] * 0.5f, py = a.cloud[pt * 3 + 1] * 0.5f, pz = a.cloud[pt * 3 + 2] * 0.5f;   // VQDIF.encode: Xbd / 2
        const float nx = enc_norm(px), ny = enc_norm(py), nz = enc_norm(pz);
        const int ix = (int)(nx * ENC_R), iy = (int)(ny * ENC_R), iz = (int)(nz * ENC_R);
        cell = ix + ENC_R * (iy + ENC_R * iz);                       // coordinate2index, c2i_order 'original'
        a.cell[pt] = cell;
        const int mx = (int)(nx * 16.f), my = (int)(ny * 16.f), mz = (int)(nz * 16.f);   // enc.py:84-91: mask[b, z, y, x]
        a.mask[(size_t)b * 4096 + (mz * 16 + my) * 16 + mx] = 1;
#pragma unroll
        for (int o = 0; o < 64; ++o) {
            float v = s_pos[192 + o];
            v = fmaf(s_pos[o * 3], px, v); v = fmaf(s_pos[o * 3 + 1], py, v); v = fmaf(s_pos[o * 3 + 2], pz, v);
            x[o] = v;
        }
    } else {
        cell = a.cell[pt];
        const float4 *nv = reinterpret_cast<const float4 *>(a.net + pt * ENC_H);
        const int4 *pv = reinterpret_cast<const int4 *>(a.pool_in + ((size_t)b * ENC_R * ENC_R * ENC_R + cell) * ENC_H);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = nv[q];
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
            const int4 k = pv[q];
            x[32 + 4 * q] = key2f(k.x); x[32 + 4 * q + 1] = key2f(k.y); x[32 + 4 * q + 2] = key2f(k.z); x[32 + 4 * q + 3] = key2f(k.w);
        }
    }
    // ResnetBlockFC (layers.py:39-48): out = shortcut(x) + fc_1(relu(fc_0(relu(x))))
    float h[ENC_H], out[ENC_H];
#pragma unroll
    for (int o = 0; o < ENC_H; ++o) {
        float acc = s_b0[o], sc = 0.f;
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            acc = fmaf(s_fc0[o * 65 + k], fmaxf(x[k], 0.f), acc);
            sc = fmaf(s_sc[o * 65 + k], x[k], sc);
        }
        h[o] = fmaxf(acc, 0.f);
        out[o] = sc;
    }
#pragma unroll
    for (int o = 0; o < ENC_H; ++o) {
        float acc = s_b1[o];
#pragma unroll
        for (int k = 0; k < ENC_H; ++k) acc = fmaf(s_fc1[o * 33 + k], h[k], acc);
        out[o] += acc;
    }
    if (!a.last) {
        float4 *dst = reinterpret_cast<float4 *>(a.net + pt * ENC_H);
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
        int *po = a.pool_out + ((size_t)b * ENC_R * ENC_R * ENC_R + cell) * ENC_H;
#pragma unroll
        for (int o = 0; o < ENC_H; ++o) atomicMax(po + o, f2key(out[o]));
    } else {
        unsigned long long *so = a.sum + ((size_t)b * ENC_R * ENC_R * ENC_R + cell) * ENC_H;
#pragma unroll
        for (int o = 0; o < ENC_H; ++o) {
            float acc = s_bc[o];
#pragma unroll
            for (int k = 0; k < ENC_H; ++k) acc = fmaf(s_fcc[o * 33 + k], out[k], acc);
            atomicAdd(so + o, (unsigned long long)(long long)llrint((double)acc * ENC_FIX));
        }
        atomicAdd(a.count + (size_t)b * ENC_R * ENC_R * ENC_R + cell, 1);
    }
}

// ---- Downsampler convolution (kernel = stride = KS in {1, 2}, no padding, no bias) + ReLU, channel-last.
//   in   (B, Ri^3, Cin)  — or, for the first layer, the fixed-point sums / counts of the scatter-mean
//   gn   (B, Cin, 2) scale / shift of the previous layer's GroupNorm applied on load (NULL: identity)
//   wT   (KS^3 * Cin, Cout) = conv.weight.permute(2, 3, 4, 1, 0)   (tap-major, then input channel)
//   out  (B, Ro^3, Cout)
// Block = Cout threads (thread = output channel), 32 output cells per block; the 32 input patches are staged in shared memory.
template <int KS>
__global__ void enc_conv_kernel(const float *in, const unsigned long long *sum, const int *count, const float *gn, const float *wT,
                                float *out, int Ri, int Cin, int Cout) {
    extern __shared__ float s_in[];     // [32 cells][KS^3 * Cin]
    const int Ro = Ri / KS, K = KS * KS * KS * Cin;
    const int b = blockIdx.y, cell0 = blockIdx.x * 32, tid = threadIdx.x;
    const size_t in_cells = (size_t)Ri * Ri * Ri;
    for (int e = tid; e < 32 * K; e += blockDim.x) {
        const int c = e / K, k = e % K, tap = k / Cin, ci = k % Cin;
        const int oc = cell0 + c;
        const int oz = oc / (Ro * Ro), oy = (oc / Ro) % Ro, ox = oc % Ro;
        const int iz = oz * KS + tap / (KS * KS), iy = oy * KS + (tap / KS) % KS, ix = ox * KS + tap % KS;
        const size_t ic = ((size_t)iz * Ri + iy) * Ri + ix;
        float v;
        if (sum) {
            const int n = count[(size_t)b * in_cells + ic];
            v = n > 0 ? (float)((double)(long long)sum[((size_t)b * in_cells + ic) * Cin + ci] / ENC_FIX) / (float)n : 0.f;
        } else {
            v = in[((size_t)b * in_cells + ic) * Cin + ci];
            if (gn) v = fmaf(v, gn[((size_t)b * Cin + ci) * 2], gn[((size_t)b * Cin + ci) * 2 + 1]);
        }
        s_in[e] = v;
    }
    __syncthreads();
    const int co = tid;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int k = 0; k < K; k += 4) {
        const float w0 = wT[(size_t)k * Cout + co], w1 = wT[(size_t)(k + 1) * Cout + co];
        const float w2 = wT[(size_t)(k + 2) * Cout + co], w3 = wT[(size_t)(k + 3) * Cout + co];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float4 v = *reinterpret_cast<const float4 *>(s_in + c * K + k);
            float t = acc[c];
            t = fmaf(w0, v.x, t); t = fmaf(w1, v.y, t); t = fmaf(w2, v.z, t); t = fmaf(w3, v.w, t);
            acc[c] = t;
        }
    }
    const size_t out_cells = (size_t)Ro * Ro * Ro;
#pragma unroll
    for (int c = 0; c < 32; ++c) out[((size_t)b * out_cells + cell0 + c) * Cout + co] = fmaxf(acc[c], 0.f);
}

// ---- GroupNorm(8) statistics of y (B, cells, C) -> per (b, channel) scale = rstd * gamma, shift = beta - mean * rstd * gamma.
// One block per (group, b); two passes (mean, then centred second moment) accumulated in fp64, fixed reduction order.
__global__ void __launch_bounds__(256) enc_gn_stats_kernel(const float *y, const float *gamma, const float *beta, float *gn, int cells,
                                                           int C) {
    __shared__ double red[256];
    __shared__ double s_mean, s_rstd;
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, cg = C / 8;
    const float *yb = y + (size_t)b * cells * C + g * cg;
    const size_t n = (size_t)cells * cg;
    for (int pass = 0; pass < 2; ++pass) {
        double s = 0.0;
        const double mean = pass ? s_mean : 0.0;
        for (size_t e = tid; e < n; e += 256) {
            const double v = (double)yb[(e / cg) * C + (e % cg)];
            s += pass ? (v - mean) * (v - mean) : v;
        }
        red[tid] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        if (tid == 0) {
            if (pass == 0) s_mean = red[0] / (double)n;
            else s_rstd = 1.0 / sqrt(red[0] / (double)n + 1e-5);
        }
        __syncthreads();
    }
    if (tid < cg) {
        const int c = g * cg + tid;
        const float sc = (float)s_rstd * gamma[c];
        gn[((size_t)b * C + c) * 2] = sc;
        gn[((size_t)b * C + c) * 2 + 1] = beta[c] - (float)s_mean * sc;
    }
}

// ---- Quantizer.forward (eval): nearest code of every cell.  Block = 64 cells x all codes in tiles of 64; 256 threads, each a 4 x 4
// register tile of dot products; distance = (|x|^2 - 2 x.w) + |w|^2 evaluated like the reference; ties -> lowest index.
constexpr int QT = 64, QD = 128, QS = QD + 4;
__global__ void __launch_bounds__(256) enc_quantize_kernel(const float *y, const float *gn, const float *codebook, const float *ww,
                                                           int64_t *ind, float *feat_out, int cells, int n_codes) {
    extern __shared__ float qs[];
    float *xs = qs, *ws = qs + QT * QS, *xx = ws + QT * QS;      // xs[64][132], ws[64][132], xx[64]
    __shared__ float best_d[QT][16];
    __shared__ int best_i[QT][16];
    const int tid = threadIdx.x, b = blockIdx.y, cell0 = blockIdx.x * QT;
    for (int e = tid; e < QT * QD; e += 256) {
        const int c = e >> 7, k = e & 127;
        float v = y[((size_t)b * cells + cell0 + c) * QD + k];
        v = fmaf(v, gn[((size_t)b * QD + k) * 2], gn[((size_t)b * QD + k) * 2 + 1]);      // GroupNorm of the last 'crg' layer
        xs[c * QS + k] = v;
        if (feat_out) feat_out[((size_t)b * QD + k) * cells + cell0 + c] = v;           // (B, 128, cells) = the reference's grid_feat
    }
    __syncthreads();
    if (tid < QT) {
        float s = 0.f;
        for (int k = 0; k < QD; ++k) s = fmaf(xs[tid * QS + k], xs[tid * QS + k], s);
        xx[tid] = s;
    }
    const int tc = tid >> 4, tw = tid & 15;         // 16 x 16 thread grid: cells tc*4.., codes tw*4..
    float bd[4];
    int bi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bd[i] = INFINITY; bi[i] = 0; }
    for (int w0 = 0; w0 < n_codes; w0 += QT) {
        __syncthreads();
        for (int e = tid; e < QT * QD; e += 256) {
            const int c = e >> 7, k = e & 127;
            ws[c * QS + k] = (w0 + c < n_codes) ? codebook[(size_t)(w0 + c) * QD + k] : 0.f;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k = 0; k < QD; k += 4) {
            float4 xv[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4 *>(xs + (tc * 4 + i) * QS + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4 *>(ws + (tw * 4 + j) * QS + k);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float t = acc[i][j];
                    t = fmaf(xv[i].x, wv[j].x, t); t = fmaf(xv[i].y, wv[j].y, t);
                    t = fmaf(xv[i].z, wv[j].z, t); t = fmaf(xv[i].w, wv[j].w, t);
                    acc[i][j] = t;
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int code = w0 + tw * 4 + j;
                if (code < n_codes) {
                    const float d = (xx[tc * 4 + i] - 2.0f * acc[i][j]) + ww[code];
                    if (d < bd[i]) { bd[i] = d; bi[i] = code; }     // codes visited in ascending order per thread
                }
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { best_d[tc * 4 + i][tw] = bd[i]; best_i[tc * 4 + i][tw] = bi[i]; }
    __syncthreads();
    if (tid < QT) {
        float d = best_d[tid][0];
        int ix = best_i[tid][0];
        for (int j = 1; j < 16; ++j) {
            const float dj = best_d[tid][j];
            const int ij = best_i[tid][j];
            if (dj < d || (dj == d && ij < ix)) { d = dj; ix = ij; }
        }
        ind[(size_t)b * cells + cell0 + tid] = ix;
    }
}

__global__ void enc_code_norms_kernel(const float *codebook, float *ww, int n_codes) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_codes) return;
    float s = 0.f;
    for (int k = 0; k < QD; ++k) s = fmaf(codebook[(size_t)c * QD + k], codebook[(size_t)c * QD + k], s);
    ww[c] = s;
}

// ---- mode of int64 values in [0, n_bins): histogram (integer atomics: deterministic) + smallest value with the largest count
__global__ void enc_hist_kernel(const int64_t *v, const unsigned char *mask, const int64_t *fill, int *hist, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int64_t x = (mask && !mask[i]) ? *fill : v[i];
        atomicAdd(hist + x, 1);
    }
}
__global__ void __launch_bounds__(1024) enc_mode_kernel(const int *hist, int n_bins, int64_t *mode) {
    __shared__ int sc[1024], sv[1024];
    int bc = -1, bv = 0;
    for (int i = threadIdx.x; i < n_bins; i += 1024)
        if (hist[i] > bc) { bc = hist[i]; bv = i; }
    sc[threadIdx.x] = bc; sv[threadIdx.x] = bv;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const int c2 = sc[threadIdx.x + o], v2 = sv[threadIdx.x + o];
            if (c2 > sc[threadIdx.x] || (c2 == sc[threadIdx.x] && v2 < sv[threadIdx.x])) { sc[threadIdx.x] = c2; sv[threadIdx.x] = v2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *mode = sv[0];
}

// quantize_cloud (vqdif.py:50-58): dense[b][cell] = mask ? raw : mode
__global__ void enc_apply_mask_kernel(const int64_t *raw, const unsigned char *mask, const int64_t *mode, int64_t *dense, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dense[i] = mask[i] ? raw[i] : *mode;
}

// batch_dense2sparse + unpack_sparse (shapeformer/common.py:84-122,152-169): per row, the cells != mode in raveled order as
// (pos, val) tuples; tokens (B, max_len, 2) pre-filled with end tokens; lengths[b] = number of tuples (uncropped).
__global__ void __launch_bounds__(1024) enc_to_tokens_kernel(const int64_t *dense, const int64_t *mode, int64_t *tokens, int32_t *lengths,
                                                             int cells, int max_len, int64_t end0, int64_t end1) {
    __shared__ int s_cnt[1024];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int64_t m = *mode;
    const int per = (cells + 1023) / 1024, c0 = tid * per;
    int n = 0;
    for (int i = 0; i < per; ++i)
        if (c0 + i < cells && dense[(size_t)b * cells + c0 + i] != m) ++n;
    s_cnt[tid] = n;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {        // inclusive scan (Hillis-Steele)
        const int v = tid >= o ? s_cnt[tid - o] : 0;
        __syncthreads();
        s_cnt[tid] += v;
        __syncthreads();
    }
    int pos = s_cnt[tid] - n;
    for (int j = tid; j < max_len; j += 1024) { tokens[((size_t)b * max_len + j) * 2] = end0; tokens[((size_t)b * max_len + j) * 2 + 1] = end1; }
    __syncthreads();
    for (int i = 0; i < per; ++i) {
        const int c = c0 + i;
        if (c < cells) {
            const int64_t v = dense[(size_t)b * cells + c];
            if (v != m) {
                if (pos < max_len) { tokens[((size_t)b * max_len + pos) * 2] = c; tokens[((size_t)b * max_len + pos) * 2 + 1] = v; }
                ++pos;
            }
        }
    }
    if (tid == 1023) lengths[b] = s_cnt[1023];
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
struct EncWs {
    size_t net, cell, pool[2], sum, count, y[2], gn, ww, hist, total;
};
static size_t enc_align(size_t v) { return (v + 255) / 256 * 256; }
static void enc_carve(int B, int T, int n_codes, EncWs *w) {
    const size_t G = (size_t)ENC_R * ENC_R * ENC_R;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = enc_align(o + bytes); return r; };
    w->net = take((size_t)B * T * ENC_H * 4);
    w->cell = take((size_t)B * T * 4);
    w->pool[0] = take((size_t)B * G * ENC_H * 4);
    w->pool[1] = take((size_t)B * G * ENC_H * 4);
    w->sum = take((size_t)B * G * ENC_H * 8);
    w->count = take((size_t)B * G * 4);
    w->y[0] = take((size_t)B * 32768 * 64 * 4);      // largest intermediate: 32^3 cells x 64 channels
    w->y[1] = take((size_t)B * 32768 * 64 * 4);
    w->gn = take((size_t)B * 128 * 2 * 4);
    w->ww = take((size_t)n_codes * 4);
    w->hist = take((size_t)n_codes * 4);
    w->total = o;
}

int64_t enc_workspace_bytes(int B, int T, int n_codes) {
    if (B < 1 || T < 1 || n_codes < 1) return -1;
    EncWs w;
    enc_carve(B, T, n_codes, &w);
    return (int64_t)w.total;
}

int launch_encode_cloud(const sfb200_enc_weights *W, const float *cloud, int B, int T, void *workspace, int64_t *raw_ind, unsigned char *mask,
                        float *grid_feat, cudaStream_t s) {
    if (!W || !cloud || !workspace || !raw_ind || !mask || B < 1 || T < 1 || W->n_codes < 1 || W->n_codes > 65536) return SFB200_E_ARG;
    EncWs w;
    enc_carve(B, T, W->n_codes, &w);
    char *ws = static_cast<char *>(workspace);
    const size_t G = (size_t)ENC_R * ENC_R * ENC_R;
    float *net = reinterpret_cast<float *>(ws + w.net);
    int *cell = reinterpret_cast<int *>(ws + w.cell);
    int *pool[2] = {reinterpret_cast<int *>(ws + w.pool[0]), reinterpret_cast<int *>(ws + w.pool[1])};
    unsigned long long *sum = reinterpret_cast<unsigned long long *>(ws + w.sum);
    int *count = reinterpret_cast<int *>(ws + w.count);
    float *y[2] = {reinterpret_cast<float *>(ws + w.y[0]), reinterpret_cast<float *>(ws + w.y[1])};
    float *gn = reinterpret_cast<float *>(ws + w.gn), *ww = reinterpret_cast<float *>(ws + w.ww);

    enc_fill_kernel<<<1024, 256, 0, s>>>(pool[0], (size_t)B * G * ENC_H, INT_MIN);
    SFB_TRY(check_launch("enc_fill"));
    enc_fill_kernel<<<1024, 256, 0, s>>>(pool[1], (size_t)B * G * ENC_H, INT_MIN);
    SFB_TRY(check_launch("enc_fill"));
    SFB_CUDA_TRY(cudaMemsetAsync(sum, 0, (size_t)B * G * ENC_H * 8, s));
    SFB_CUDA_TRY(cudaMemsetAsync(count, 0, (size_t)B * G * 4, s));
    SFB_CUDA_TRY(cudaMemsetAsync(mask, 0, (size_t)B * 4096, s));
    for (int st = 0; st < 5; ++st) {
        EncStageArgs a;
        a.cloud = cloud; a.net = net; a.cell = cell; a.pool_in = pool[(st + 1) & 1]; a.pool_out = st < 4 ? pool[st & 1] : nullptr;
        a.sum = sum; a.count = count; a.mask = mask;
        a.fc_pos_w = W->fc_pos_w; a.fc_pos_b = W->fc_pos_b;
        a.fc0_w = W->fc0_w[st]; a.fc0_b = W->fc0_b[st]; a.fc1_w = W->fc1_w[st]; a.fc1_b = W->fc1_b[st]; a.sc_w = W->sc_w[st];
        a.fcc_w = W->fcc_w; a.fcc_b = W->fcc_b; a.T = T; a.stage = st; a.last = st == 4;
        if (st >= 2) {      // pool[st & 1] was last written by stage st - 2 and read by stage st - 1: clear it for this stage
            enc_fill_kernel<<<1024, 256, 0, s>>>(pool[st & 1], (size_t)B * G * ENC_H, INT_MIN);
            SFB_TRY(check_launch("enc_fill"));
        }
        enc_stage_kernel<<<dim3((T + 127) / 128, B), 128, 0, s>>>(a);
        SFB_TRY(check_launch("enc_stage"));
    }
    // Downsampler: 64^3 x 32 -> 32^3 x 64 (k2s2) -> 32^3 x 64 (k1) -> 16^3 x 128 (k2s2) -> 16^3 x 128 (k1), each conv + ReLU + GroupNorm(8)
    static unsigned long long attr_done = 0;
    if (first_use_on_device(attr_done)) {
        SFB_CUDA_TRY(cudaFuncSetAttribute(enc_conv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 8 * 64 * 4));
        SFB_CUDA_TRY(cudaFuncSetAttribute(enc_quantize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * QT * QS + QT) * 4));
    }
    enc_conv_kernel<2><<<dim3(32768 / 32, B), 64, 32 * 8 * 32 * 4, s>>>(nullptr, sum, count, nullptr, W->ds_wT[0], y[0], 64, 32, 64);
    SFB_TRY(check_launch("enc_conv0"));
    enc_gn_stats_kernel<<<dim3(8, B), 256, 0, s>>>(y[0], W->ds_gn_w[0], W->ds_gn_b[0], gn, 32768, 64);
    SFB_TRY(check_launch("enc_gn0"));
    enc_conv_kernel<1><<<dim3(32768 / 32, B), 64, 32 * 64 * 4, s>>>(y[0], nullptr, nullptr, gn, W->ds_wT[1], y[1], 32, 64, 64);
    SFB_TRY(check_launch("enc_conv1"));
    enc_gn_stats_kernel<<<dim3(8, B), 256, 0, s>>>(y[1], W->ds_gn_w[1], W->ds_gn_b[1], gn, 32768, 64);
    SFB_TRY(check_launch("enc_gn1"));
    enc_conv_kernel<2><<<dim3(4096 / 32, B), 128, 32 * 8 * 64 * 4, s>>>(y[1], nullptr, nullptr, gn, W->ds_wT[2], y[0], 32, 64, 128);
    SFB_TRY(check_launch("enc_conv2"));
    enc_gn_stats_kernel<<<dim3(8, B), 256, 0, s>>>(y[0], W->ds_gn_w[2], W->ds_gn_b[2], gn, 4096, 128);
    SFB_TRY(check_launch("enc_gn2"));
    enc_conv_kernel<1><<<dim3(4096 / 32, B), 128, 32 * 128 * 4, s>>>(y[0], nullptr, nullptr, gn, W->ds_wT[3], y[1], 16, 128, 128);
    SFB_TRY(check_launch("enc_conv3"));
    enc_gn_stats_kernel<<<dim3(8, B), 256, 0, s>>>(y[1], W->ds_gn_w[3], W->ds_gn_b[3], gn, 4096, 128);
    SFB_TRY(check_launch("enc_gn3"));
    enc_code_norms_kernel<<<(W->n_codes + 255) / 256, 256, 0, s>>>(W->codebook, ww, W->n_codes);
    SFB_TRY(check_launch("enc_code_norms"));
    enc_quantize_kernel<<<dim3(4096 / QT, B), 256, (2 * QT * QS + QT) * 4, s>>>(y[1], gn, W->codebook, ww, raw_ind, grid_feat, 4096, W->n_codes);
    return check_launch("enc_quantize");
}

int launch_dense_to_tokens(const int64_t *raw_ind, const unsigned char *mask, int B, int cells, int n_codes, int max_len, int64_t end0,
                           int64_t end1, void *workspace, int64_t *dense, int64_t *tokens, int32_t *lengths, int64_t *modes, cudaStream_t s) {
    if (!raw_ind || !mask || !workspace || !dense || !tokens || !lengths || !modes || B < 1 || cells < 1 || max_len < 1 || n_codes < 1)
        return SFB200_E_ARG;
    int *hist = static_cast<int *>(workspace);      // n_codes ints
    const size_t n = (size_t)B * cells;
    // mode of the raw indices over the whole batch (pth_get_mode, vqdif.py:53) -> unoccupied cells
    SFB_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)n_codes * 4, s));
    enc_hist_kernel<<<64, 256, 0, s>>>(raw_ind, nullptr, nullptr, hist, n);
    SFB_TRY(check_launch("enc_hist"));
    enc_mode_kernel<<<1, 1024, 0, s>>>(hist, n_codes, modes);
    SFB_TRY(check_launch("enc_mode"));
    enc_apply_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(raw_ind, mask, modes, dense, n);
    SFB_TRY(check_launch("enc_apply_mask"));
    // torch.mode of the masked grid (batch_dense2sparse, common.py:156) = the empty index
    SFB_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)n_codes * 4, s));
    enc_hist_kernel<<<64, 256, 0, s>>>(dense, nullptr, nullptr, hist, n);
    SFB_TRY(check_launch("enc_hist"));
    enc_mode_kernel<<<1, 1024, 0, s>>>(hist, n_codes, modes + 1);
    SFB_TRY(check_launch("enc_mode"));
    enc_to_tokens_kernel<<<B, 1024, 0, s>>>(dense, modes + 1, tokens, lengths, cells, max_len, end0, end1);
    return check_launch("enc_to_tokens");
}

}  // namespace sfb
