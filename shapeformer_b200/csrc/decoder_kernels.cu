// VQDIF decoder side: codebook gather, feature-grid layout change, the fused per-point implicit decoder (trilinear feature
// sampling + 16-layer ResNet-FC MLP) in its fp32 FFMA form, and the token -> dense code grid scatter.
#include "decoder_kernels.cuh"

namespace sfb {

// ---------------------------------------------------------------------------------------------------------------------
// Quantizer.get_code (vqdif/quantizer.py:19-30): out[b][c][cell] = codebook[ind[b][cell]][c]
// A CTA transposes a 32-cell x C tile through shared memory so both the gather and the store are coalesced.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) code_gather_kernel(const int64_t *__restrict__ ind, const float *__restrict__ cb,
                                                          float *__restrict__ out, int cells, int C, int n_codes) {
    extern __shared__ float tile[];  // [32][C + 1]
    const int b = blockIdx.y, cell0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < 32; r += 8) {
        const int cell = cell0 + r;
        if (cell < cells) {
            int64_t code = ind[(size_t)b * cells + cell];
            code = code < 0 ? 0 : (code >= n_codes ? n_codes - 1 : code);
            for (int c = lane; c < C; c += 32) tile[r * (C + 1) + c] = cb[(size_t)code * C + c];
        }
    }
    __syncthreads();
    for (int c = warp; c < C; c += 8) {
        const int cell = cell0 + lane;
        if (cell < cells) out[((size_t)b * C + c) * cells + cell] = tile[lane * (C + 1) + c];
    }
}

// (B, C, S) -> (B, S, C), 32 x 32 tiles
__global__ void __launch_bounds__(256) to_channels_last_kernel(const float *__restrict__ src, float *__restrict__ dst, int C,
                                                               int64_t S) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t s0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const int64_t s = s0 + tx;
        tile[r][tx] = (c < C && s < S) ? src[((size_t)b * C + c) * S + s] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t s = s0 + r;
        const int c = c0 + tx;
        if (c < C && s < S) dst[((size_t)b * S + s) * C + c] = tile[tx][r];
    }
}

// filter_end_tokens + batch_sparse2dense (shapeformer/common.py:50-55,171-189).  One CTA per row; the scatter is done by a
// single thread in token order so that later duplicates win deterministically.
__global__ void __launch_bounds__(256) tokens_to_dense_kernel(const int64_t *__restrict__ tokens,
                                                              const int64_t *__restrict__ empty, int64_t *__restrict__ dense,
                                                              int T, int cells, int64_t end_pos, int64_t end_val) {
    const int b = blockIdx.x;
    int64_t *dr = dense + (size_t)b * cells;
    const int64_t fill = empty[b];
    for (int i = threadIdx.x; i < cells; i += blockDim.x) dr[i] = fill;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int64_t *tr = tokens + (size_t)b * T * 2;
        for (int t = 0; t < T; ++t) {
            const int64_t p = tr[2 * t], v = tr[2 * t + 1];
            if (p != end_pos && v != end_val && p >= 0 && p < cells) dr[p] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused implicit decoder, fp32 FFMA form: one thread per query point, MLP weights in the constant bank (every FFMA takes
// its weight as a uniform constant operand), activations in registers.
// ---------------------------------------------------------------------------------------------------------------------
__constant__ float c_mlp[SFB200_DEC_MLP_FLOATS];

constexpr int OFF_WP = 0, OFF_BP = 96, OFF_BLK = 128, BLK_FLOATS = 3 * (32 * 32 + 32);
constexpr int OFF_WO = OFF_BLK + 5 * BLK_FLOATS, OFF_BO = OFF_WO + 32;

// normalize_3d_coordinate (vqdif/common.py:260-276) followed by grid_sample's align_corners=True un-normalisation with
// border clamping (vqdif/dec.py:62-68): returns the continuous voxel coordinate in [0, R-1].
__device__ __forceinline__ float voxel_coord(float p, int R) {
    float pn = p / 1.101f + 0.5f;
    pn = (pn >= 1.0f) ? 0.999f : pn;
    pn = (pn < 0.0f) ? 0.0f : pn;
    const float vg = 2.0f * pn - 1.0f;
    float f = ((vg + 1.0f) / 2.0f) * (float)(R - 1);
    return fminf(fmaxf(f, 0.0f), (float)(R - 1));
}

template <int OFFW, int OFFB, bool RELU_IN, bool ACCUM>
__device__ __forceinline__ void fc32(const float (&in)[32], float (&out)[32]) {
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float a = c_mlp[OFFB + o];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float v = RELU_IN ? fmaxf(in[i], 0.f) : in[i];
            a = fmaf(c_mlp[OFFW + o * 32 + i], v, a);
        }
        out[o] = ACCUM ? out[o] + a : a;
    }
}

template <int BLK>
__device__ __forceinline__ void mlp_block(const float (&c)[32], float (&net)[32]) {
    constexpr int base = OFF_BLK + BLK * BLK_FLOATS;
    float h[32];
    fc32<base, base + 1024, false, true>(c, net);                   // net += fc_c(c)
    fc32<base + 1056, base + 1056 + 1024, true, false>(net, h);     // h = fc_0(relu(net))
    fc32<base + 2112, base + 2112 + 1024, true, true>(h, net);      // net += fc_1(relu(h))
}

__global__ void __launch_bounds__(128) decoder_points_ffma_kernel(const float *__restrict__ grid,
                                                                  const float *__restrict__ xtg, int64_t xtg_bstride,
                                                                  float *__restrict__ logits, int R, int64_t N,
                                                                  int sigmoid) {
    const int b = blockIdx.y;
    const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const float *pt = xtg + (size_t)b * xtg_bstride + n * 3;
    const float px = pt[0] * 0.5f, py = pt[1] * 0.5f, pz = pt[2] * 0.5f;   // VQDIF.decode: Xtg / 2 (vqdif/vqdif.py:71)
    const float fx = voxel_coord(px, R), fy = voxel_coord(py, R), fz = voxel_coord(pz, R);
    const float x0f = floorf(fx), y0f = floorf(fy), z0f = floorf(fz);
    const int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
    const int x1 = min(x0 + 1, R - 1), y1 = min(y0 + 1, R - 1), z1 = min(z0 + 1, R - 1);
    const float tx = fx - x0f, ty = fy - y0f, tz = fz - z0f;
    const float wx[2] = {1.0f - tx, tx}, wy[2] = {1.0f - ty, ty}, wz[2] = {1.0f - tz, tz};
    const int xs[2] = {x0, x1}, ys[2] = {y0, y1}, zs[2] = {z0, z1};
    const float *gb = grid + (size_t)b * R * R * R * 32;

    float c[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = 0.f;
    // component 0 of the point indexes the LAST grid dim (grid_sample's x -> W), 1 -> H, 2 -> D  
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const float w = wx[dx] * wy[dy] * wz[dz];
                const float *cell = gb + (((size_t)zs[dz] * R + ys[dy]) * R + xs[dx]) * 32;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(cell) + q);
                    c[4 * q + 0] = fmaf(w, v.x, c[4 * q + 0]);
                    c[4 * q + 1] = fmaf(w, v.y, c[4 * q + 1]);
                    c[4 * q + 2] = fmaf(w, v.z, c[4 * q + 2]);
                    c[4 * q + 3] = fmaf(w, v.w, c[4 * q + 3]);
                }
            }

    float net[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float a = c_mlp[OFF_BP + o];
        a = fmaf(c_mlp[OFF_WP + o * 3 + 0], px, a);
        a = fmaf(c_mlp[OFF_WP + o * 3 + 1], py, a);
        a = fmaf(c_mlp[OFF_WP + o * 3 + 2], pz, a);
        net[o] = a;
    }
    mlp_block<0>(c, net);
    mlp_block<1>(c, net);
    mlp_block<2>(c, net);
    mlp_block<3>(c, net);
    mlp_block<4>(c, net);
    float out = c_mlp[OFF_BO];
#pragma unroll
    for (int i = 0; i < 32; ++i) out = fmaf(c_mlp[OFF_WO + i], fmaxf(net[i], 0.f), out);
    if (sigmoid) out = 1.0f / (1.0f + expf(-out));
    logits[(size_t)b * N + n] = out;
}

// ---------------------------------------------------------------------------------------------------------------------
int launch_code_gather(const int64_t *ind, const float *cb, float *out, int B, int cells, int C, int n_codes, cudaStream_t s) {
    if (B <= 0 || cells <= 0 || C <= 0) return SFB200_E_ARG;
    const size_t smem = (size_t)32 * (C + 1) * sizeof(float);
    if (smem > 48 * 1024) return SFB200_E_ARG;
    code_gather_kernel<<<dim3((cells + 31) / 32, B), 256, smem, s>>>(ind, cb, out, cells, C, n_codes);
    return check_launch("code_gather");
}
int launch_to_channels_last(const float *src, float *dst, int B, int C, int64_t S, cudaStream_t s) {
    if (B <= 0 || C <= 0 || S <= 0) return SFB200_E_ARG;
    to_channels_last_kernel<<<dim3((unsigned)((S + 31) / 32), (C + 31) / 32, B), 256, 0, s>>>(src, dst, C, S);
    return check_launch("to_channels_last");
}
int launch_tokens_to_dense(const int64_t *tokens, const int64_t *empty, int64_t *dense, int B, int T, int cells,
                           int64_t end_pos, int64_t end_val, cudaStream_t s) {
    if (B <= 0 || T < 0 || cells <= 0) return SFB200_E_ARG;
    tokens_to_dense_kernel<<<B, 256, 0, s>>>(tokens, empty, dense, T, cells, end_pos, end_val);
    return check_launch("tokens_to_dense");
}
int decoder_set_weights_ffma(const float *w, cudaStream_t s) {
    SFB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_mlp, w, sizeof(float) * SFB200_DEC_MLP_FLOATS, 0, cudaMemcpyDeviceToDevice, s));
    return SFB200_OK;
}
int launch_decoder_points_ffma(const float *grid, const float *xtg, int64_t xtg_bstride, float *logits, int B, int R,
                               int64_t N, int sigmoid, cudaStream_t s) {
    if (B <= 0 || R < 2 || N <= 0) return SFB200_E_ARG;
    decoder_points_ffma_kernel<<<dim3((unsigned)((N + 127) / 128), B), 128, 0, s>>>(grid, xtg, xtg_bstride, logits, R, N, sigmoid);
    return check_launch("decoder_points_ffma");
}

}  // namespace sfb
