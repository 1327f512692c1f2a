// Fused implicit decoder on the tcgen05 tensor cores (the default sfb200_decoder_points path).
//
// One persistent CTA per SM keeps THREE 128-point tiles in flight (12 "point" warps = 3 groups x 4 warps, thread = one query
// point = one TMEM lane) plus one MMA-issuing warp.  Per tile:
//   point warps : trilinear gather of the 32-channel feature (8 corners x 128 B from the channel-last grid) -> tcgen05.st: the
//                 feature vector c is the resident half of the A operand, living in TENSOR MEMORY for the whole tile
//   11 round trips, every GEMM [128 x 32] with K = 32 or 64 (12 / 24 tcgen05.mma.kind::tf32, weights resident in smem):
//       point warps: tcgen05.ld accumulator -> + biases -> ReLU -> TF32 hi (raw bits, truncated by the MMA) | lo -> tcgen05.st
//       MMA warp   : D = A . W^T, serving whichever group's operand is ready (non-blocking mbarrier probes)
//     the fc_c projection of the next block rides in the fc_1 GEMM of the current one (K concatenation [h | c])
//   point warps : fc_out dot product, optional sigmoid, coalesced store.
// Products are 3xTF32 (hi*hi + hi*lo + lo*hi, fp32 accumulate, chains of 12 MMAs): logits stay within ~1e-6 of the fp32
// reference (tolerance 1e-4; single-pass TF32 misses it); chains of at most 24 MMAs.  fc_p (K = 3) and fc_out (N = 1) are CUDA-core.
#include "decoder_kernels.cuh"
#include "tc_common.cuh"

namespace sfb {

using namespace tc;

// small parameters, uniform per warp -> constant bank:  fc_p (96 + 32), biases, fc_out
struct DecSmall {
    float wp[96], bp[32];
    float bc[5][32], b0[5][32], b1[5][32];
    float wo[32], bo;
};
__constant__ DecSmall c_dec;

constexpr int DTC_G = 3;                  // point tiles in flight per CTA (one warpgroup each)
constexpr int DTC_THREADS = DTC_G * 128 + 32;   // + 1 MMA warp
constexpr int DTC_TILE_W = 32 * 32 * 4;   // one 32x32 fp32 weight tile (K-major SWIZZLE_128B): 4 KB
// shared memory: 15 matrices (per block: fc_c, fc_0, fc_1) x (hi | lo)
constexpr int DTC_OFF_W = 0;                                    // [15][2][4 KB]
constexpr int DTC_OFF_BAR = DTC_OFF_W + 15 * 2 * DTC_TILE_W;
constexpr int DTC_SMEM = DTC_OFF_BAR + 128;
// tensor memory per group (160 columns): A hi [act 32 | c 32] | A lo [act 32 | c 32] | D 32
constexpr int DTC_GCOLS = 160, DTC_COL_AH = 0, DTC_COL_AL = 64, DTC_COL_D = 128;

// device copy of the packed MLP weights for the tile builder (filled by decoder_set_weights_tc)
__device__ float g_dec_w[SFB200_DEC_MLP_FLOATS];

__device__ __forceinline__ float dvoxel_coord(float p, int R) {
    float pn = p / 1.101f + 0.5f;
    pn = (pn >= 1.0f) ? 0.999f : pn;
    pn = (pn < 0.0f) ? 0.0f : pn;
    const float vg = 2.0f * pn - 1.0f;
    const float f = ((vg + 1.0f) / 2.0f) * (float)(R - 1);
    return fminf(fmaxf(f, 0.0f), (float)(R - 1));
}

// write element (row, k) of a K-major SWIZZLE_128B tile (rows of 32 floats)
__device__ __forceinline__ int sw128_index(int row, int k) { return row * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3)); }

// activation vector (32 fp32 in registers) -> A operand in tensor memory.  hi = the raw fp32 bits: the tensor core reads only
// the upper 19 bits of a tf32 operand (truncation, no instruction); lo = rna_tf32(a - trunc_tf32(a)), the subtraction is exact.
__device__ __forceinline__ void store_activation(uint32_t taddr_hi, uint32_t taddr_lo, const float (&a)[32]) {
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        hi[i] = __float_as_uint(a[i]);
        lo[i] = __float_as_uint(a[i] - __uint_as_float(hi[i] & 0xFFFFE000u)) + 0x1000u;
    }
    tmem_st32(taddr_hi, hi);
    tmem_st32(taddr_lo, lo);
    tmem_st_wait();
}

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// One persistent CTA per SM keeps THREE 128-point tiles in flight (3 warpgroups of point threads + 1 MMA warp).  Per tile
// (thread = query point = TMEM lane), 11 tensor-core round trips, all K = 32 / 64, N = 32, 3xTF32:
//   init   D = c . Wc[0]^T                                    net = fc_p(p) + D + bc[0]
//   blk b  D = relu(net) . W0[b]^T                            h = relu(D + b0[b])
//          D = [h | c] . [W1[b] | Wc[b+1]]^T  (K = 64)        net += D + b1[b] + bc[b+1]        (last block: K = 32, no fc_c)
// i.e. the fc_c projection of the NEXT block rides in the fc_1 GEMM of the current one (K concatenation): the feature vector c
// stays resident in tensor memory for the whole tile and no 160-column side accumulator is needed, so three tiles fit in TMEM.
__global__ void __launch_bounds__(DTC_THREADS, 1)
decoder_points_tc_kernel(const float *__restrict__ grid, const float *__restrict__ xtg, int64_t xtg_bstride,
                         float *__restrict__ out, int B, int R, int64_t N, int sigmoid) {
    extern __shared__ __align__(1024) unsigned char dsm[];
    uint64_t *ready = reinterpret_cast<uint64_t *>(dsm + DTC_OFF_BAR);   // [G] A operand of group g is in TMEM
    uint64_t *done = ready + DTC_G;                                       // [G] GEMM for group g finished
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + DTC_G);
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time: weights -> TF32 hi/lo swizzled tiles in shared memory.  Matrix m = blk * 3 + {0: fc_c, 1: fc_0, 2: fc_1}
    // packed layout (decoder.py::pack_mlp_weights): fc_p (96+32), per block [fc_c W 1024, b 32, fc_0 W, b, fc_1 W, b], fc_out
    for (int e = tid; e < 15 * 1024; e += DTC_THREADS) {
        const int mtx = e >> 10, blk = mtx / 3, which = mtx % 3, o = (e >> 5) & 31, i = e & 31;
        uint32_t h, l;
        split_tf32(g_dec_w[128 + blk * 3168 + 1056 * which + o * 32 + i], h, l);
        float *t_hi = reinterpret_cast<float *>(dsm + DTC_OFF_W + (mtx * 2) * DTC_TILE_W);
        float *t_lo = reinterpret_cast<float *>(dsm + DTC_OFF_W + (mtx * 2 + 1) * DTC_TILE_W);
        const int idx = sw128_index(o, i);
        t_hi[idx] = __uint_as_float(h); t_lo[idx] = __uint_as_float(l);
    }
    if (tid == 0) {
        for (int g = 0; g < DTC_G; ++g) { mbar_init(&ready[g], 4); mbar_init(&done[g], 1); }
        mbar_fence_init();
    }
    if (warp == DTC_G * 4) tmem_alloc<512>(tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t tiles_per_shape = (N + 127) / 128;
    const int64_t n_tiles = tiles_per_shape * B;
    const int64_t stride = (int64_t)gridDim.x * DTC_G;

    if (warp < DTC_G * 4) {
        // =========================================== point warps ===========================================
        const int g = warp >> 2, lane = tid & 31;
        const uint32_t tg = tmem_base + g * DTC_GCOLS + ((uint32_t)(32 * (warp & 3)) << 16);
        uint32_t ph_done = 0;   // number of completed GEMM phases for this group
        auto wait_gemm = [&]() { mbar_wait(&done[g], ph_done & 1); ++ph_done; tc_fence_after(); };
        auto signal = [&]() { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&ready[g]); };
        for (int64_t tile = (int64_t)blockIdx.x * DTC_G + g; tile < n_tiles; tile += stride) {
            const int b = (int)(tile / tiles_per_shape);
            const int64_t n = (tile % tiles_per_shape) * 128 + (tid & 127);
            const bool valid = n < N;
            const int64_t nc = valid ? n : N - 1;
            const float *pt = xtg + (size_t)b * xtg_bstride + nc * 3;
            const float px = pt[0] * 0.5f, py = pt[1] * 0.5f, pz = pt[2] * 0.5f;   // VQDIF.decode: Xtg / 2
            float net[32];
            {
                // ---- trilinear feature (vqdif/dec.py:62-68) -> the resident c half of the A operand
                const float fx = dvoxel_coord(px, R), fy = dvoxel_coord(py, R), fz = dvoxel_coord(pz, R);
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
#pragma unroll
                for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 2; ++dx) {
                            const float w = wx[dx] * wy[dy] * wz[dz];
                            const float4 *cell = reinterpret_cast<const float4 *>(gb + (((size_t)zs[dz] * R + ys[dy]) * R + xs[dx]) * 32);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 v = __ldg(cell + q);
                                c[4 * q + 0] = fmaf(w, v.x, c[4 * q + 0]);
                                c[4 * q + 1] = fmaf(w, v.y, c[4 * q + 1]);
                                c[4 * q + 2] = fmaf(w, v.z, c[4 * q + 2]);
                                c[4 * q + 3] = fmaf(w, v.w, c[4 * q + 3]);
                            }
                        }
                store_activation(tg + DTC_COL_AH + 32, tg + DTC_COL_AL + 32, c);
            }
            signal();                                   // -> init: D = c . Wc[0]^T
            // net = fc_p(p) while the tensor core works
#pragma unroll
            for (int o = 0; o < 32; ++o) {
                float a = c_dec.bp[o];
                a = fmaf(c_dec.wp[o * 3 + 0], px, a);
                a = fmaf(c_dec.wp[o * 3 + 1], py, a);
                a = fmaf(c_dec.wp[o * 3 + 2], pz, a);
                net[o] = a;
            }
            wait_gemm();
#pragma unroll 1
            for (int blk = 0; blk < 5; ++blk) {
                uint32_t v[32];
                float a[32];
                // net += (fc_c[blk](c), and for blk > 0 also fc_1[blk-1](h)) + biases; A <- relu(net)
                tmem_ld32(tg + DTC_COL_D, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float add = __uint_as_float(v[i]) + c_dec.bc[blk][i];
                    if (blk > 0) add += c_dec.b1[blk - 1][i];
                    net[i] += add;
                    a[i] = fmaxf(net[i], 0.f);
                }
                store_activation(tg + DTC_COL_AH, tg + DTC_COL_AL, a);
                signal();                               // -> D = relu(net) . fc_0^T
                wait_gemm();
                tmem_ld32(tg + DTC_COL_D, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] = fmaxf(__uint_as_float(v[i]) + c_dec.b0[blk][i], 0.f);
                store_activation(tg + DTC_COL_AH, tg + DTC_COL_AL, a);
                signal();                               // -> D = [h | c] . [fc_1 | fc_c[blk+1]]^T
                wait_gemm();
            }
            {
                uint32_t v[32];
                tmem_ld32(tg + DTC_COL_D, v);
                tmem_ld_wait();
                float o = c_dec.bo;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float r = net[i] + (__uint_as_float(v[i]) + c_dec.b1[4][i]);
                    o = fmaf(c_dec.wo[i], fmaxf(r, 0.f), o);
                }
                if (sigmoid) o = 1.0f / (1.0f + expf(-o));
                if (valid) out[(size_t)b * N + n] = o;
            }
        }
        tc_fence_before();
    } else {
        // =========================================== MMA issuer ===========================================
        // the whole warp stays converged and one elected lane issues (tc_common.cuh); the three groups are served in whatever
        // order their operands become ready (non-blocking mbarrier probes)
        constexpr uint32_t IDESC32 = instr_desc(2, 128, 32);
        const uint32_t w0 = smem_u32(dsm + DTC_OFF_W);
        uint32_t ph[DTC_G], step[DTC_G];
        int64_t next_tile[DTC_G];
        int live = 0;
#pragma unroll
        for (int g = 0; g < DTC_G; ++g) {
            ph[g] = 0; step[g] = 0;
            next_tile[g] = (int64_t)blockIdx.x * DTC_G + g;
            live += next_tile[g] < n_tiles;
        }
        while (live > 0) {
#pragma unroll
            for (int g = 0; g < DTC_G; ++g) {
                if (next_tile[g] >= n_tiles) continue;
                if (!mbar_test(&ready[g], ph[g] & 1)) continue;
                ++ph[g];
                tc_fence_after();
                const uint32_t tgb = tmem_base + g * DTC_GCOLS;
                const uint32_t d = tgb + DTC_COL_D;
                const uint32_t st = step[g];
                // step 0: c . Wc[0]; odd steps 2b+1: act . W0[b]; even steps 2b+2: act . W1[b] (+ c . Wc[b+1] for b < 4)
                const int blk = st == 0 ? 0 : (int)(st - 1) >> 1;
                const int mtx = st == 0 ? 0 : ((st & 1) ? blk * 3 + 1 : blk * 3 + 2);
                const uint32_t a_off = st == 0 ? 32u : 0u;      // the c half for the init step
                const bool concat = st != 0 && (st & 1) == 0 && blk < 4;
                const uint64_t bh = smem_desc_k128(w0 + (mtx * 2) * DTC_TILE_W), bl = smem_desc_k128(w0 + (mtx * 2 + 1) * DTC_TILE_W);
                const uint64_t ch = smem_desc_k128(w0 + ((blk + 1) * 3 * 2) * DTC_TILE_W);
                const uint64_t cl = smem_desc_k128(w0 + ((blk + 1) * 3 * 2 + 1) * DTC_TILE_W);
                if (elect_one()) {
                    const uint32_t a_hi = tgb + DTC_COL_AH + a_off, a_lo = tgb + DTC_COL_AL + a_off;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        mma_tf32_ts(d, a_lo + 8 * k, bh + 2 * k, IDESC32, k != 0);
                        mma_tf32_ts(d, a_hi + 8 * k, bl + 2 * k, IDESC32, 1);
                        mma_tf32_ts(d, a_hi + 8 * k, bh + 2 * k, IDESC32, 1);
                    }
                    if (concat) {
                        const uint32_t c_hi = tgb + DTC_COL_AH + 32, c_lo = tgb + DTC_COL_AL + 32;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            mma_tf32_ts(d, c_lo + 8 * k, ch + 2 * k, IDESC32, 1);
                            mma_tf32_ts(d, c_hi + 8 * k, cl + 2 * k, IDESC32, 1);
                            mma_tf32_ts(d, c_hi + 8 * k, ch + 2 * k, IDESC32, 1);
                        }
                    }
                    mma_commit(&done[g]);
                }
                __syncwarp();
                if (++step[g] == 11) {
                    step[g] = 0;
                    next_tile[g] += stride;
                    if (next_tile[g] >= n_tiles) --live;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == DTC_G * 4) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
int decoder_set_weights_tc(const float *w, cudaStream_t s) {
    // packed order: fc_p W (96), b (32) | per block: fc_c W (1024) b (32), fc_0 W b, fc_1 W b | fc_out W (32) b (1)
    SFB_CUDA_TRY(cudaMemcpyToSymbolAsync(g_dec_w, w, sizeof(float) * SFB200_DEC_MLP_FLOATS, 0, cudaMemcpyDeviceToDevice, s));
    const size_t F = sizeof(float);
    SFB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_dec, w, 128 * F, offsetof(DecSmall, wp), cudaMemcpyDeviceToDevice, s));
    for (int blk = 0; blk < 5; ++blk) {
        const float *base = w + 128 + blk * 3168;
        SFB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_dec, base + 1024, 32 * F, offsetof(DecSmall, bc) + blk * 32 * F,
                                             cudaMemcpyDeviceToDevice, s));
        SFB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_dec, base + 1056 + 1024, 32 * F, offsetof(DecSmall, b0) + blk * 32 * F,
                                             cudaMemcpyDeviceToDevice, s));
        SFB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_dec, base + 2112 + 1024, 32 * F, offsetof(DecSmall, b1) + blk * 32 * F,
                                             cudaMemcpyDeviceToDevice, s));
    }
    SFB_CUDA_TRY(cudaMemcpyToSymbolAsync(c_dec, w + 128 + 5 * 3168, 33 * F, offsetof(DecSmall, wo), cudaMemcpyDeviceToDevice, s));
    return SFB200_OK;
}

int launch_decoder_points_tc(const float *grid, const float *xtg, int64_t xtg_bstride, float *logits, int B, int R, int64_t N,
                             int sigmoid, cudaStream_t s) {
    if (B <= 0 || R < 2 || N <= 0) return SFB200_E_ARG;
    static bool attr_done = false;
    static int n_sm = 148;
    if (!attr_done) {
        SFB_CUDA_TRY(cudaFuncSetAttribute(decoder_points_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DTC_SMEM));
        int dev = 0;
        SFB_CUDA_TRY(cudaGetDevice(&dev));
        SFB_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int64_t n_tiles = ((N + 127) / 128) * B;
    int64_t ctas = (n_tiles + 1) / 2;
    if (ctas > n_sm) ctas = n_sm;
    decoder_points_tc_kernel<<<(unsigned)ctas, DTC_THREADS, DTC_SMEM, s>>>(grid, xtg, xtg_bstride, logits, B, R, N, sigmoid);
    return check_launch("decoder_points_tc");
}

}  // namespace sfb
