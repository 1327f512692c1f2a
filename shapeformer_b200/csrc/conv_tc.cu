// Conv prologue of the VQDIF decoder as kernels of this library (UNet3D + Upsampler: vqdif/unet3d.py:449-474, updown.py:79-132):
//
//   conv3d_tc_kernel  3x3x3 (pad 1, no bias) or 1x1x1 (+ bias) convolution over channels-last (N, D, H, W, C) fp32 tensors as an
//                     implicit GEMM on the 5th-generation tensor cores with fp32-level accuracy (3xTF32):
//                         D[voxel, co] += sum_{tap, ci} in[voxel + tap][ci] * w[tap][co][ci]
//                     UMMA M = 128 output voxels (a 3-D box of the volume), UMMA N = NT output channels, K = taps x Cin in
//                     32-channel chunks.  Both operands are TMA-fed, no im2col in registers: the A tile of a (tap, chunk) is ONE
//                     5-D tiled TMA box of the input shifted by the tap offset — out-of-volume voxels are zero-filled by the TMA
//                     unit, which IS the zero padding — and lands in shared memory as a K-major SWIZZLE_128B tile (one 128-byte
//                     row per voxel).  hi operand = the raw fp32 tensor (the tensor core truncates it to tf32), lo operand = a
//                     second tensor written by whoever produced the input (conv_prep_kernel).  Weights are pre-packed
//                     [tap][co][ci] (+ lo).  Accumulators are promoted to fp32 registers every 4 chunks (two TMEM buffers).
//                     Epilogue: bias / ReLU, channels-last store, and per-(sample, channel) sum / sum of squares of the output in
//                     fp64 atomics — the statistics of the GroupNorm that follows.
//                     A convolution that FOLLOWS a nearest x2 upsampling runs in sub-pixel form (taps = 8): each of the 8 output
//                     phases is a 2x2x2 convolution of the LOW-resolution input with pre-summed weights — 27/8 fewer FLOPs and the
//                     upsampled tensor is never written.
//   conv_prep_kernel  everything between two convolutions, in one elementwise pass: GroupNorm(8) from those per-channel sums
//                     (no second pass over the producer), nearest-neighbour x2 upsampling and channel concatenation of two
//                     sources on load, and the hi / lo operand split on store (the lo tensor is what conv3d_tc reads).
//   pool_stats_kernel max-pool 2x2x2 (or a plain pass, window 1) + the per-channel sums of its output.
//   gather_codes_cl   Quantizer.get_code (vqdif/quantizer.py:19-30) straight into the channels-last layout + its sums.
#include <cuda.h>
#include <string.h>

#include "ar_kernels.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace sfb {

using namespace tc;

constexpr int CV_THREADS = 320;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: promotion + epilogue
constexpr int CV_G = 4;                  // chunks per promotion group (48 MMAs per TMEM accumulation chain)
constexpr int CV_A_TILE = 128 * 32 * 4;  // 16 KB

template <int NT>
struct CvCfg {
    static constexpr int B_TILE = NT * 32 * 4;
    static constexpr int STAGE = 2 * CV_A_TILE + 2 * B_TILE;
    static constexpr int NS = (200 * 1024) / STAGE > 5 ? 5 : (200 * 1024) / STAGE;
    static constexpr int OFF_BAR = NS * STAGE;
    static constexpr int SMEM = OFF_BAR + 256;
    static constexpr int TM_COLS = 2 * NT < 32 ? 32 : 2 * NT;
};

struct ConvArgs {
    TensorMapBlob in, in_lo;     // 5-D channels-last activations (C, X, Y, Z, B), box (32, bx, by, bz, bn)
    TensorMapBlob w, w_lo;       // packed weights (taps * Cout, Cin), box (32, NT)
    const float *bias;           // (Cout) or NULL
    float *out;                  // (B, Z, Y, X, Cout)
    double *stats;               // (B, Cout, 2) sum | sum of squares of the stored output, accumulated; or NULL
    int B, Z, Y, X, Cin, Cout;
    int bx, by, bz, bn;          // voxel box of an M tile: bx * by * bz * bn == 128
    int taps, relu;
    int up;                      // 1: the convolution follows a nearest-neighbour x2 upsampling that is NOT materialised: the input is
                                 // the low-resolution tensor, blockIdx.z = output phase (pz, py, px), taps = 8 per phase, weights
                                 // packed [phase][tap][co][ci] with the taps that hit the same input voxel summed (sub-pixel form);
                                 // out is (B, 2Z, 2Y, 2X, Cout) and this CTA writes the voxels (2z+pz, 2y+py, 2x+px)
};

__device__ __forceinline__ void cv_tma_2d(void *dst, const void *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
                     "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cv_tma_5d(void *dst, const void *map, int c0, int c1, int c2, int c3, int c4, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::
            "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cv_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

template <int NT>
__global__ void __launch_bounds__(CV_THREADS, 1) conv3d_tc_kernel(const __grid_constant__ ConvArgs a) {
    using C = CvCfg<NT>;
    constexpr int NS = C::NS;
    extern __shared__ __align__(1024) unsigned char cv_smem[];
    unsigned char *smem = cv_smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);   // [NS] stage landed (TMA tx bytes)
    uint64_t *empty = full + NS;                                        // [NS] MMAs that read the stage finished
    uint64_t *dfull = empty + NS;                                       // [2]  promotion group finished in D[b]
    uint64_t *dfree = dfull + 2;                                        // [2]  D[b] drained by the 8 epilogue warps
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dfree + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // tile -> box origin
    const int nx = a.X / a.bx, ny = a.Y / a.by, nz = a.Z / a.bz;
    int t = blockIdx.x;
    const int x0 = (t % nx) * a.bx; t /= nx;
    const int y0 = (t % ny) * a.by; t /= ny;
    const int z0 = (t % nz) * a.bz; t /= nz;
    const int b0 = t * a.bn;
    const int n0 = blockIdx.y * NT;
    const int cpt = a.Cin >> 5;                 // chunks per tap
    const int nch = a.taps * cpt;
    const int phase = a.up ? blockIdx.z : 0, pz = (phase >> 2) & 1, py = (phase >> 1) & 1, px = phase & 1;

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dfree[i], 8); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<C::TM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            asm volatile("fence.proxy.async;\n" ::: "memory");
            for (int i = 0; i < nch; ++i) {
                const int s = i % NS;
                if (i >= NS) mbar_wait(&empty[s], ((i / NS) - 1) & 1);
                const int tap = i / cpt, ch = i - tap * cpt;
                int dz = 0, dy = 0, dx = 0;
                if (a.up) {                      // phase p reads the input voxels q + p - 1 and q + p of every axis
                    dz = pz - 1 + ((tap >> 2) & 1); dy = py - 1 + ((tap >> 1) & 1); dx = px - 1 + (tap & 1);
                } else if (a.taps == 27) {
                    dz = tap / 9 - 1; dy = (tap / 3) % 3 - 1; dx = tap % 3 - 1;
                }
                const int wrow = (phase * a.taps + tap) * a.Cout + n0;
                unsigned char *st = smem + s * C::STAGE;
                cv_expect_tx(&full[s], C::STAGE);
                cv_tma_5d(st, &a.in, ch * 32, x0 + dx, y0 + dy, z0 + dz, b0, &full[s]);
                cv_tma_5d(st + CV_A_TILE, &a.in_lo, ch * 32, x0 + dx, y0 + dy, z0 + dz, b0, &full[s]);
                cv_tma_2d(st + 2 * CV_A_TILE, &a.w, ch * 32, wrow, &full[s]);
                cv_tma_2d(st + 2 * CV_A_TILE + C::B_TILE, &a.w_lo, ch * 32, wrow, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp converged, one elected lane issues) ================================
        constexpr uint32_t IDESC = instr_desc(2, 128, NT);
        const uint32_t ring0 = smem_u32(smem);
        for (int i = 0; i < nch; ++i) {
            const int s = i % NS, g = i / CV_G, b = g & 1;
            const bool first = (i % CV_G) == 0, last = (i % CV_G) == CV_G - 1 || i == nch - 1;
            if (first && g >= 2) mbar_wait(&dfree[b], ((g >> 1) - 1) & 1);
            mbar_wait(&full[s], (i / NS) & 1);
            tc_fence_after();
            const uint32_t d = tmem_base + b * NT;
            const uint32_t st = ring0 + s * C::STAGE;
            const uint64_t ah = smem_desc_k128(st), al = smem_desc_k128(st + CV_A_TILE);
            const uint64_t bh = smem_desc_k128(st + 2 * CV_A_TILE), bl = smem_desc_k128(st + 2 * CV_A_TILE + C::B_TILE);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mma_tf32_ss(d, al + 2 * k, bh + 2 * k, IDESC, !(first && k == 0));
                    mma_tf32_ss(d, ah + 2 * k, bl + 2 * k, IDESC, 1);
                    mma_tf32_ss(d, ah + 2 * k, bh + 2 * k, IDESC, 1);
                }
                mma_commit(&empty[s]);
                if (last) mma_commit(&dfull[b]);
            }
            __syncwarp();
        }
    }

    // ================================ promotion + epilogue warps ================================
    constexpr int HB = NT / 2;
    float acc[HB];                 // acc[j] = D[voxel row][half * HB + j]
    const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
    const int row = quad * 32 + lane;
    if (warp >= 2) {
#pragma unroll
        for (int j = 0; j < HB; ++j) acc[j] = 0.f;
        const uint32_t lane_off = (uint32_t)(32 * quad) << 16;
        const int ngroups = (nch + CV_G - 1) / CV_G;
        for (int g = 0; g < ngroups; ++g) {
            const int b = g & 1;
            mbar_wait(&dfull[b], (g >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < HB / 16; ++hh) {
                uint32_t v[16];
                tmem_ld16(tmem_base + lane_off + b * NT + half * HB + hh * 16, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[hh * 16 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dfree[b]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TM_COLS>(tmem_base);
    }
    if (warp < 2) return;

    // ---- epilogue: this thread's voxel, HB consecutive output channels
    int r = row;
    const int x = x0 + r % a.bx; r /= a.bx;
    const int y = y0 + r % a.by; r /= a.by;
    const int z = z0 + r % a.bz; r /= a.bz;
    const int b = b0 + r;
    const bool valid = b < a.B;
    const int c0 = n0 + half * HB;
#pragma unroll
    for (int j = 0; j < HB; ++j) {
        float v = acc[j];
        if (a.bias) v += __ldg(a.bias + c0 + j);
        if (a.relu) v = fmaxf(v, 0.f);
        acc[j] = valid ? v : 0.f;
    }
    if (valid) {
        const int us = a.up ? 2 : 1;
        float *o = a.out + ((((size_t)b * (a.Z * us) + (z * us + pz)) * (a.Y * us) + (y * us + py)) * (a.X * us) + (x * us + px)) * a.Cout + c0;
#pragma unroll
        for (int j = 0; j < HB; j += 4) st4(o + j, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
    }
    if (a.stats) {
        // the 32 voxels of a warp belong to one sample (a box holds >= 32 voxels per sample)
        const int bw = __shfl_sync(0xffffffffu, b, 0);
        const bool any = __shfl_sync(0xffffffffu, (int)valid, 0) != 0;
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            const float s = warp_sum(acc[j]), q = warp_sum(acc[j] * acc[j]);
            if (lane == 0 && any) {
                double *st = a.stats + ((size_t)bw * a.Cout + c0 + j) * 2;
                atomicAdd(st, (double)s);
                atomicAdd(st + 1, (double)q);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Elementwise pass between two convolutions.  dst[b, z, y, x, c] for c < C0 comes from src0[b, z >> sh0, ...][c], for c >= C0 from
// src1[b, z >> sh1, ...][c - C0] (channel concatenation; sh = 1: nearest-neighbour x2 upsampling on load).  With groups > 0 the value is
// GroupNorm-ed over the concatenated channels: statistics of group g = channels [g * C / groups, ...) come from the per-channel
// (sum, sum of squares) of the SOURCES (an upsampled source has the statistics of its low-resolution tensor), eps 1e-5, affine
// gamma / beta.  dst_lo (optional) receives the low part of the TF32 operand split of dst.
struct PrepArgs {
    const float *src0, *src1;
    const double *st0, *st1;     // (B, C0, 2) / (B, C1, 2) per-channel sums of the sources, or NULL when groups == 0
    const float *gamma, *beta;   // (C0 + C1)
    float *dst, *dst_lo;
    int B, Z, Y, X, C0, C1, sh0, sh1, groups;
    double n0, n1;               // voxels per sample behind st0 / st1
};

constexpr int PREP_MAXC = 768;

__global__ void __launch_bounds__(256) conv_prep_kernel(const PrepArgs a) {
    __shared__ float s_scale[PREP_MAXC], s_shift[PREP_MAXC];
    __shared__ float s_mean[32], s_rstd[32];
    const int b = blockIdx.y, C = a.C0 + a.C1, tid = threadIdx.x;
    if (a.groups > 0) {
        const int gs = C / a.groups;
        if (tid < a.groups) {
            double s = 0.0, q = 0.0, n = 0.0;
            for (int c = tid * gs; c < (tid + 1) * gs; ++c) {
                const bool first = c < a.C0;
                const double *st = first ? a.st0 + ((size_t)b * a.C0 + c) * 2 : a.st1 + ((size_t)b * a.C1 + (c - a.C0)) * 2;
                // every source voxel appears (dst voxels / source voxels) times: the factor cancels in mean and variance when
                // both sources have the same replication, otherwise weight the sums so that each counts dst voxels
                const double w = 1.0 / (first ? a.n0 : a.n1);
                s += st[0] * w; q += st[1] * w; n += 1.0;
            }
            const double mean = s / n, var = q / n - mean * mean;
            s_mean[tid] = (float)mean;
            s_rstd[tid] = (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + 1e-5));
        }
        __syncthreads();
        for (int c = tid; c < C; c += blockDim.x) {
            const int g = c / gs;
            const float sc = s_rstd[g] * __ldg(a.gamma + c);
            s_scale[c] = sc;
            s_shift[c] = __ldg(a.beta + c) - s_mean[g] * sc;
        }
        __syncthreads();
    }
    const int C4 = C >> 2;
    const size_t vox = (size_t)a.Z * a.Y * a.X;
    const size_t total = vox * C4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        size_t v = i / C4;
        const int x = (int)(v % a.X); v /= a.X;
        const int y = (int)(v % a.Y);
        const int z = (int)(v / a.Y);
        float4 val;
        if (c < a.C0) {
            const int Z0 = a.Z >> a.sh0, Y0 = a.Y >> a.sh0, X0 = a.X >> a.sh0;
            val = __ldg(reinterpret_cast<const float4 *>(a.src0 + ((((size_t)b * Z0 + (z >> a.sh0)) * Y0 + (y >> a.sh0)) * X0 + (x >> a.sh0)) * a.C0 + c));
        } else {
            const int Z1 = a.Z >> a.sh1, Y1 = a.Y >> a.sh1, X1 = a.X >> a.sh1;
            val = __ldg(reinterpret_cast<const float4 *>(a.src1 + ((((size_t)b * Z1 + (z >> a.sh1)) * Y1 + (y >> a.sh1)) * X1 + (x >> a.sh1)) * a.C1 + (c - a.C0)));
        }
        if (a.groups > 0) {
            val.x = val.x * s_scale[c] + s_shift[c];
            val.y = val.y * s_scale[c + 1] + s_shift[c + 1];
            val.z = val.z * s_scale[c + 2] + s_shift[c + 2];
            val.w = val.w * s_scale[c + 3] + s_shift[c + 3];
        }
        const size_t o = ((size_t)b * vox) * C + i * 4;
        st4(a.dst + o, val);
        if (a.dst_lo) st4(a.dst_lo + o, make_float4(tf32_lo(val.x), tf32_lo(val.y), tf32_lo(val.z), tf32_lo(val.w)));
    }
}

// dst[b, v, c] = max over the win^3 window of src (win = 1: copy, or no store at all when dst == NULL) + per-channel sums of
// the result.  One thread per channel, a CTA covers a run of output voxels of one sample.
__global__ void __launch_bounds__(256) pool_stats_kernel(const float *__restrict__ src, float *dst, double *stats, int Zo, int Yo, int Xo, int C,
                                                         int win, int vox_per_cta) {
    const int b = blockIdx.y, c = threadIdx.x;
    if (c >= C) return;
    const int vox = Zo * Yo * Xo, Zi = Zo * win, Yi = Yo * win, Xi = Xo * win;
    const int v0 = blockIdx.x * vox_per_cta, v1 = min(vox, v0 + vox_per_cta);
    double s = 0.0, q = 0.0;
    for (int v = v0; v < v1; ++v) {
        const int x = v % Xo, y = (v / Xo) % Yo, z = v / (Xo * Yo);
        float m = -INFINITY;
        for (int dz = 0; dz < win; ++dz)
            for (int dy = 0; dy < win; ++dy)
                for (int dx = 0; dx < win; ++dx)
                    m = fmaxf(m, __ldg(src + ((((size_t)b * Zi + z * win + dz) * Yi + y * win + dy) * Xi + x * win + dx) * C + c));
        if (dst) dst[((size_t)b * vox + v) * C + c] = m;
        s += (double)m; q += (double)m * (double)m;
    }
    atomicAdd(stats + ((size_t)b * C + c) * 2, s);
    atomicAdd(stats + ((size_t)b * C + c) * 2 + 1, q);
}

// out[b, v, :] = codebook[idx[b, v], :] (channels-last), + per-channel sums
__global__ void __launch_bounds__(256) gather_codes_cl_kernel(const int64_t *__restrict__ idx, const float *__restrict__ codebook, float *out,
                                                              double *stats, int cells, int C, int n_codes, int vox_per_cta) {
    const int b = blockIdx.y, c = threadIdx.x;
    if (c >= C) return;
    const int v0 = blockIdx.x * vox_per_cta, v1 = min(cells, v0 + vox_per_cta);
    double s = 0.0, q = 0.0;
    for (int v = v0; v < v1; ++v) {
        int64_t k = idx[(size_t)b * cells + v];
        k = k < 0 ? 0 : (k >= n_codes ? n_codes - 1 : k);
        const float m = __ldg(codebook + (size_t)k * C + c);
        out[((size_t)b * cells + v) * C + c] = m;
        s += (double)m; q += (double)m * (double)m;
    }
    atomicAdd(stats + ((size_t)b * C + c) * 2, s);
    atomicAdd(stats + ((size_t)b * C + c) * 2 + 1, q);
}

// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*CvEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CvEncodeFn cv_encode_fn() {
    static CvEncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<CvEncodeFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
static int cv_encode(TensorMapBlob *out, const float *p, int rank, const cuuint64_t *dims, const cuuint32_t *box) {
    CvEncodeFn enc = cv_encode_fn();
    if (!enc) { set_cuda_error(cudaErrorNotSupported, "cuTensorMapEncodeTiled entry point"); return SFB200_E_CUDA; }
    if (!p || (reinterpret_cast<uintptr_t>(p) & 15)) return SFB200_E_ARG;
    cuuint64_t strides[4];
    cuuint64_t acc = 4;
    for (int i = 0; i < rank - 1; ++i) { acc *= dims[i]; strides[i] = acc; }
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float *>(p), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (conv)"); return SFB200_E_CUDA; }
    return SFB200_OK;
}

// voxel box of an M tile: x fastest, 128 voxels in total (two samples per tile when the volume has only 64 voxels)
static bool cv_box(int Z, int Y, int X, int *bx, int *by, int *bz, int *bn) {
    int rem = 128;
    *bx = X < 8 ? X : 8; rem /= *bx;
    *by = Y < rem ? Y : rem; if (*by > 4 && rem > 4 && Z >= rem / 4) *by = 4;
    rem /= *by;
    *bz = Z < rem ? Z : rem; rem /= *bz;
    *bn = rem;
    if (*bx * *by * *bz * *bn != 128 || X % *bx || Y % *by || Z % *bz) return false;
    if (*bn > 1 && *bx * *by * *bz < 32) return false;     // a warp of the epilogue must stay inside one sample
    return true;
}

template <int NT>
static int launch_conv_t(ConvArgs &a, const float *w, const float *w_lo, cudaStream_t stream) {
    const cuuint64_t wd[2] = {(cuuint64_t)a.Cin, (cuuint64_t)(a.up ? 8 : 1) * a.taps * a.Cout};
    const cuuint32_t wb[2] = {32, NT};
    SFB_TRY(cv_encode(&a.w, w, 2, wd, wb));
    SFB_TRY(cv_encode(&a.w_lo, w_lo, 2, wd, wb));
    static unsigned long long attr_done = 0;   // bit per device
    if (first_use_on_device(attr_done))
        SFB_CUDA_TRY(cudaFuncSetAttribute(conv3d_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, CvCfg<NT>::SMEM));
    const int tiles = (a.X / a.bx) * (a.Y / a.by) * (a.Z / a.bz) * ((a.B + a.bn - 1) / a.bn);
    // no PDL attribute: the kernel reads its inputs without a dependency wait (plain stream order)
    conv3d_tc_kernel<NT><<<dim3(tiles, a.Cout / NT, a.up ? 8 : 1), dim3(CV_THREADS), CvCfg<NT>::SMEM, stream>>>(a);
    return check_launch("conv3d_tc");
}

int launch_conv3d_tc(const float *in, const float *in_lo, const float *w, const float *w_lo, const float *bias, float *out, double *stats,
                     int B, int Z, int Y, int X, int Cin, int Cout, int taps, int relu, cudaStream_t stream) {
    // taps: 27 = 3x3x3 pad 1; 1 = 1x1x1; 8 = sub-pixel form of [nearest x2 upsampling -> 3x3x3 pad 1] (Z, Y, X = INPUT size)
    if (!in || !in_lo || !w || !w_lo || !out || B < 1 || Cin < 32 || Cin % 32 || Cout < 32 || Cout % 32 ||
        (taps != 27 && taps != 1 && taps != 8))
        return SFB200_E_ARG;
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.bias = bias; a.out = out; a.stats = stats; a.B = B; a.Z = Z; a.Y = Y; a.X = X; a.Cin = Cin; a.Cout = Cout; a.taps = taps; a.relu = relu;
    a.up = taps == 8;
    if (!cv_box(Z, Y, X, &a.bx, &a.by, &a.bz, &a.bn)) return SFB200_E_ARG;
    const cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)B};
    const cuuint32_t box[5] = {32, (cuuint32_t)a.bx, (cuuint32_t)a.by, (cuuint32_t)a.bz, (cuuint32_t)a.bn};
    SFB_TRY(cv_encode(&a.in, in, 5, dims, box));
    SFB_TRY(cv_encode(&a.in_lo, in_lo, 5, dims, box));
    if (Cout % 128 == 0) return launch_conv_t<128>(a, w, w_lo, stream);
    if (Cout == 64) return launch_conv_t<64>(a, w, w_lo, stream);
    if (Cout == 32) return launch_conv_t<32>(a, w, w_lo, stream);
    return SFB200_E_ARG;
}

int launch_conv_prep(const float *src0, int C0, int sh0, const double *st0, double n0, const float *src1, int C1, int sh1, const double *st1,
                     double n1, const float *gamma, const float *beta, int groups, float *dst, float *dst_lo, int B, int Z, int Y, int X,
                     cudaStream_t stream) {
    const int C = C0 + C1;
    if (!src0 || !dst || C0 < 4 || C0 % 4 || C1 % 4 || (C1 > 0 && !src1) || C > PREP_MAXC || B < 1) return SFB200_E_ARG;
    if (groups > 0 && (groups > 32 || C % groups || !st0 || (C1 > 0 && !st1) || !gamma || !beta)) return SFB200_E_ARG;
    PrepArgs a;
    a.src0 = src0; a.src1 = src1; a.st0 = st0; a.st1 = st1; a.gamma = gamma; a.beta = beta; a.dst = dst; a.dst_lo = dst_lo;
    a.B = B; a.Z = Z; a.Y = Y; a.X = X; a.C0 = C0; a.C1 = C1; a.sh0 = sh0; a.sh1 = sh1; a.groups = groups; a.n0 = n0; a.n1 = n1;
    const size_t total = (size_t)Z * Y * X * (C / 4);
    size_t blocks = (total + 256 * 8 - 1) / (256 * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 4) blocks = 148 * 4;
    conv_prep_kernel<<<dim3((unsigned)blocks, B), dim3(256), 0, stream>>>(a);
    return check_launch("conv_prep");
}

int launch_pool_stats(const float *src, float *dst, double *stats, int B, int Zo, int Yo, int Xo, int C, int win, cudaStream_t stream) {
    if (!src || !stats || C < 1 || C > 256 || (win != 1 && win != 2) || B < 1) return SFB200_E_ARG;
    const int vox = Zo * Yo * Xo, per = 32;
    pool_stats_kernel<<<dim3((vox + per - 1) / per, B), dim3(C <= 128 ? 128 : 256), 0, stream>>>(src, dst, stats, Zo, Yo, Xo, C, win, per);
    return check_launch("pool_stats");
}

int launch_gather_codes_cl(const int64_t *idx, const float *codebook, float *out, double *stats, int B, int cells, int C, int n_codes,
                           cudaStream_t stream) {
    if (!idx || !codebook || !out || !stats || C < 1 || C > 256 || B < 1) return SFB200_E_ARG;
    const int per = 32;
    gather_codes_cl_kernel<<<dim3((cells + per - 1) / per, B), dim3(C <= 128 ? 128 : 256), 0, stream>>>(idx, codebook, out, stats, cells, C,
                                                                                                   n_codes, per);
    return check_launch("gather_codes_cl");
}

}  // namespace sfb
