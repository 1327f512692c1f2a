// y = act(x W^T + bias) + residual for LARGE row counts (decode batches of > 64 rows, prefill) on the 5th-generation tensor cores
// with fp32-level accuracy ("3xTF32"), as a pure TMA -> shared memory -> tcgen05.mma pipeline: no per-element work on the
// operands inside the kernel (nn.Linear of Block.forward / the heads, transformer/mingpt.py:74-111,222-231).
//
// Operand split.  v = hi + lo with hi = trunc_tf32(v) — which costs nothing: the tensor core reads only the upper 19 bits of a
// 32-bit tf32 operand, so the RAW fp32 array is the hi operand — and lo = rna_tf32(v - hi) kept as a second fp32 array by whoever
// produced v (the weights once at load time: sfb200_ar_set_lo_weights; activations by the producing kernel's epilogue or by
// split_lo_kernel).  Products lo*hi + hi*lo + hi*hi, the dropped lo*lo term is <= 2^-20 relative.
//
// Swap-AB mapping as in tc_gemm.cu: UMMA M = 128 output features (rows of W), UMMA N = BN activation rows:
//     D[n, m] (+)= sum_k W[n, k] * x[m, k]          A = W tile (128 x 32), B = x tile (BN x 32), both K-major SWIZZLE_128B.
// Per 32-wide K chunk one stage of the ring holds  W hi | W lo | x hi | x lo  (4 TMA tile loads, one mbarrier), and one thread
// issues 12 tcgen05.mma.kind::tf32 (4 k-steps x 3 products, both operands from shared memory).
// The tensor-core accumulator does not round to nearest, so accumulation chains are kept at 24 MMAs: two TMEM accumulators
// alternate and 8 epilogue warps drain the finished one into fp32 registers (IEEE adds) while the other accumulates.
// Split-K (small M x N grids, at most one CTA per SM) goes through L2: every split stores its partial tile, waits for its
// siblings at the tile's arrival counter, then sums the partials in split order (deterministic) for ITS slice of the rows and
// runs the epilogue there — the reduction is spread over the splits; no cluster, no second kernel.
#include <cuda.h>
#include <string.h>

#include "ar_kernels.cuh"
#include "tc_common.cuh"

namespace sfb {

using namespace tc;

constexpr int BG_THREADS = 320;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: promotion + epilogue
constexpr int BG_G = 2;                  // chunks per promotion group (24 MMAs per TMEM accumulation chain; 4 costs 3e-5 on the logits)
constexpr int BG_A_TILE = 128 * 32 * 4;  // 16 KB

template <int BN>
struct BgCfg {
    static constexpr int NS = BN <= 128 ? 3 : 2;          // ring stages (227 KB limit)
    static constexpr int B_TILE = BN * 32 * 4;
    static constexpr int STAGE = 2 * BG_A_TILE + 2 * B_TILE;
    static constexpr int OFF_BAR = NS * STAGE;
    static constexpr int SMEM = OFF_BAR + 256;
    static constexpr int TM_COLS = 2 * BN;                // two accumulators
};

struct BigArgs {
    TensorMapBlob w, wlo, x, xlo;
    const float *bias, *residual;
    float *y, *y_lo;
    float *partial;          // split-K scratch: [tile][split][BN][128]
    int *tile_cnt;           // 2 x BG_MAX_TILES counters (arrived | finished) per tile, zero between launches
    int M, N, K, act, splits;
};

__device__ __forceinline__ void bg_tma_2d(void *smem_dst, const void *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
                     "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bg_prefetch_desc(const void *map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void bg_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ float bg_gelu(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

template <int BN>
__global__ void __launch_bounds__(BG_THREADS, 1) tc_big_linear_kernel(const __grid_constant__ BigArgs a) {
    using C = BgCfg<BN>;
    constexpr int NS = C::NS;
    extern __shared__ __align__(1024) unsigned char bg_smem[];
    unsigned char *smem = bg_smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);   // [NS] stage landed (TMA tx bytes)
    uint64_t *empty = full + NS;                                        // [NS] MMAs that read the stage finished
    uint64_t *dfull = empty + NS;                                       // [2]  promotion group finished in D[b]
    uint64_t *dfree = dfull + 2;                                        // [2]  D[b] drained by the 8 epilogue warps
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dfree + 2);

    pdl_trigger();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * 128, sp = blockIdx.y, m0 = blockIdx.z * BN;
    const int nch_total = a.K >> 5;
    const int c_beg = (int)(((long long)sp * nch_total) / a.splits), c_end = (int)(((long long)(sp + 1) * nch_total) / a.splits);
    const int nch = c_end - c_beg;

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
            bg_prefetch_desc(&a.w); bg_prefetch_desc(&a.wlo); bg_prefetch_desc(&a.x); bg_prefetch_desc(&a.xlo);
            // the weights do not depend on the previous kernel: their tiles of the first stages are requested before the
            // programmatic-dependency wait, the activation tiles after it
            const int pre = nch < NS ? nch : NS;
            for (int i = 0; i < pre; ++i) {
                unsigned char *st = smem + i * C::STAGE;
                bg_expect_tx(&full[i], C::STAGE);
                bg_tma_2d(st, &a.w, (c_beg + i) * 32, n0, &full[i]);
                bg_tma_2d(st + BG_A_TILE, &a.wlo, (c_beg + i) * 32, n0, &full[i]);
            }
            pdl_wait();
            asm volatile("fence.proxy.async;\n" ::: "memory");
            for (int i = 0; i < pre; ++i) {
                unsigned char *st = smem + i * C::STAGE;
                bg_tma_2d(st + 2 * BG_A_TILE, &a.x, (c_beg + i) * 32, m0, &full[i]);
                bg_tma_2d(st + 2 * BG_A_TILE + C::B_TILE, &a.xlo, (c_beg + i) * 32, m0, &full[i]);
            }
            for (int i = NS; i < nch; ++i) {
                const int s = i % NS;
                mbar_wait(&empty[s], ((i / NS) - 1) & 1);
                unsigned char *st = smem + s * C::STAGE;
                bg_expect_tx(&full[s], C::STAGE);
                bg_tma_2d(st, &a.w, (c_beg + i) * 32, n0, &full[s]);
                bg_tma_2d(st + BG_A_TILE, &a.wlo, (c_beg + i) * 32, n0, &full[s]);
                bg_tma_2d(st + 2 * BG_A_TILE, &a.x, (c_beg + i) * 32, m0, &full[s]);
                bg_tma_2d(st + 2 * BG_A_TILE + C::B_TILE, &a.xlo, (c_beg + i) * 32, m0, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp converged, one elected lane issues) ================================
        constexpr uint32_t IDESC = instr_desc(2, 128, BN);
        const uint32_t ring0 = smem_u32(smem);
        for (int i = 0; i < nch; ++i) {
            const int s = i % NS, g = i / BG_G, b = g & 1;
            const bool first = (i % BG_G) == 0, last = (i % BG_G) == BG_G - 1 || i == nch - 1;
            if (first && g >= 2) mbar_wait(&dfree[b], ((g >> 1) - 1) & 1);      // D[b] drained (group g-2)
            mbar_wait(&full[s], (i / NS) & 1);
            tc_fence_after();
            const uint32_t d = tmem_base + b * BN;
            const uint32_t st = ring0 + s * C::STAGE;
            const uint64_t ah = smem_desc_k128(st), al = smem_desc_k128(st + BG_A_TILE);
            const uint64_t bh = smem_desc_k128(st + 2 * BG_A_TILE), bl = smem_desc_k128(st + 2 * BG_A_TILE + C::B_TILE);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // small products first, the dominant hi*hi term last
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
    constexpr int HB = BN / 2;
    float acc[HB];                 // acc[j] = D[feature row][half * HB + j]
    const int ew = warp - 2;       // 0..7 for the epilogue warps
    const int quad = warp & 3, half = ew >> 2;
    const int row = quad * 32 + lane;
    if (warp >= 2) {
        pdl_wait();      // residual / partials / y belong to earlier kernels (the operands arrive through the producer's wait)
#pragma unroll
        for (int j = 0; j < HB; ++j) acc[j] = 0.f;
        const uint32_t lane_off = (uint32_t)(32 * quad) << 16;
        const int ngroups = (nch + BG_G - 1) / BG_G;
        for (int g = 0; g < ngroups; ++g) {
            const int b = g & 1;
            mbar_wait(&dfull[b], (g >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < HB / 32; ++hh) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_off + b * BN + half * HB + hh * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[hh * 32 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dfree[b]);
        }
    }
    tc_fence_before();
    __syncthreads();      // all MMAs retired and drained: the ring and TMEM are idle
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TM_COLS>(tmem_base);
    }

    // ---- split-K: every split stores its partial tile; once all splits of the tile have arrived (they are co-resident: the
    //      grid never exceeds one CTA per SM when splits > 1) each of them reduces and finishes its own slice of the activation
    //      rows, summing the partials in split order (deterministic)
    const int tile = blockIdx.z * gridDim.x + blockIdx.x;
    if (a.splits > 1) {
        float *part = a.partial + ((size_t)tile * a.splits + sp) * BN * 128;
        if (warp >= 2) {
#pragma unroll
            for (int j = 0; j < HB; ++j) __stcg(part + (size_t)(half * HB + j) * 128 + row, acc[j]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            int *ctr = a.tile_cnt + tile;
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;\n" ::"l"(ctr) : "memory");
            int v;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(ctr) : "memory");
            } while (v < a.splits);
        }
        __syncthreads();
        if (warp >= 2) {
            const int per = (BN + a.splits - 1) / a.splits;
            const int c_lo = sp * per, c_hi = min(BN, c_lo + per);
            const float *p0 = a.partial + (size_t)tile * a.splits * BN * 128;
            const int et = tid - 64, n = n0 + (et & 127);
            const float bv = (a.bias && n < a.N) ? __ldg(a.bias + n) : 0.f;
            for (int c = c_lo + (et >> 7); c < c_hi; c += 2) {
                float r = 0.f;
                for (int s2 = 0; s2 < a.splits; ++s2) r += __ldcg(p0 + ((size_t)s2 * BN + c) * 128 + (et & 127));
                const int m = m0 + c;
                if (m < a.M && n < a.N) {
                    r += bv;
                    if (a.act == 1) r = bg_gelu(r);
                    const size_t off = (size_t)m * a.N + n;
                    if (a.residual) r += __ldcg(a.residual + off);
                    a.y[off] = r;
                    if (a.y_lo) a.y_lo[off] = tf32_lo(r);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {      // the last split to finish its slice re-arms the tile's counters for the next launch
            __threadfence();
            if (atomicAdd(a.tile_cnt + BG_MAX_TILES + tile, 1) == a.splits - 1) {
                a.tile_cnt[tile] = 0;
                a.tile_cnt[BG_MAX_TILES + tile] = 0;
            }
        }
        return;
    }
    // ---- epilogue: bias, activation, residual; y (and the lo part of y when a GEMM consumes it next)
    if (warp >= 2) {
        const int n = n0 + row;
        if (n < a.N) {
            const float bv = a.bias ? __ldg(a.bias + n) : 0.f;
#pragma unroll
            for (int j = 0; j < HB; ++j) {
                const int m = m0 + half * HB + j;
                if (m < a.M) {
                    float r = acc[j] + bv;
                    if (a.act == 1) r = bg_gelu(r);
                    const size_t off = (size_t)m * a.N + n;
                    if (a.residual) r += __ldcg(a.residual + off);
                    a.y[off] = r;
                    if (a.y_lo) a.y_lo[off] = tf32_lo(r);
                }
            }
        }
    }
}

// lo[i] = rna_tf32(x[i] - trunc_tf32(x[i]))
__global__ void __launch_bounds__(256) split_lo_kernel(const float *x, float *lo, size_t n4, size_t n) {
    pdl_trigger();
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) lo[n4 * 4 + threadIdx.x] = tf32_lo(x[n4 * 4 + threadIdx.x]);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(x) + i);
        reinterpret_cast<float4 *>(lo)[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
    }
}

int launch_split_lo(const float *x, float *lo, size_t n, cudaStream_t s) {
    if (!x || !lo || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15)) return SFB200_E_ARG;
    if (n == 0) return SFB200_OK;
    const size_t n4 = n / 4, blocks = (n4 + 255) / 256;
    const int grid = (int)(blocks < 1 ? 1 : (blocks < 148 * 8 ? blocks : 148 * 8));
    return launch_ex("split_lo", split_lo_kernel, dim3(grid), dim3(256), 0, s, dim3(1, 1, 1), x, lo, n4, n);
}

// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*BgEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static BgEncodeFn bg_encode_fn() {
    static BgEncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<BgEncodeFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
// Row-major fp32 matrix (rows, K) -> tensor map with box 32 k x box_rows rows, SWIZZLE_128B, rows past the end read as zeros.
static int bg_map(const float *p, int rows, int K, int box_rows, TensorMapBlob *out) {
    if (!p || rows <= 0 || K <= 0 || K % 32 != 0 || (reinterpret_cast<uintptr_t>(p) & 15)) return SFB200_E_ARG;
    BgEncodeFn enc = bg_encode_fn();
    if (!enc) { set_cuda_error(cudaErrorNotSupported, "cuTensorMapEncodeTiled entry point"); return SFB200_E_CUDA; }
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p), gdim, gstride,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled"); return SFB200_E_CUDA; }
    return SFB200_OK;
}

size_t big_partial_floats() { return (size_t)BG_MAX_UNITS * 128 * 128; }

template <int BN>
static int launch_big_t(BigArgs &a, const float *x, const float *x_lo, cudaStream_t stream) {
    const int n_tiles = (a.N + 127) / 128, m_tiles = (a.M + BN - 1) / BN, tiles = n_tiles * m_tiles, nch = a.K / 32;
    if (m_tiles > 65535) return SFB200_E_ARG;
    int splits = 1;
    const int sms = chain_grid_size() > 0 ? chain_grid_size() : 148;     // splits spin on their siblings: all must be co-resident
    if (tiles < sms) {
        splits = sms / tiles;
        if (splits > nch / 4) splits = nch / 4;
        if (splits > 16) splits = 16;
        if (splits < 1) splits = 1;
        while (splits > 1 && (size_t)tiles * splits * BN * 128 > big_partial_floats()) --splits;
    }
    if (splits > 1 && (!a.partial || !a.tile_cnt || tiles > BG_MAX_TILES)) splits = 1;
    a.splits = splits;
    SFB_TRY(bg_map(x, a.M, a.K, BN, &a.x));
    SFB_TRY(bg_map(x_lo, a.M, a.K, BN, &a.xlo));
    static unsigned long long attr_done = 0;   // bit per device
    if (first_use_on_device(attr_done))
        SFB_CUDA_TRY(cudaFuncSetAttribute(tc_big_linear_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, BgCfg<BN>::SMEM));
    return launch_ex("tc_big_linear", tc_big_linear_kernel<BN>, dim3(n_tiles, splits, m_tiles), dim3(BG_THREADS), BgCfg<BN>::SMEM, stream,
                     dim3(1, 1, 1), a);
}

// x, x_lo (M, K); W, W_lo (N, K); y (M, N); y_lo optional.  partial / tile_cnt: split-K scratch (big_partial_floats() floats,
// 2 * BG_MAX_TILES zeroed ints) or NULL (no split-K).
int launch_linear_big(const float *x, const float *x_lo, const float *W, const float *W_lo, const float *bias, const float *residual,
                      float *y, float *y_lo, int M, int N, int K, int act, float *partial, int *tile_cnt, cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0 || K % 32 != 0 || !x || !x_lo || !W || !W_lo || !y) return SFB200_E_ARG;
    BigArgs a;
    memset(&a, 0, sizeof(a));
    a.bias = bias; a.residual = residual; a.y = y; a.y_lo = y_lo; a.partial = partial; a.tile_cnt = tile_cnt;
    a.M = M; a.N = N; a.K = K; a.act = act;
    SFB_TRY(bg_map(W, N, K, 128, &a.w));
    SFB_TRY(bg_map(W_lo, N, K, 128, &a.wlo));
    return launch_big_t<128>(a, x, x_lo, stream);
}

}  // namespace sfb
