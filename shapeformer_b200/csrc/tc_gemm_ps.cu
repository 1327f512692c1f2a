// Decode-step nn.Linear (M <= 64 activation rows) from PRE-SPLIT weight tiles: the leanest tcgen05 form of the 3xTF32 GEMM.
//
// Offline (sfb200_ar_pretile) every GEMM weight is rewritten as TF32 hi/lo tiles in the exact shared-memory image the tensor
// core reads: Wt[n_tile][k_chunk][hi|lo][128 rows x 32 k, SWIZZLE_128B] — 32 KB per (tile, chunk), contiguous.  In the kernel
//   loader warp : ONE cp.async.bulk (TMA engine, mbarrier complete_tx) per chunk brings hi+lo straight into an operand stage
//                 — issued before the programmatic-dependency wait, the weights do not depend on the previous kernel;
//   split warps : only the small activation chunk (64 x 32) is split into hi/lo operand tiles (register double-buffered
//                 loads from L2), and the TMEM accumulator is promoted to fp32 registers every 2 chunks;
//   MMA warp    : 12 tcgen05.mma.kind::tf32 per chunk with BOTH operands from shared memory;
//   epilogue    : cluster split-K through distributed shared memory, as in tc_gemm.cu.
// Costs 2x the weight bytes in HBM (2.4 GB more for the shipped model) and 2x weight traffic for M > 4 rows — where the step
// is bound by per-kernel latency, not by HBM bandwidth; the fp32 blob stays the source for M <= 4 (GEMV).
#include <cooperative_groups.h>

#include "ar_kernels.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace sfb {

using namespace tc;

constexpr int PS_BN = 64;            // activation rows per tile (UMMA N)
constexpr int PS_NS = 4;             // operand stages (weights and activations share the ring index)
constexpr int PS_G = 2;              // chunks per promotion group
constexpr int PS_THREADS = 192;      // warps 0-3: activation split + promotion + epilogue, warp 4: MMA, warp 5: loader
constexpr int PS_W_STAGE = 2 * 128 * 32 * 4;   // hi + lo weight tiles: 32 KB
constexpr int PS_X_TILE = PS_BN * 32 * 4;      // 8 KB
constexpr int PS_OFF_W = 0;
constexpr int PS_OFF_XH = PS_OFF_W + PS_NS * PS_W_STAGE;
constexpr int PS_OFF_XL = PS_OFF_XH + PS_NS * PS_X_TILE;
constexpr int PS_OFF_BAR = PS_OFF_XL + PS_NS * PS_X_TILE;
constexpr int PS_SMEM = PS_OFF_BAR + 256;

// Optional timeline probe (development aid): when a buffer is registered, CTA (0,0) stamps %globaltimer at the
// phase boundaries of the kernel.  slots: 0 entry, 1 setup done, 2 dependency wait done, 3 first activation tile handed over,
// 4 first MMA issued, 5 last MMA committed, 6 accumulator drained, 7 partials exchanged (cluster.sync), 8 exit.
__device__ unsigned long long *g_ps_timeline = nullptr;
__device__ __forceinline__ void ps_stamp(int slot) {
    if (g_ps_timeline && blockIdx.x == 0 && blockIdx.y == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_ps_timeline[slot] = t;
    }
}

__device__ __forceinline__ float gelu_erf_ps(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// ---- offline: W (N, K) fp32 row-major -> hi/lo swizzled tiles; grid (K/32, ceil(N/128)), 256 threads
__global__ void __launch_bounds__(256) tc_pretile_kernel(const float *__restrict__ W, float *__restrict__ Wt, int N, int K) {
    const int kc = blockIdx.x, nt = blockIdx.y;
    float *hi = Wt + ((size_t)nt * (K / 32) + kc) * (2 * 4096), *lo = hi + 4096;
    for (int e = threadIdx.x; e < 4096; e += 256) {
        const int r = e >> 5, k = e & 31, n = nt * 128 + r;
        const float v = n < N ? W[(size_t)n * K + kc * 32 + k] : 0.f;
        uint32_t h, l;
        split_tf32(v, h, l);
        const int idx = r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
        hi[idx] = __uint_as_float(h);
        lo[idx] = __uint_as_float(l);
    }
}

__global__ void __launch_bounds__(PS_THREADS, 1)
tc_linear_ps_kernel(const float *x, const float *__restrict__ Wt, const float *__restrict__ bias, const float *residual, float *y,
                    int M, int N, int K, int act, int splits) {
    pdl_trigger();
    if (threadIdx.x == 0) ps_stamp(0);
    extern __shared__ __align__(1024) unsigned char ps_smem[];
    unsigned char *smem = ps_smem;
    uint64_t *wfull = reinterpret_cast<uint64_t *>(smem + PS_OFF_BAR);   // [NS] weight stage landed (TMA tx bytes)
    uint64_t *xfull = wfull + PS_NS;                                     // [NS] activation hi/lo tiles written
    uint64_t *done = xfull + PS_NS;                                      // [NS] MMAs of the chunk finished -> stage free
    uint64_t *dfull = done + PS_NS;                                      // [2]  promotion group finished in D[b]
    uint64_t *dfree = dfull + 2;                                         // [2]  D[b] drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dfree + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int n0 = blockIdx.x * 128, sp = blockIdx.y;
    const int nch_total = K / 32;
    const int c_beg = (int)(((long long)sp * nch_total) / splits), c_end = (int)(((long long)(sp + 1) * nch_total) / splits);
    const int nch = c_end - c_beg;
    const int ngroups = (nch + PS_G - 1) / PS_G;
    constexpr uint32_t IDESC = instr_desc(2, 128, PS_BN);
    const float *wt_tile = Wt + ((size_t)blockIdx.x * nch_total + c_beg) * (2 * 4096);   // chunk c at + c * 8192 floats

    if (tid == 0) {
        for (int i = 0; i < PS_NS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&xfull[i], 128); mbar_init(&done[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dfree[i], 128); }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) ps_stamp(1);

    float acc[PS_BN];   // split warps: acc[j] = D[row tid][col j]
#pragma unroll
    for (int j = 0; j < PS_BN; ++j) acc[j] = 0.f;

    if (warp < 4) {
        // ================================ activation split + promotion ================================
        const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
        auto drain = [&](int g) {
            const int b = g & 1;
            mbar_wait(&dfull[b], (g >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < PS_BN / 32; ++h) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_off + b * PS_BN + h * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[h * 32 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            mbar_arrive(&dfree[b]);
        };
        // thread -> 4 float4 of the 64 x 32 chunk: element idx = tid + 128 j: row = idx >> 3, 16-byte chunk = idx & 7
        auto load_x = [&](int c, float4 (&v)[4]) {
            const int k0 = (c_beg + c) * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = tid + 128 * j, r = idx >> 3, ch = idx & 7;
                v[j] = r < M ? ld4(x + (size_t)r * K + k0 + ch * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        pdl_wait();
        if (tid == 0) ps_stamp(2);
        float4 cur[4], nxt[4];
        if (nch > 0) load_x(0, cur);
        for (int i = 0; i < nch; ++i) {
            if (i + 1 < nch) load_x(i + 1, nxt);
            const int s = i % PS_NS;
            if (i >= PS_NS) mbar_wait(&done[s], ((i / PS_NS) - 1) & 1);
            float *xh = reinterpret_cast<float *>(smem + PS_OFF_XH + s * PS_X_TILE);
            float *xl = reinterpret_cast<float *>(smem + PS_OFF_XL + s * PS_X_TILE);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = tid + 128 * j, r = idx >> 3, ch = idx & 7;
                const int o = r * 32 + ((ch ^ (r & 7)) << 2);
                uint32_t h[4], l[4];
                split_tf32(cur[j].x, h[0], l[0]); split_tf32(cur[j].y, h[1], l[1]);
                split_tf32(cur[j].z, h[2], l[2]); split_tf32(cur[j].w, h[3], l[3]);
                *reinterpret_cast<uint4 *>(xh + o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4 *>(xl + o) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&xfull[s]);
            if (tid == 0 && i == 0) ps_stamp(3);
            if ((i % PS_G) == PS_G - 1 || i == nch - 1) {
                const int g = i / PS_G;
                if (g >= 1) drain(g - 1);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
        }
        drain(ngroups - 1);
        if (tid == 0) ps_stamp(6);
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        pdl_wait();
        {   // the whole warp stays converged, one elected lane issues (operands stay in uniform registers)
            const uint32_t w0 = smem_u32(smem + PS_OFF_W), xh0 = smem_u32(smem + PS_OFF_XH), xl0 = smem_u32(smem + PS_OFF_XL);
            for (int i = 0; i < nch; ++i) {
                const int s = i % PS_NS, g = i / PS_G, b = g & 1;
                const bool first = (i % PS_G) == 0;
                if (first && g >= 2) mbar_wait(&dfree[b], ((g >> 1) - 1) & 1);
                mbar_wait(&wfull[s], (i / PS_NS) & 1);
                mbar_wait(&xfull[s], (i / PS_NS) & 1);
                tc_fence_after();
                if (i == 0 && (tid & 31) == 0) ps_stamp(4);
                const uint32_t d = tmem_base + b * PS_BN;
                const uint64_t ah = smem_desc_k128(w0 + s * PS_W_STAGE), al = smem_desc_k128(w0 + s * PS_W_STAGE + 16384);
                const uint64_t bh = smem_desc_k128(xh0 + s * PS_X_TILE), bl = smem_desc_k128(xl0 + s * PS_X_TILE);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        mma_tf32_ss(d, al + 2 * k, bh + 2 * k, IDESC, !(first && k == 0));
                        mma_tf32_ss(d, ah + 2 * k, bl + 2 * k, IDESC, 1);
                        mma_tf32_ss(d, ah + 2 * k, bh + 2 * k, IDESC, 1);
                    }
                    mma_commit(&done[s]);
                    if ((i % PS_G) == PS_G - 1 || i == nch - 1) mma_commit(&dfull[b]);
                }
                __syncwarp();
            }
            if ((tid & 31) == 0) ps_stamp(5);
        }
    } else {
        // ================================ weight loader (TMA bulk copies) ================================
        // weights are constants: no dependency wait before streaming them
        if ((tid & 31) == 0) {
            for (int i = 0; i < nch; ++i) {
                const int s = i % PS_NS;
                if (i >= PS_NS) mbar_wait(&done[s], ((i / PS_NS) - 1) & 1);
                mbar_expect_tx(&wfull[s], PS_W_STAGE);
                bulk_load(smem + PS_OFF_W + s * PS_W_STAGE, wt_tile + (size_t)i * (2 * 4096), PS_W_STAGE, &wfull[s]);
            }
        }
        __syncwarp();
        pdl_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<128>(tmem_base);
    }

    // ================================ epilogue (same scheme as tc_gemm.cu) ================================
    if (splits == 1) {
        if (warp < 4) {
            const int n = n0 + tid;
            const float bv = (bias && n < N) ? bias[n] : 0.f;
#pragma unroll
            for (int j = 0; j < PS_BN; ++j) {
                if (j < M && n < N) {
                    float r = acc[j] + bv;
                    if (act == 1) r = gelu_erf_ps(r);
                    if (residual) r += residual[(size_t)j * N + n];
                    y[(size_t)j * N + n] = r;
                }
            }
        }
    } else {
        cg::cluster_group cluster = cg::this_cluster();
        float *red = reinterpret_cast<float *>(smem);   // [BN][128] partial tile, reusing the weight stages
        if (warp < 4) {
#pragma unroll
            for (int j = 0; j < PS_BN; ++j) red[j * 128 + tid] = acc[j];
        }
        cluster.sync();
        if (tid == 0) ps_stamp(7);
        const int rank = (int)cluster.block_rank();
        const float *peer[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) peer[s] = cluster.map_shared_rank(red, s < splits ? s : 0);
        constexpr int TOTAL4 = PS_BN * 128 / 4;
        const int per4 = (TOTAL4 + splits - 1) / splits;
        const int e_beg = rank * per4, e_end = min(TOTAL4, e_beg + per4);
        for (int e4 = e_beg + tid; e4 < e_end; e4 += PS_THREADS) {
            float4 v[16];
#pragma unroll
            for (int s = 0; s < 16; ++s)
                if (s < splits) v[s] = ld4(peer[s] + 4 * e4);
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int s = 0; s < 16; ++s)
                if (s < splits) { r.x += v[s].x; r.y += v[s].y; r.z += v[s].z; r.w += v[s].w; }
            const int m = e4 >> 5, nn = n0 + ((e4 & 31) << 2);
            if (m < M) {
                float o[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (nn + q < N) {
                        float t = o[q];
                        if (bias) t += bias[nn + q];
                        if (act == 1) t = gelu_erf_ps(t);
                        const size_t off = (size_t)m * N + nn + q;
                        if (residual) t += residual[off];
                        y[off] = t;
                    }
                }
            }
        }
        cluster.sync();
    }
    if (tid == 0) ps_stamp(8);
}

// ---------------------------------------------------------------------------------------------------------------------
int set_ps_timeline(unsigned long long *buf) {
    SFB_CUDA_TRY(cudaMemcpyToSymbol(g_ps_timeline, &buf, sizeof(buf)));
    return SFB200_OK;
}

int64_t tc_pretiled_floats(int N, int K) { return (int64_t)((N + 127) / 128) * 128 * K * 2; }

int launch_tc_pretile(const float *W, float *Wt, int N, int K, cudaStream_t s) {
    if (N <= 0 || K <= 0 || K % 32 != 0) return SFB200_E_ARG;
    tc_pretile_kernel<<<dim3(K / 32, (N + 127) / 128), 256, 0, s>>>(W, Wt, N, K);
    return check_launch("tc_pretile");
}

static int ps_max_cluster(int tiles) {
    static int cache[17] = {0};
    for (int S = 16; S >= 2; --S) {
        if (cache[S] == 0) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(1, S, 1);
            cfg.blockDim = dim3(PS_THREADS);
            cfg.dynamicSmemBytes = PS_SMEM;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = S; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, tc_linear_ps_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = -1; }
            cache[S] = n > 0 ? n : -1;
        }
        if (cache[S] >= tiles) return S;
    }
    return 1;
}

int launch_linear_tc_ps(const float *x, const float *Wt, const float *bias, const float *residual, float *y, int M, int N, int K,
                        int act, cudaStream_t stream) {
    if (M <= 0 || M > PS_BN || N <= 0 || K <= 0 || K % 32 != 0) return SFB200_E_ARG;
    static unsigned long long attr_done = 0;   // bit per device
    if (first_use_on_device(attr_done)) {
        SFB_CUDA_TRY(cudaFuncSetAttribute(tc_linear_ps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM));
        SFB_CUDA_TRY(cudaFuncSetAttribute(tc_linear_ps_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    const int tiles = (N + 127) / 128, nch = K / 32;
    int s = 148 / tiles;
    if (s < 1) s = 1;
    if (s > nch / 2) s = nch / 2 > 0 ? nch / 2 : 1;
    if (s > 1) {
        const int cap = ps_max_cluster(tiles);
        if (s > cap) s = cap;
    }
    return launch_ex("tc_linear_ps", tc_linear_ps_kernel, dim3(tiles, s, 1), dim3(PS_THREADS), PS_SMEM, stream, dim3(1, s, 1), x, Wt, bias,
                     residual, y, M, N, K, act, s);
}

}  // namespace sfb
