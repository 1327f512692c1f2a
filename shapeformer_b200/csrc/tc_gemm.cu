// y = act(x W^T + bias) + residual on the 5th-generation tensor cores with fp32-level accuracy ("3xTF32").
//
// Every fp32 operand is split into two TF32-representable parts (hi = rna_tf32(v), lo = rna_tf32(v - hi)) and the product is
// accumulated as  lo*hi + hi*lo + hi*hi  (the dropped lo*lo term is ~2^-22 relative).  The tensor-core accumulator does not
// round to nearest, so its error grows with the length of the accumulation chain (measured: ~1e-5 relative after 384 MMAs);
// therefore the TMEM accumulator is PROMOTED every TCG_G chunks: drained with tcgen05.ld and added (IEEE RN) into fp32
// registers while the next group accumulates into the other TMEM buffer.  Net accuracy ~1e-6, which keeps the sampled AR
// tokens identical to the fp32 reference (single-pass TF32 / BF16 would flip them — SURVEY.md App. C-1/C-6).
//
// Swap-AB mapping for skinny activations: the UMMA M dimension (128 TMEM lanes) runs over OUTPUT FEATURES (rows of W), the
// UMMA N dimension over activation rows, so a 64-row decode batch still fills the 128-lane datapath:
//     D[n, m] (+)= sum_k W[n, k] * x[m, k]         A = W tile (128 x 32 per chunk), B = x tile (BN x 32), both K-major.
//
// Data flow per 32-wide K chunk (one 128-byte swizzle row of fp32):
//   cp.async (16 B, coalesced, XOR-swizzled destination)  ->  raw W / x tiles in shared memory (NS-deep ring)
//   warps 0-3: thread t reads W row t (conflict-free through the swizzle), splits it and writes hi|lo straight into TENSOR
//              MEMORY with tcgen05.st (A operand lives in TMEM: no second shared-memory pass for the weights);
//              the x tile is split element-wise into two UMMA-layout shared tiles (hi, lo)
//   warp 4   : one thread issues 12 tcgen05.mma.kind::tf32 per chunk (4 k-steps x 3 products), commits to mbarriers
// Split-K runs as a thread-block cluster (<= 16 CTAs): partial tiles are exchanged through distributed shared memory and
// summed in rank order (deterministic), then bias / GELU / residual are applied once.
#include <cooperative_groups.h>

#include "ar_kernels.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace sfb {

using namespace tc;

constexpr int TCG_NT = 4;   // TMEM A slots == x hi/lo slots (split warps may run this many chunks ahead of the MMAs)
constexpr int TCG_G = 2;    // chunks per promotion group (24 MMAs per TMEM accumulation chain)
constexpr int TCG_THREADS = 288;   // 8 load/split/promote warps (two per W row: K halves) + 1 MMA warp
constexpr int TCG_SPLIT = 256;     // threads in the split warps

template <int BN>
struct TcgSmem {
    static constexpr int NS = BN <= 64 ? 4 : 3;       // raw shared-memory stages (BN = 128 is capped by the 227 KB limit)
    static constexpr int W_TILE = 128 * 32 * 4;       // bytes
    static constexpr int X_TILE = BN * 32 * 4;
    static constexpr int OFF_W = 0;
    static constexpr int OFF_X = OFF_W + NS * W_TILE;
    static constexpr int OFF_XH = OFF_X + NS * X_TILE;
    static constexpr int OFF_XL = OFF_XH + TCG_NT * X_TILE;
    static constexpr int OFF_BAR = OFF_XL + TCG_NT * X_TILE;
    static constexpr int TOTAL = OFF_BAR + 128;
    static_assert(BN * 128 * 4 <= NS * W_TILE + NS * X_TILE, "partial tile must fit in the operand ring");
};

__device__ __forceinline__ float gelu_erf_tc(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

template <int BN>
__global__ void __launch_bounds__(TCG_THREADS, 1)
tc_linear_kernel(const float *x, const float *__restrict__ W, const float *__restrict__ bias,
                 const float *residual, float *y, int M, int N, int K, int act, int splits) {
    using L = TcgSmem<BN>;
    constexpr int TCG_NS = L::NS;
    extern __shared__ __align__(1024) unsigned char tcg_smem[];
    unsigned char *smem = tcg_smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::OFF_BAR);        // [NT] A slot + x hi/lo slot ready
    uint64_t *done = full + TCG_NT;                                           // [NT] MMAs of the chunk finished
    uint64_t *dfull = done + TCG_NT;                                          // [2]  promotion group finished in D[b]
    uint64_t *dfree = dfull + 2;                                              // [2]  D[b] drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dfree + 2);

    pdl_trigger();
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, half = (tid >> 7) & 1;   // split warps: W row / TMEM lane, and which half of K (and of D columns)
    const int n0 = blockIdx.x * 128, m0 = blockIdx.z * BN, sp = blockIdx.y;
    const int nch_total = K / 32;
    const int c_beg = (int)(((long long)sp * nch_total) / splits), c_end = (int)(((long long)(sp + 1) * nch_total) / splits);
    const int nch = c_end - c_beg;
    const int ngroups = (nch + TCG_G - 1) / TCG_G;
    constexpr int TM_COLS = (2 * BN + TCG_NT * 64) <= 256 ? 256 : 512;
    constexpr int A_COL0 = 2 * BN;
    constexpr uint32_t IDESC = instr_desc(2, 128, BN);

    if (tid == 0) {
        for (int i = 0; i < TCG_NT; ++i) { mbar_init(&full[i], TCG_SPLIT); mbar_init(&done[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dfree[i], TCG_SPLIT); }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc<TM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr int HB = BN / 2;   // D columns owned by one thread of a row pair
    float acc[HB];               // split warps: acc[j] = D[row][half * HB + j]
#pragma unroll
    for (int j = 0; j < HB; ++j) acc[j] = 0.f;

    if (warp < 8) {
        // ================================ load + split + promote warps ================================
        auto issue_w = [&](int c, int s) {
            const int k0 = (c_beg + c) * 32;
            float *wdst = reinterpret_cast<float *>(smem + L::OFF_W + s * L::W_TILE);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = tid + TCG_SPLIT * j, r = idx >> 3, ch = idx & 7, n = n0 + r;
                const float *src = W + (size_t)(n < N ? n : N - 1) * K + k0 + ch * 4;
                cp_async16(wdst + r * 32 + ((ch ^ (r & 7)) << 2), src, n < N ? 16 : 0);
            }
        };
        auto issue_x = [&](int c, int s) {
            const int k0 = (c_beg + c) * 32;
            float *xdst = reinterpret_cast<float *>(smem + L::OFF_X + s * L::X_TILE);
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) {
                const int idx = tid + TCG_SPLIT * j, r = idx >> 3, ch = idx & 7, m = m0 + r;
                const float *src = x + (size_t)(m < M ? m : M - 1) * K + k0 + ch * 4;
                cp_async16(xdst + r * 32 + ((ch ^ (r & 7)) << 2), src, m < M ? 16 : 0);
            }
        };
        const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
        // drain promotion group g: acc += D[g & 1]
        auto drain = [&](int g) {
            const int b = g & 1;
            mbar_wait(&dfull[b], (g >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < HB / 32; ++h) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_off + b * BN + half * HB + h * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[h * 32 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            mbar_arrive(&dfree[b]);
        };
        // the weights do not depend on the previous kernel: start streaming them before the programmatic-dependency wait
#pragma unroll
        for (int c = 0; c < TCG_NS - 1; ++c)
            if (c < nch) issue_w(c, c);
        pdl_wait();
#pragma unroll
        for (int c = 0; c < TCG_NS - 1; ++c) {
            if (c < nch) issue_x(c, c);
            cp_async_commit();     // group 0 also carries every prologue W copy
        }
        for (int i = 0; i < nch; ++i) {
            cp_async_wait<TCG_NS - 2>();     // this thread's copies of chunk i have landed
            bar_sync(1, TCG_SPLIT);          // ... everybody's; and everybody finished splitting chunk i-1
            if (i + TCG_NS - 1 < nch) {
                issue_w(i + TCG_NS - 1, (i + TCG_NS - 1) % TCG_NS);
                issue_x(i + TCG_NS - 1, (i + TCG_NS - 1) % TCG_NS);
            }
            cp_async_commit();
            const int slot = i % TCG_NT;
            if (i >= TCG_NT) mbar_wait(&done[slot], ((i / TCG_NT) - 1) & 1);   // MMAs of chunk i-NT released the slot
            tc_fence_after();
            // ---- half of W row `row` (16 of its 32 k) -> hi | lo in tensor memory
            const float *wrow = reinterpret_cast<const float *>(smem + L::OFF_W + (i % TCG_NS) * L::W_TILE) + row * 32;
            {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int ch = half * 4 + c4;
                    const float4 v = ld4(wrow + ((ch ^ (row & 7)) << 2));
                    split_tf32(v.x, hi[4 * c4 + 0], lo[4 * c4 + 0]);
                    split_tf32(v.y, hi[4 * c4 + 1], lo[4 * c4 + 1]);
                    split_tf32(v.z, hi[4 * c4 + 2], lo[4 * c4 + 2]);
                    split_tf32(v.w, hi[4 * c4 + 3], lo[4 * c4 + 3]);
                }
                const uint32_t a_col = tmem_base + lane_off + A_COL0 + slot * 64 + half * 16;
                tmem_st16(a_col, hi);
                tmem_st16(a_col + 32, lo);
            }
            // ---- x tile -> hi / lo shared tiles (same swizzled positions)
            const float *xr = reinterpret_cast<const float *>(smem + L::OFF_X + (i % TCG_NS) * L::X_TILE);
            float *xh = reinterpret_cast<float *>(smem + L::OFF_XH + slot * L::X_TILE);
            float *xl = reinterpret_cast<float *>(smem + L::OFF_XL + slot * L::X_TILE);
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) {
                const int o = (tid + TCG_SPLIT * j) * 4;
                const float4 v = ld4(xr + o);
                uint32_t h[4], l[4];
                split_tf32(v.x, h[0], l[0]); split_tf32(v.y, h[1], l[1]);
                split_tf32(v.z, h[2], l[2]); split_tf32(v.w, h[3], l[3]);
                *reinterpret_cast<uint4 *>(xh + o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4 *>(xl + o) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            tmem_st_wait();
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&full[slot]);
            // ---- after handing over the last chunk of group g, promote group g-1 (its MMAs finished long ago)
            if ((i % TCG_G) == TCG_G - 1 || i == nch - 1) {
                const int g = i / TCG_G;
                if (g >= 1) drain(g - 1);
            }
        }
        cp_async_wait<0>();
        drain(ngroups - 1);
    } else {
        // ================================ MMA issuer (one elected thread of warp 4) ================================
        pdl_wait();
        {   // warp 8: the whole warp stays converged, one elected lane issues (operands stay in uniform registers)
            const uint32_t xh0 = smem_u32(smem + L::OFF_XH), xl0 = smem_u32(smem + L::OFF_XL);
            for (int i = 0; i < nch; ++i) {
                const int slot = i % TCG_NT, g = i / TCG_G, b = g & 1;
                const bool first = (i % TCG_G) == 0;
                if (first && g >= 2) {                 // D[b] must have been drained (group g-2)
                    mbar_wait(&dfree[b], ((g >> 1) - 1) & 1);
                }
                mbar_wait(&full[slot], (i / TCG_NT) & 1);
                tc_fence_after();
                const uint32_t d = tmem_base + b * BN;
                const uint32_t a_hi = tmem_base + A_COL0 + slot * 64, a_lo = a_hi + 32;
                const uint64_t bh = smem_desc_k128(xh0 + slot * L::X_TILE), bl = smem_desc_k128(xl0 + slot * L::X_TILE);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // small products first, the dominant hi*hi term last
                        mma_tf32_ts(d, a_lo + 8 * k, bh + 2 * k, IDESC, !(first && k == 0));
                        mma_tf32_ts(d, a_hi + 8 * k, bl + 2 * k, IDESC, 1);
                        mma_tf32_ts(d, a_hi + 8 * k, bh + 2 * k, IDESC, 1);
                    }
                    mma_commit(&done[slot]);
                    if ((i % TCG_G) == TCG_G - 1 || i == nch - 1) mma_commit(&dfull[b]);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();      // operand ring is idle from here on; TMEM is no longer needed
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<TM_COLS>(tmem_base);
    }

    // ================================ epilogue ================================
    const int n = n0 + row;
    if (splits == 1) {
        if (warp < 8) {
            const float bv = (bias && n < N) ? bias[n] : 0.f;
#pragma unroll
            for (int j = 0; j < HB; ++j) {
                const int m = m0 + half * HB + j;
                if (m < M && n < N) {
                    float r = acc[j] + bv;
                    if (act == 1) r = gelu_erf_tc(r);
                    if (residual) r += residual[(size_t)m * N + n];
                    y[(size_t)m * N + n] = r;
                }
            }
        }
    } else {
        cg::cluster_group cluster = cg::this_cluster();
        float *red = reinterpret_cast<float *>(smem);   // [BN][128] partial tile, reusing the operand ring
        if (warp < 8) {
#pragma unroll
            for (int j = 0; j < HB; ++j) red[(half * HB + j) * 128 + row] = acc[j];
        }
        cluster.sync();
        const int rank = (int)cluster.block_rank();
        // Each CTA finalises a contiguous slice of the tile.  All peers' values of an element group are fetched first
        // (independent DSMEM loads in flight, ~215 cycles each) and only then summed, in rank order.
        const float *peer[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) peer[s] = cluster.map_shared_rank(red, s < splits ? s : 0);
        constexpr int TOTAL4 = BN * 128 / 4;
        const int per4 = (TOTAL4 + splits - 1) / splits;
        const int e_beg = rank * per4, e_end = min(TOTAL4, e_beg + per4);
        for (int e4 = e_beg + tid; e4 < e_end; e4 += TCG_THREADS) {
            float4 v[16];
#pragma unroll
            for (int s = 0; s < 16; ++s)
                if (s < splits) v[s] = ld4(peer[s] + 4 * e4);
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int s = 0; s < 16; ++s)
                if (s < splits) { r.x += v[s].x; r.y += v[s].y; r.z += v[s].z; r.w += v[s].w; }
            const int m = m0 + (e4 >> 5), nn = n0 + ((e4 & 31) << 2);
            if (m < M) {
                float o[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (nn + q < N) {
                        float t = o[q];
                        if (bias) t += bias[nn + q];
                        if (act == 1) t = gelu_erf_tc(t);
                        const size_t off = (size_t)m * N + nn + q;
                        if (residual) t += residual[off];
                        y[off] = t;
                    }
                }
            }
        }
        cluster.sync();   // keep shared memory alive until every peer has read it
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Largest cluster size S <= 16 such that all `tiles` clusters of S CTAs can be co-resident (a cluster must fit in one GPC:
// with 16 the 8 GPCs of a B200 do not all have room, which would serialise the launch into two waves).
template <int BN>
static int tc_max_cluster(int tiles) {
    static int cache[17] = {0};   // cache[S] = max co-resident clusters of size S (0 = not queried yet)
    for (int S = 16; S >= 2; --S) {
        if (cache[S] == 0) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(1, S, 1);
            cfg.blockDim = dim3(TCG_THREADS);
            cfg.dynamicSmemBytes = TcgSmem<BN>::TOTAL;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = S; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, tc_linear_kernel<BN>, &cfg) != cudaSuccess) { cudaGetLastError(); n = -1; }
            cache[S] = n > 0 ? n : -1;
        }
        if (cache[S] >= tiles) return S;
    }
    return 1;
}

template <int BN>
static int tc_pick_splits(int M, int N, int K) {
    const int tiles = ((N + 127) / 128) * ((M + BN - 1) / BN);
    int s = 148 / tiles;
    if (s < 1) s = 1;
    const int nch = K / 32;
    if (s > nch / 2) s = nch / 2 > 0 ? nch / 2 : 1;   // at least two chunks per split
    if (s > 1) {
        const int cap = tc_max_cluster<BN>(tiles);
        if (s > cap) s = cap;
    }
    return s;
}

template <int BN>
static int launch_tc_t(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                       int act, cudaStream_t stream) {
    const int n_tiles = (N + 127) / 128, m_tiles = (M + BN - 1) / BN;
    if (m_tiles > 65535) return SFB200_E_ARG;
    static unsigned long long attr_done = 0;   // bit per device
    if (first_use_on_device(attr_done)) {
        SFB_CUDA_TRY(cudaFuncSetAttribute(tc_linear_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcgSmem<BN>::TOTAL));
        SFB_CUDA_TRY(cudaFuncSetAttribute(tc_linear_kernel<BN>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    const int splits = tc_pick_splits<BN>(M, N, K);
    return launch_ex("tc_linear", tc_linear_kernel<BN>, dim3(n_tiles, splits, m_tiles), dim3(TCG_THREADS), TcgSmem<BN>::TOTAL, stream,
                     dim3(1, splits, 1), x, W, bias, residual, y, M, N, K, act, splits);
}

int launch_linear_tc(const float *x, const float *W, const float *bias, const float *residual, float *y, int M, int N, int K,
                     int act, cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0 || K % 32 != 0) return SFB200_E_ARG;
    if (M <= 64) return launch_tc_t<64>(x, W, bias, residual, y, M, N, K, act, stream);
    return launch_tc_t<128>(x, W, bias, residual, y, M, N, K, act, stream);
}

}  // namespace sfb
