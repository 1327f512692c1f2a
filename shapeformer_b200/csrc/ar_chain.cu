// Decode-step GEMM chain: ONE persistent kernel runs a sequence of nn.Linear phases of a transformer block for the newest
// position of up to 64 rows —  proj(+residual) -> LN2 -> FC1(+GELU) -> FC2(+residual) -> LN1' -> QKV'  (or the head) —
// replacing one launch per GEMM / LayerNorm (Block.forward, transformer/mingpt.py:108-111; heads :222-231).
//
// One CTA per SM (grid = SM count, all co-resident), warp-specialised, decoupled through mbarriers:
//   warp 13 (loader)  : streams the fp32 weight tiles (128 output features x 32 k, 16 KB) of EVERY phase of the launch through
//                       a 10-deep shared-memory ring with TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B tensor maps over the
//                       row-major fp32 weights) — weights are constants, so the stream never waits for a phase boundary and
//                       runs through the grid barriers; the fp32 blob is read exactly once per step (no pre-split copy);
//   warps 0-3 (W split): thread = one weight row: conflict-free read of its 128-byte swizzled row, hi = rna_tf32(w),
//                       lo = rna_tf32(w - hi), tcgen05.st of hi|lo into a 4-deep TENSOR-MEMORY ring (the A operand lives in
//                       TMEM); may run up to 4 chunks ahead of the activations, i.e. into the next phase;
//   warps 4-11 (X)    : activation operand: coherent loads of x[64 x 32] from L2, optional LayerNorm applied on the fly from
//                       per-row statistics, TF32 hi/lo split into UMMA-layout shared tiles; promotion of the TMEM accumulator
//                       into fp32 registers every 2 chunks (the tensor-core accumulator is not round-to-nearest: chains stay
//                       at 24 MMAs); split-K partial tile -> L2 scratch; grid barrier; distributed reduction;
//   warp 12 (MMA)     : 12 tcgen05.mma.kind::tf32 per chunk (lo*hi + hi*lo + hi*hi), swap-AB: UMMA M = 128 output features,
//                       UMMA N = 64 activation rows.
// Split-K across CTAs (tiles x splits <= grid) is reduced THROUGH L2 in a fixed order (deterministic): every (tile, split)
// unit stores its 64 x 128 partial, one grid barrier, then each CTA finalises (row, 512-column) items: sum over the splits in
// split order + bias (+ exact-erf GELU) (+ residual) -> y, and — when a LayerNorm follows — the item's (mean, M2) so that the
// consumer merges row statistics with Chan's formula instead of re-reading the row.
#include <cuda.h>

#include "ar_kernels.cuh"
#include "tc_common.cuh"

namespace sfb {

using namespace tc;

constexpr int CH_THREADS = 448;            // warps 0-3 W split, 4-11 X (two per TMEM lane quadrant), 12 MMA issuer, 13 TMA loader
constexpr int CH_XT = 256;                 // X threads
constexpr int CH_WS = 10;                    // raw weight stages
constexpr int CH_XS = 3;                     // activation hi/lo stages
constexpr int CH_AS = 4;                     // TMEM A-operand stages
constexpr int CH_G = 2;                      // chunks per promotion group
constexpr int CH_BN = 64;                    // activation rows (UMMA N)
constexpr int CH_W_TILE = 128 * 32 * 4;      // 16 KB
constexpr int CH_X_TILE = CH_BN * 32 * 4;    // 8 KB
constexpr int CH_OFF_W = 0;
constexpr int CH_OFF_XH = CH_OFF_W + CH_WS * CH_W_TILE;
constexpr int CH_OFF_XL = CH_OFF_XH + CH_XS * CH_X_TILE;
constexpr int CH_N_MBAR = 2 * CH_WS + 2 * CH_AS + 2 * CH_XS + 4;    // mbarriers
constexpr int CH_OFF_BAR = CH_OFF_XL + CH_XS * CH_X_TILE;
constexpr int CH_OFF_STAT = CH_OFF_BAR + ((CH_N_MBAR * 8 + 8 + 127) / 128) * 128;   // (+ the TMEM address slot)                   // mean[64] rstd[64] red[16]
constexpr int CH_SMEM = CH_OFF_STAT + (64 + 64 + 16) * 4 + 64;
constexpr int CH_COL_D = 0, CH_COL_A = 2 * CH_BN;               // TMEM columns: D0 | D1 | A ring (AS x (hi 32 | lo 32))
constexpr int CH_ITEM_COLS = 512;                               // columns per reduction item = per LayerNorm statistics piece

static_assert(CH_SMEM <= 232448, "chain kernel shared memory exceeds the 227 KB limit");
static_assert(CH_COL_A + CH_AS * 64 <= 512, "TMEM columns");

__device__ __forceinline__ float gelu_erf_ch(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
                     "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_ch(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }

// ---- grid barrier among the X warps' leaders (the 256 X threads of every CTA take part; all other warps are decoupled by mbarriers)
__device__ __forceinline__ void chain_grid_barrier(unsigned int *ctr, unsigned int n_cta, int xt) {
    bar_sync(2, CH_XT);
    if (xt == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < n_cta);
        __threadfence();
    }
    bar_sync(2, CH_XT);
}

// sum over the first 128 X threads (4 warps), result broadcast; `red` = 8 floats of shared memory
__device__ __forceinline__ float x_block_sum(float v, float *red, int xt) {
    v = warp_sum(v);
    bar_sync(4, 128);                 // previous use of `red` is over
    if ((xt & 31) == 0) red[xt >> 5] = v;
    bar_sync(4, 128);
    return (red[0] + red[1]) + (red[2] + red[3]);
}

__global__ void __launch_bounds__(CH_THREADS, 1) ar_chain_kernel(const __grid_constant__ ChainArgs a) {
    pdl_trigger();
    extern __shared__ __align__(1024) unsigned char ch_smem[];
    unsigned char *smem = ch_smem;
    uint64_t *wfull = reinterpret_cast<uint64_t *>(smem + CH_OFF_BAR);   // [WS] raw weight tile landed (TMA tx bytes)
    uint64_t *wfree = wfull + CH_WS;                                     // [WS] raw tile consumed by the W-split warps
    uint64_t *afull = wfree + CH_WS;                                     // [AS] A operand (hi|lo) written to TMEM
    uint64_t *afree = afull + CH_AS;                                     // [AS] MMAs that read the A stage finished
    uint64_t *xfull = afree + CH_AS;                                     // [XS] activation hi/lo tiles written
    uint64_t *xfree = xfull + CH_XS;                                     // [XS] MMAs that read the x stage finished
    uint64_t *dfull = xfree + CH_XS;                                     // [2]  promotion group finished in D[b]
    uint64_t *dfree = dfull + 2;                                         // [2]  D[b] drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dfree + 2);
    float *s_mean = reinterpret_cast<float *>(smem + CH_OFF_STAT), *s_rstd = s_mean + 64, *s_red = s_rstd + 64;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, n_cta = gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < CH_WS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wfree[i], 128); }
        for (int i = 0; i < CH_AS; ++i) { mbar_init(&afull[i], 128); mbar_init(&afree[i], 1); }
        for (int i = 0; i < CH_XS; ++i) { mbar_init(&xfull[i], CH_XT); mbar_init(&xfree[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dfree[i], CH_XT); }
        mbar_fence_init();
    }
    if (warp == 12) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit of this CTA in phase p: (tile, split) and its K-chunk range; false when the CTA idles in that phase
    auto unit_of = [&](int p, int &tile, int &c_beg, int &c_end) -> bool {
        const ChainPhase &ph = a.ph[p];
        if (ph.tiles <= 0 || cta >= ph.tiles * ph.splits) return false;
        tile = cta / ph.splits;
        const int sp = cta % ph.splits, nch = ph.K >> 5;
        c_beg = (int)(((long long)sp * nch) / ph.splits);
        c_end = (int)(((long long)(sp + 1) * nch) / ph.splits);
        return c_end > c_beg;
    };

    if (warp == 13) {
        // ================================ weight loader (TMA), free-running across phases ================================
        if (lane == 0) {
            for (int p = 0; p < a.n_phases; ++p)
                if (a.ph[p].tiles > 0) tma_prefetch_desc(reinterpret_cast<const CUtensorMap *>(&a.wmap[p]));
            uint32_t it = 0;
            for (int p = 0; p < a.n_phases; ++p) {
                int tile, c_beg, c_end;
                if (!unit_of(p, tile, c_beg, c_end)) continue;
                for (int c = c_beg; c < c_end; ++c, ++it) {
                    const uint32_t s = it % CH_WS;
                    if (it >= CH_WS) mbar_wait(&wfree[s], ((it / CH_WS) - 1) & 1);
                    mbar_expect_tx_ch(&wfull[s], CH_W_TILE);
                    tma_load_2d(smem + CH_OFF_W + s * CH_W_TILE, reinterpret_cast<const CUtensorMap *>(&a.wmap[p]), c * 32, tile * 128, &wfull[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // ================================ W split: raw fp32 tile -> TF32 hi | lo in tensor memory ================================
        const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
        const int row = tid;   // weight row inside the tile == TMEM lane
        uint32_t it = 0;
        for (int p = 0; p < a.n_phases; ++p) {
            int tile, c_beg, c_end;
            if (!unit_of(p, tile, c_beg, c_end)) continue;
            for (int c = c_beg; c < c_end; ++c, ++it) {
                const uint32_t s = it % CH_WS, sa = it % CH_AS;
                mbar_wait(&wfull[s], (it / CH_WS) & 1);
                const float *wrow = reinterpret_cast<const float *>(smem + CH_OFF_W + s * CH_W_TILE) + row * 32;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const float4 v = ld4(wrow + ((ch ^ (row & 7)) << 2));
                    split_tf32(v.x, hi[4 * ch + 0], lo[4 * ch + 0]);
                    split_tf32(v.y, hi[4 * ch + 1], lo[4 * ch + 1]);
                    split_tf32(v.z, hi[4 * ch + 2], lo[4 * ch + 2]);
                    split_tf32(v.w, hi[4 * ch + 3], lo[4 * ch + 3]);
                }
                mbar_arrive(&wfree[s]);                      // the row is in registers: the raw stage may be refilled
                if (it >= CH_AS) mbar_wait(&afree[sa], ((it / CH_AS) - 1) & 1);
                tc_fence_after();
                const uint32_t a_col = tmem_base + lane_off + CH_COL_A + sa * 64;
                tmem_st32(a_col, hi);
                tmem_st32(a_col + 32, lo);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&afull[sa]);
            }
        }
    } else if (warp == 12) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t IDESC = instr_desc(2, 128, CH_BN);
            const uint32_t xh0 = smem_u32(smem + CH_OFF_XH), xl0 = smem_u32(smem + CH_OFF_XL);
            uint32_t it = 0, gg = 0;
            for (int p = 0; p < a.n_phases; ++p) {
                int tile, c_beg, c_end;
                if (!unit_of(p, tile, c_beg, c_end)) continue;
                const int nch = c_end - c_beg;
                for (int i = 0; i < nch; ++i, ++it) {
                    const uint32_t sa = it % CH_AS, sx = it % CH_XS, b = gg & 1;
                    const bool first = (i % CH_G) == 0;
                    if (first && gg >= 2) mbar_wait(&dfree[b], ((gg >> 1) - 1) & 1);
                    mbar_wait(&afull[sa], (it / CH_AS) & 1);
                    mbar_wait(&xfull[sx], (it / CH_XS) & 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + CH_COL_D + b * CH_BN;
                    const uint32_t a_hi = tmem_base + CH_COL_A + sa * 64, a_lo = a_hi + 32;
                    const uint64_t bh = smem_desc_k128(xh0 + sx * CH_X_TILE), bl = smem_desc_k128(xl0 + sx * CH_X_TILE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        mma_tf32_ts(d, a_lo + 8 * k, bh + 2 * k, IDESC, !(first && k == 0));
                        mma_tf32_ts(d, a_hi + 8 * k, bl + 2 * k, IDESC, 1);
                        mma_tf32_ts(d, a_hi + 8 * k, bh + 2 * k, IDESC, 1);
                    }
                    mma_commit(&afree[sa]);
                    mma_commit(&xfree[sx]);
                    if ((i % CH_G) == CH_G - 1 || i == nch - 1) { mma_commit(&dfull[b]); ++gg; }
                }
            }
        }
        __syncwarp();
    } else {
        // ================================ X warps: activations, promotion, partials, barriers, reduction ================================
        const int xt = tid - 128;                                   // 0..255
        const int row = xt & 127, half = xt >> 7;                   // TMEM lane (weight row of the tile) / which 32 D columns
        const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
        constexpr int HB = CH_BN / 2;
        const int M = a.M;
        pdl_wait();                                                 // everything below touches data of earlier kernels
        uint32_t it = 0, gg = 0;
        int bar_i = 0;
        for (int p = 0; p < a.n_phases; ++p) {
            const ChainPhase &ph = a.ph[p];
            if (ph.wait_before) { chain_grid_barrier(a.bar + bar_i, n_cta, xt); ++bar_i; }
            int tile, c_beg, c_end;
            const bool has = unit_of(p, tile, c_beg, c_end);
            if (has) {
                const int K = ph.K;
                // ---- LayerNorm row statistics: merge the (mean, M2) pieces of each row (Chan et al.)
                if (ph.ln_g) {
                    if (xt < CH_BN) {
                        float mean = 0.f, rstd = 0.f;
                        if (xt < M) {
                            float n = 0.f, m2 = 0.f;
                            const int pieces = (K + CH_ITEM_COLS - 1) / CH_ITEM_COLS;
                            for (int q = 0; q < pieces; ++q) {
                                const float2 st = __ldcg(reinterpret_cast<const float2 *>(ph.stats_in) + (size_t)xt * pieces + q);
                                const float nb = (float)min(CH_ITEM_COLS, K - q * CH_ITEM_COLS);
                                const float delta = st.x - mean, tot = n + nb;
                                mean += delta * (nb / tot);
                                m2 += st.y + delta * delta * (n * nb / tot);
                                n = tot;
                            }
                            rstd = 1.0f / sqrtf(m2 / (float)K + 1e-5f);
                        }
                        s_mean[xt] = mean; s_rstd[xt] = rstd;
                    }
                    bar_sync(3, CH_XT);
                }
                float acc[HB];             // acc[j] = D[row][half * 32 + j]
#pragma unroll
                for (int j = 0; j < HB; ++j) acc[j] = 0.f;
                auto drain = [&]() {          // acc += D[gg & 1]; advances gg
                    const uint32_t b = gg & 1;
                    mbar_wait(&dfull[b], (gg >> 1) & 1);
                    tc_fence_after();
                    {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + lane_off + CH_COL_D + b * CH_BN + half * HB, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
                    }
                    tc_fence_before();
                    mbar_arrive(&dfree[b]);
                    ++gg;
                };
                // thread -> 2 float4 of the 64 x 32 chunk: element idx = xt + 256 j: row = idx >> 3, 16-byte chunk = idx & 7
                auto load_x = [&](int c, float4 (&v)[2]) {
                    const int k0 = c * 32;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int idx = xt + CH_XT * j, r = idx >> 3, chk = idx & 7;
                        v[j] = r < M ? ldcg4(ph.x + (size_t)r * K + k0 + chk * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                };
                const int nch = c_end - c_beg;
                float4 cur[2], nxt[2];
                load_x(c_beg, cur);
                int pending = 0;              // promotion groups committed by the MMA warp but not drained yet
                for (int i = 0; i < nch; ++i, ++it) {
                    if (i + 1 < nch) load_x(c_beg + i + 1, nxt);
                    const uint32_t sx = it % CH_XS;
                    if (it >= CH_XS) mbar_wait(&xfree[sx], ((it / CH_XS) - 1) & 1);
                    float *xh = reinterpret_cast<float *>(smem + CH_OFF_XH + sx * CH_X_TILE);
                    float *xl = reinterpret_cast<float *>(smem + CH_OFF_XL + sx * CH_X_TILE);
                    const int k0 = (c_beg + i) * 32;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int idx = xt + CH_XT * j, r = idx >> 3, chk = idx & 7;
                        float4 v = cur[j];
                        if (ph.ln_g && r < M) {
                            const float mean = s_mean[r], rstd = s_rstd[r];
                            const float4 g = __ldg(reinterpret_cast<const float4 *>(ph.ln_g + k0 + chk * 4));
                            const float4 bb = __ldg(reinterpret_cast<const float4 *>(ph.ln_b + k0 + chk * 4));
                            v.x = (v.x - mean) * rstd * g.x + bb.x;
                            v.y = (v.y - mean) * rstd * g.y + bb.y;
                            v.z = (v.z - mean) * rstd * g.z + bb.z;
                            v.w = (v.w - mean) * rstd * g.w + bb.w;
                        }
                        const int o = r * 32 + ((chk ^ (r & 7)) << 2);
                        uint32_t h[4], l[4];
                        split_tf32(v.x, h[0], l[0]); split_tf32(v.y, h[1], l[1]);
                        split_tf32(v.z, h[2], l[2]); split_tf32(v.w, h[3], l[3]);
                        *reinterpret_cast<uint4 *>(xh + o) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4 *>(xl + o) = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&xfull[sx]);
                    if ((i % CH_G) == CH_G - 1 || i == nch - 1) {
                        // group g of this unit handed over: promote the previous one (its MMAs finished long ago)
                        if (pending > 0) { drain(); --pending; }
                        ++pending;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) cur[j] = nxt[j];
                }
                while (pending > 0) { drain(); --pending; }
                // ---- partial tile of this (tile, split) unit -> L2 scratch [unit][m][128]
                float *part = a.scratch + ((size_t)cta * CH_BN + half * HB) * 128 + row;
#pragma unroll
                for (int j = 0; j < HB; ++j)
                    if (half * HB + j < M) __stcg(part + j * 128, acc[j]);
            }
            if (ph.tiles > 0) { chain_grid_barrier(a.bar + bar_i, n_cta, xt); ++bar_i; }
            // ---- distributed reduction: items (row m, 512-column block cb), by the first 128 X threads
            if (xt < 128) {
                const int N = ph.N, S = ph.tiles > 0 ? ph.splits : 0;
                const int CB = (N + CH_ITEM_COLS - 1) / CH_ITEM_COLS;
                const bool vec = (N & 3) == 0;
                for (int item = cta; item < M * CB; item += n_cta) {
                    const int m = item / CB, cb = item % CB;
                    const int n = cb * CH_ITEM_COLS + xt * 4;
                    const int cnt = min(CH_ITEM_COLS, N - cb * CH_ITEM_COLS);
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    const int nv = max(0, min(4, N - n));      // valid columns of this thread
                    if (nv > 0) {
                        if (S > 0) {
                            const int t = n >> 7, nl = n & 127;
                            const float *src = a.scratch + (((size_t)t * S) * CH_BN + m) * 128 + nl;
                            // two batches of <= 9 independent loads in flight; summed in split order (deterministic)
#pragma unroll
                            for (int s0 = 0; s0 < 18; s0 += 9) {
                                float4 pv[9];
#pragma unroll
                                for (int s = 0; s < 9; ++s)
                                    if (s0 + s < S) pv[s] = ldcg4(src + (size_t)(s0 + s) * CH_BN * 128);
#pragma unroll
                                for (int s = 0; s < 9; ++s)
                                    if (s0 + s < S) { v[0] += pv[s].x; v[1] += pv[s].y; v[2] += pv[s].z; v[3] += pv[s].w; }
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q < nv) {
                                float t = v[q];
                                if (ph.bias) t += __ldg(ph.bias + n + q);
                                if (ph.act == 1) t = gelu_erf_ch(t);
                                if (ph.residual) t += __ldcg(ph.residual + (size_t)m * N + n + q);
                                v[q] = t;
                            }
                        }
                        if (ph.y) {
                            float *dst = ph.y + (size_t)m * N + n;
                            if (vec) __stcg(reinterpret_cast<float4 *>(dst), make_float4(v[0], v[1], v[2], v[3]));
                            else
                                for (int q = 0; q < nv; ++q) __stcg(dst + q, v[q]);
                        }
                    }
                    if (ph.stats_out) {
                        float s = 0.f;
                        for (int q = 0; q < nv; ++q) s += v[q];
                        const float mean = x_block_sum(s, s_red, xt) / (float)cnt;
                        float d2 = 0.f;
                        for (int q = 0; q < nv; ++q) d2 += (v[q] - mean) * (v[q] - mean);
                        const float m2 = x_block_sum(d2, s_red, xt);
                        if (xt == 0) __stcg(reinterpret_cast<float2 *>(ph.stats_out) + (size_t)m * CB + cb, make_float2(mean, m2));
                    }
                }
            }
        }
        // ---- reset the barrier counters for the next launch: the last CTA to get here knows everybody passed every barrier
        bar_sync(2, CH_XT);
        if (xt == 0) {
            __threadfence();
            const unsigned int old = atomicAdd(a.bar + CH_MAX_BARRIERS, 1u);
            if (old == (unsigned int)n_cta - 1) {
                for (int i = 0; i <= CH_MAX_BARRIERS; ++i) a.bar[i] = 0u;
                __threadfence();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// Tensor map over a row-major fp32 weight matrix W (N, K): box = 32 k x 128 rows, SWIZZLE_128B (the UMMA K-major layout),
// rows past N read as zeros.
int chain_weight_map(const float *W, int N, int K, void *map_out) {
    if (!W || N <= 0 || K <= 0 || K % 32 != 0 || (reinterpret_cast<uintptr_t>(W) & 15)) return SFB200_E_ARG;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_cuda_error(cudaErrorNotSupported, "cuTensorMapEncodeTiled entry point"); return SFB200_E_CUDA; }
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {32, 128};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(static_cast<CUtensorMap *>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(W), gdim,
                           gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled"); return SFB200_E_CUDA; }
    return SFB200_OK;
}

int chain_grid_size() {
    static int n_sm[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return 0;
    if (n_sm[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
        n_sm[dev] = n;
    }
    return n_sm[dev];
}

// tiles / splits of a GEMM phase for a grid of `grid` CTAs: every (tile, split) unit gets its own CTA
void chain_plan(int N, int K, int grid, int *tiles, int *splits) {
    const int t = (N + 127) / 128, nch = K / 32;
    int s = grid / t;
    if (s < 1) s = 1;
    if (s > nch) s = nch;
    if (s > 18) s = 18;
    *tiles = t; *splits = s;
}

size_t chain_scratch_floats(int grid) { return (size_t)grid * CH_BN * 128; }

int launch_chain(const ChainArgs &args, cudaStream_t stream) {
    const int grid = chain_grid_size();
    if (grid <= 0 || args.n_phases < 1 || args.n_phases > CH_MAXP || args.M < 1 || args.M > CH_BN) return SFB200_E_ARG;
    int nbar = 0;
    for (int p = 0; p < args.n_phases; ++p) {
        const ChainPhase &ph = args.ph[p];
        if (ph.tiles > 0 && (ph.tiles * ph.splits > grid || ph.splits > 18 || ph.K % 32 != 0)) return SFB200_E_ARG;
        nbar += (ph.wait_before ? 1 : 0) + (ph.tiles > 0 ? 1 : 0);
    }
    if (nbar > CH_MAX_BARRIERS) return SFB200_E_ARG;
    static unsigned long long attr_done = 0;
    if (first_use_on_device(attr_done)) {
        SFB_CUDA_TRY(cudaFuncSetAttribute(ar_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
        int per_sm = 0;
        SFB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_chain_kernel, CH_THREADS, CH_SMEM));
        if (per_sm < 1) { set_cuda_error(cudaErrorLaunchOutOfResources, "ar_chain_kernel does not fit an SM"); return SFB200_E_CUDA; }
    }
    return launch_ex("ar_chain", ar_chain_kernel, dim3(grid), dim3(CH_THREADS), CH_SMEM, stream, dim3(1, 1, 1), args);
}

}  // namespace sfb
