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
constexpr int CH_XT = 256;                   // X threads
constexpr int CH_WS = 6;                     // raw weight stages
constexpr int CH_XR = 8;                     // activation prefetch depth (chunks), thread-private staging in shared memory
constexpr int CH_XRL = CH_XR / 2;            // ... per X group (the two groups convert alternate chunks)
constexpr int CH_LNP = 2;                    // LayerNorm-fed GEMM phases per launch whose gamma / beta slices are staged
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
constexpr int CH_OFF_STAT = CH_OFF_BAR + ((CH_N_MBAR * 8 + 8 + 127) / 128) * 128;   // (+ the TMEM address slot)   mean[64] rstd[64] red[16]
constexpr int CH_OFF_XRAW = CH_OFF_STAT + 640;                  // [XR][2][256 threads] float4: per-thread activation prefetch
constexpr int CH_OFF_LN = CH_OFF_XRAW + CH_XR * 2 * CH_XT * 16; // [LNP][gamma | beta][XR chunks][32] floats
constexpr int CH_OFF_STST = CH_OFF_LN + CH_LNP * 2 * CH_XR * 32 * 4;   // [64 rows][8 pieces] float2: statistics staging
constexpr int CH_SMEM = CH_OFF_STST + 64 * 8 * 8;
constexpr int CH_COL_D = 0, CH_COL_A = 2 * CH_BN;               // TMEM columns: D0 | D1 | A ring (AS x (hi 32 | lo 32))
constexpr int CH_ITEM_COLS = 512;                               // columns per reduction item (128 threads x float4)
constexpr int CH_STAT_COLS = 128;                               // columns per LayerNorm statistics piece (one warp)

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

// Optional timeline probe (development aid): when a buffer is registered (sfb200_debug_chain_timeline), thread xt == 0 of
// CTA 0 adds the SM cycles (clock64) since the previous stamp to slot [phase * 8 + stage] and counts the visits:
// stages 0 wait_before barrier, 1 row statistics, 2 operand prefetch issued, 3 chunk loop + final promotion, 4 partial store,
// 5 post-GEMM barrier, 6 reduction; slot 62 = kernel entry -> dependency wait done, 63 = launches.
__device__ unsigned long long *g_chain_timeline = nullptr;
#ifndef SFB_CHAIN_PROBE
#define SFB_CHAIN_PROBE 0      // build with -DSFB_CHAIN_PROBE=1 for the timeline (costs registers and ~10 % of the kernel time)
#endif
#if SFB_CHAIN_PROBE
struct ChainProbe {     // accumulates in (thread-local) memory, flushed to the global buffer once at the end
    unsigned long long t;
    unsigned int acc[64], cnt[64];
    bool on;
    __device__ __forceinline__ void start(bool enable) {
        on = enable && g_chain_timeline != nullptr;
        if (on) {
            for (int i = 0; i < 64; ++i) { acc[i] = 0; cnt[i] = 0; }
            t = clock64();
        }
    }
    __device__ __forceinline__ void stamp(int slot) {
        if (on) {
            const unsigned long long now = clock64();      // SM cycles (the probing thread stays on one SM)
            acc[slot] += (unsigned int)(now - t);
            cnt[slot] += 1;
            t = now;
        }
    }
    __device__ __forceinline__ void flush() {
        if (on)
            for (int i = 0; i < 64; ++i)
                if (cnt[i]) {
                    atomicAdd(&g_chain_timeline[i], (unsigned long long)acc[i]);
                    atomicAdd(&g_chain_timeline[64 + i], (unsigned long long)cnt[i]);
                }
    }
};
#else
struct ChainProbe {
    __device__ __forceinline__ void start(bool) {}
    __device__ __forceinline__ void stamp(int) {}
    __device__ __forceinline__ void flush() {}
};
#endif

// ---- grid barrier among the X warps' leaders (the 256 X threads of every CTA take part; all other warps are decoupled by mbarriers)
__device__ __forceinline__ void chain_grid_barrier(unsigned int *ctr, unsigned int n_cta, int xt) {
    bar_sync(2, CH_XT);          // every X thread's global writes happen-before the leader's release (cumulativity)
    if (xt == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < n_cta);
    }
    bar_sync(2, CH_XT);
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
    float *s_mean = reinterpret_cast<float *>(smem + CH_OFF_STAT), *s_rstd = s_mean + 64;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, n_cta = gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < CH_WS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wfree[i], 4); }
        for (int i = 0; i < CH_AS; ++i) { mbar_init(&afull[i], 4); mbar_init(&afree[i], 1); }
        for (int i = 0; i < CH_XS; ++i) { mbar_init(&xfull[i], CH_XT / 64); mbar_init(&xfree[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dfree[i], CH_XT / 32); }
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
        ChainProbe wp;
        wp.start(cta == 0 && tid == 0);
        for (int p = 0; p < a.n_phases; ++p) {
            int tile, c_beg, c_end;
            if (!unit_of(p, tile, c_beg, c_end)) continue;
            for (int c = c_beg; c < c_end; ++c, ++it) {
                const uint32_t s = it % CH_WS, sa = it % CH_AS;
                wp.stamp(43);
                mbar_wait(&wfull[s], (it / CH_WS) & 1);
                wp.stamp(40);
                const float *wrow = reinterpret_cast<const float *>(smem + CH_OFF_W + s * CH_W_TILE) + row * 32;
                // hi = the raw fp32 bits: the tensor core reads only the upper 19 bits of a tf32 operand, i.e. it truncates
                // (hi_t = trunc_tf32(w), no instruction needed); lo = rna_tf32(w - hi_t), the subtraction being exact
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const float4 v = ld4(wrow + ((ch ^ (row & 7)) << 2));
                    hi[4 * ch + 0] = __float_as_uint(v.x); hi[4 * ch + 1] = __float_as_uint(v.y);
                    hi[4 * ch + 2] = __float_as_uint(v.z); hi[4 * ch + 3] = __float_as_uint(v.w);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&wfree[s]);       // the rows are in registers: the raw stage may be refilled
                wp.stamp(41);
                if (it >= CH_AS) mbar_wait(&afree[sa], ((it / CH_AS) - 1) & 1);
                wp.stamp(42);
                tc_fence_after();
                const uint32_t a_col = tmem_base + lane_off + CH_COL_A + sa * 64;
                tmem_st32(a_col, hi);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float w = __uint_as_float(hi[i]);
                    lo[i] = __float_as_uint(w - __uint_as_float(hi[i] & 0xFFFFE000u)) + 0x1000u;   // low 13 bits ignored by the MMA
                }
                tmem_st32(a_col + 32, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&afull[sa]);      // one arrival per warp: 4 instead of 128 shared-memory atomics
            }
        }
        wp.flush();
    } else if (warp == 12) {
        // ================================ MMA issuer (whole warp converged; one elected lane issues) ================================
        {
            constexpr uint32_t IDESC = instr_desc(2, 128, CH_BN);
            const uint32_t xh0 = smem_u32(smem + CH_OFF_XH), xl0 = smem_u32(smem + CH_OFF_XL);
            uint32_t it = 0, gg = 0;
            ChainProbe mp;
            mp.start(cta == 0 && lane == 0);
            for (int p = 0; p < a.n_phases; ++p) {
                int tile, c_beg, c_end;
                if (!unit_of(p, tile, c_beg, c_end)) continue;
                const int nch = c_end - c_beg;
                for (int i = 0; i < nch; ++i, ++it) {
                    const uint32_t sa = it % CH_AS, sx = it % CH_XS, b = gg & 1;
                    const bool first = (i % CH_G) == 0;
                    const bool last = (i % CH_G) == CH_G - 1 || i == nch - 1;
                    mp.stamp(47);
                    if (first && gg >= 2) mbar_wait(&dfree[b], ((gg >> 1) - 1) & 1);
                    mp.stamp(44);
                    mbar_wait(&xfull[sx], (it / CH_XS) & 1);
                    mp.stamp(46);
                    mbar_wait(&afull[sa], (it / CH_AS) & 1);
                    mp.stamp(45);
                    tc_fence_after();
                    const uint32_t d = tmem_base + CH_COL_D + b * CH_BN;
                    const uint32_t a_hi = tmem_base + CH_COL_A + sa * 64, a_lo = a_hi + 32;
                    const uint64_t bh = smem_desc_k128(xh0 + sx * CH_X_TILE), bl = smem_desc_k128(xl0 + sx * CH_X_TILE);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            mma_tf32_ts(d, a_lo + 8 * k, bh + 2 * k, IDESC, !(first && k == 0));
                            mma_tf32_ts(d, a_hi + 8 * k, bl + 2 * k, IDESC, 1);
                            mma_tf32_ts(d, a_hi + 8 * k, bh + 2 * k, IDESC, 1);
                        }
                        mma_commit(&afree[sa]);
                        mma_commit(&xfree[sx]);
                        if (last) mma_commit(&dfull[b]);
                    }
                    __syncwarp();
                    if (last) ++gg;
                }
            }
            mp.flush();
        }
    } else {
        // ================================ X warps: activations, promotion, partials, barriers, reduction ================================
        const int xt = tid - 128;                                   // 0..255
        const int row = xt & 127, half = xt >> 7;                   // TMEM lane (weight row of the tile) / which 32 D columns
        const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
        constexpr int HB = CH_BN / 2;
        const int M = a.M;
        ChainProbe probe;
        probe.start(cta == 0 && xt == 0);
        float4 *xraw = reinterpret_cast<float4 *>(smem + CH_OFF_XRAW);
        float *s_ln = reinterpret_cast<float *>(smem + CH_OFF_LN);
        // LayerNorm gamma / beta slices of this CTA's K ranges are weights: staged before the dependency wait
        {
            int lnp = 0;
            for (int p = 0; p < a.n_phases; ++p) {
                const ChainPhase &ph = a.ph[p];
                if (!ph.ln_g) continue;
                int tile, c_beg, c_end;
                if (lnp < CH_LNP && unit_of(p, tile, c_beg, c_end) && c_end - c_beg <= CH_XR) {
                    const int n = (c_end - c_beg) * 32;
                    float *dst = s_ln + lnp * 2 * CH_XR * 32;
                    for (int i = xt; i < 2 * n; i += CH_XT)
                        dst[(i < n ? 0 : CH_XR * 32) + (i < n ? i : i - n)] = __ldg((i < n ? ph.ln_g : ph.ln_b) + c_beg * 32 + (i < n ? i : i - n));
                }
                ++lnp;
            }
        }
        pdl_wait();                                                 // everything below touches data of earlier kernels
        probe.stamp(62);
        uint32_t it = 0, gg = 0;
        int bar_i = 0, lnp = 0;
        for (int p = 0; p < a.n_phases; ++p) {
            const ChainPhase &ph = a.ph[p];
            if (ph.wait_before) { chain_grid_barrier(a.bar + bar_i, n_cta, xt); ++bar_i; }
            probe.stamp(p * 8 + 0);
            int tile, c_beg, c_end;
            const bool has = unit_of(p, tile, c_beg, c_end);
            const int my_lnp = ph.ln_g ? lnp++ : -1;
            if (has) {
                const int K = ph.K;
                const int nch = c_end - c_beg;
                const int grp = xt >> 7, gt = xt & 127;     // the two X groups (warps 4-7 / 8-11) produce alternate chunks
                // ---- prefetch of the activation chunks this thread will convert (chunks i with (i & 1) == grp): cp.async into
                //      thread-private staging (the copying thread is the reading thread: no barrier), 4 own chunks in flight; one
                //      commit group per slot, empty groups keep the group arithmetic uniform
                auto issue_x = [&](int i) {
                    const int k0 = (c_beg + i) * 32, slot = (i >> 1) % CH_XRL;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int idx = gt + 128 * j, r = idx >> 3, chk = idx & 7;
                        cp_async16(xraw + (slot * 4 + j) * CH_XT + xt, ph.x + (size_t)(r < M ? r : 0) * K + k0 + chk * 4, r < M ? 16 : 0);
                    }
                };
                // ---- LayerNorm row statistics (loads first, so that they do not queue behind the operand prefetch burst):
                //      thread xt < M fetches the (mean, M2) pieces of row xt — 8 pieces = 64 bytes through cp.async into its private
                //      slot when K <= 1024, plain loads otherwise — and merges them with Chan's formula
                const int pieces = ph.ln_g ? (K + CH_STAT_COLS - 1) / CH_STAT_COLS : 0;
                const bool st_async = pieces > 0 && pieces <= 8 && (pieces & 1) == 0;   // 16-byte aligned rows
                float2 *my_st = reinterpret_cast<float2 *>(smem + CH_OFF_STST) + (xt & 63) * 8;
                if (st_async && xt < M) {
                    const float2 *sp = reinterpret_cast<const float2 *>(ph.stats_in) + (size_t)xt * pieces;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (2 * q < pieces) cp_async16(my_st + 2 * q, sp + 2 * q, 2 * q + 1 < pieces ? 16 : 8);
                }
                cp_async_commit();                           // group "statistics" (possibly empty)
#pragma unroll
                for (int li = 0; li < CH_XRL; ++li) {
                    if (2 * li + grp < nch) issue_x(2 * li + grp);
                    cp_async_commit();
                }
                if (ph.ln_g) {
                    if (xt < CH_BN) {
                        float mean = 0.f, rstd = 0.f;
                        if (xt < M) {
                            float n = 0.f, m2 = 0.f;
                            const float2 *sp = reinterpret_cast<const float2 *>(ph.stats_in) + (size_t)xt * pieces;
                            if (st_async) cp_async_wait<CH_XRL>();      // the statistics group has landed
                            for (int q0 = 0; q0 < pieces; q0 += 8) {
                                float2 stp[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q)      // 8 pieces in flight at once
                                    if (q0 + q < pieces) stp[q] = st_async ? my_st[q] : __ldcg(sp + q0 + q);
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    if (q0 + q < pieces) {
                                        const float nb = (float)min(CH_STAT_COLS, K - (q0 + q) * CH_STAT_COLS);
                                        const float delta = stp[q].x - mean, tot = n + nb;
                                        mean += delta * (nb / tot);
                                        m2 += stp[q].y + delta * delta * (n * nb / tot);
                                        n = tot;
                                    }
                                }
                            }
                            rstd = 1.0f / sqrtf(m2 / (float)K + 1e-5f);
                        }
                        s_mean[xt] = mean; s_rstd[xt] = rstd;
                    }
                    bar_sync(3, CH_XT);
                }
                probe.stamp(p * 8 + 1);
                constexpr int HBc = CH_BN / 2;
                float acc[HBc];            // acc[j] = D[row][half * 32 + j]
#pragma unroll
                for (int j = 0; j < HBc; ++j) acc[j] = 0.f;
                auto drain = [&]() {          // acc += D[gg & 1]; advances gg
                    const uint32_t b = gg & 1;
                    mbar_wait(&dfull[b], (gg >> 1) & 1);
                    probe.stamp(51);
                    tc_fence_after();
                    {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + lane_off + CH_COL_D + b * CH_BN + half * HB, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&dfree[b]);
                    ++gg;
                };
                const bool ln_staged = my_lnp >= 0 && my_lnp < CH_LNP && nch <= CH_XR;
                const float *sg = s_ln + (my_lnp < 0 ? 0 : my_lnp) * 2 * CH_XR * 32, *sb = sg + CH_XR * 32;
                probe.stamp(p * 8 + 2);
                int pending = 0;              // promotion groups committed by the MMA warp but not drained yet
                for (int i = 0; i < nch; ++i, ++it) {
                    if ((i & 1) == grp) {
                        // ---- this group's chunk: staged fp32 -> (LayerNorm) -> TF32 hi / lo UMMA tiles
                        probe.stamp(48);
                        cp_async_wait<CH_XRL - 1>();       // this thread's copies of chunk i have landed
                        const int slot = (i >> 1) % CH_XRL;
                        float4 cur[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) cur[j] = xraw[(slot * 4 + j) * CH_XT + xt];
                        if (i + 2 * CH_XRL < nch) issue_x(i + 2 * CH_XRL);
                        cp_async_commit();
                        const uint32_t sx = it % CH_XS;
                        if (it >= CH_XS) mbar_wait(&xfree[sx], ((it / CH_XS) - 1) & 1);
                        probe.stamp(49);
                        float *xh = reinterpret_cast<float *>(smem + CH_OFF_XH + sx * CH_X_TILE);
                        float *xl = reinterpret_cast<float *>(smem + CH_OFF_XL + sx * CH_X_TILE);
                        const int k0 = (c_beg + i) * 32;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int idx = gt + 128 * j, r = idx >> 3, chk = idx & 7;
                            float4 v = cur[j];
                            if (ph.ln_g && r < M) {
                                const float mean = s_mean[r], rstd = s_rstd[r];
                                const float4 g = ln_staged ? ld4(sg + i * 32 + chk * 4) : __ldg(reinterpret_cast<const float4 *>(ph.ln_g + k0 + chk * 4));
                                const float4 bb = ln_staged ? ld4(sb + i * 32 + chk * 4) : __ldg(reinterpret_cast<const float4 *>(ph.ln_b + k0 + chk * 4));
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
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&xfull[sx]);
                        probe.stamp(50);
                    }
                    if ((i % CH_G) == CH_G - 1 || i == nch - 1) {
                        // promotion group g of this unit is in the MMA warp's hands: promote the previous one
                        if (pending > 0) { drain(); --pending; }
                        ++pending;
                    }
                }
                cp_async_wait<0>();
                while (pending > 0) { drain(); --pending; }
                probe.stamp(p * 8 + 3);
                // ---- partial tile of this (tile, split) unit -> L2 scratch [unit][m][128]
                float *part = a.scratch + ((size_t)cta * CH_BN + half * HB) * 128 + row;
#pragma unroll
                for (int j = 0; j < HB; ++j)
                    if (half * HB + j < M) __stcg(part + j * 128, acc[j]);
                probe.stamp(p * 8 + 4);
            }
            if (ph.tiles > 0) { chain_grid_barrier(a.bar + bar_i, n_cta, xt); ++bar_i; }
            probe.stamp(p * 8 + 5);
            // ---- distributed reduction: items (row m, 512-column block cb); the two groups of 128 X threads take alternate
            //      items; every load of an item is issued before the first use; LayerNorm statistics per warp (128 columns)
            {
                const int grp = xt >> 7, gt = xt & 127, wq = gt >> 5;
                const int N = ph.N, S = ph.tiles > 0 ? ph.splits : 0;
                const int CB = (N + CH_ITEM_COLS - 1) / CH_ITEM_COLS, PC = (N + CH_STAT_COLS - 1) / CH_STAT_COLS;
                const bool vec = (N & 3) == 0;
                for (int item = cta + grp * n_cta; item < M * CB; item += 2 * n_cta) {
                    const int m = item / CB, cb = item % CB;
                    const int n = cb * CH_ITEM_COLS + gt * 4;
                    const int nv = max(0, min(4, N - n));      // valid columns of this thread
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (nv > 0) {
                        const int t = n >> 7, nl = n & 127;
                        const float *src = a.scratch + (((size_t)t * S) * CH_BN + m) * 128 + nl;
                        float4 pv[9];
                        float bs[4] = {0.f, 0.f, 0.f, 0.f}, rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int s = 0; s < 9; ++s)
                            if (s < S) pv[s] = ldcg4(src + (size_t)s * CH_BN * 128);
                        if (vec) {
                            if (ph.bias) { const float4 b4 = __ldg(reinterpret_cast<const float4 *>(ph.bias + n)); bs[0] = b4.x; bs[1] = b4.y; bs[2] = b4.z; bs[3] = b4.w; }
                            if (ph.residual) { const float4 r4 = ldcg4(ph.residual + (size_t)m * N + n); rs[0] = r4.x; rs[1] = r4.y; rs[2] = r4.z; rs[3] = r4.w; }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (q < nv) {
                                    if (ph.bias) bs[q] = __ldg(ph.bias + n + q);
                                    if (ph.residual) rs[q] = __ldcg(ph.residual + (size_t)m * N + n + q);
                                }
                            }
                        }
#pragma unroll
                        for (int s = 0; s < 9; ++s)
                            if (s < S) { v[0] += pv[s].x; v[1] += pv[s].y; v[2] += pv[s].z; v[3] += pv[s].w; }
                        if (S > 9) {                             // splits 9..17, still in split order (deterministic)
#pragma unroll
                            for (int s = 0; s < 9; ++s)
                                if (9 + s < S) pv[s] = ldcg4(src + (size_t)(9 + s) * CH_BN * 128);
#pragma unroll
                            for (int s = 0; s < 9; ++s)
                                if (9 + s < S) { v[0] += pv[s].x; v[1] += pv[s].y; v[2] += pv[s].z; v[3] += pv[s].w; }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float tq = v[q];
                            if (ph.bias) tq += bs[q];
                            if (ph.act == 1) tq = gelu_erf_ch(tq);
                            if (ph.residual) tq += rs[q];
                            v[q] = q < nv ? tq : 0.f;
                        }
                        if (ph.y) {
                            float *dst = ph.y + (size_t)m * N + n;
                            if (vec) __stcg(reinterpret_cast<float4 *>(dst), make_float4(v[0], v[1], v[2], v[3]));
                            else
                                for (int q = 0; q < nv; ++q) __stcg(dst + q, v[q]);
                        }
                    }
                    if (ph.stats_out) {
                        // (mean, M2) of this warp's 128 columns of row m — merged by the consumer (Chan et al.)
                        const int piece = cb * (CH_ITEM_COLS / CH_STAT_COLS) + wq;
                        const int cnt = max(0, min(CH_STAT_COLS, N - piece * CH_STAT_COLS));
                        const float mean = warp_sum((v[0] + v[1]) + (v[2] + v[3])) / (float)max(cnt, 1);
                        float d2 = 0.f;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q < nv) d2 += (v[q] - mean) * (v[q] - mean);
                        const float m2 = warp_sum(d2);
                        if (lane == 0 && cnt > 0) __stcg(reinterpret_cast<float2 *>(ph.stats_out) + (size_t)m * PC + piece, make_float2(mean, m2));
                    }
                }
            }
            probe.stamp(p * 8 + 6);
        }
        probe.flush();
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

int set_chain_timeline(unsigned long long *buf) {
    SFB_CUDA_TRY(cudaMemcpyToSymbol(g_chain_timeline, &buf, sizeof(buf)));
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
