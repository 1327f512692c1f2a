// Decode attention for batches whose rows come in groups of G identical conditionings (the reference's sample_n expansion,
// shapeformer.py:229): CausalSelfAttention.forward for the newest query of every row (transformer/mingpt.py:74-91) + KV append.
//
// The conditioning-prefix K/V of a group is bit-identical in its G rows, so it is streamed ONCE per (group, head) into shared
// memory and scored against all G queries there — the prefix crosses HBM *and* the L2->SM fabric once per group, not once per
// row (the per-row kernel let siblings hit each other's lines in L2, which saved HBM traffic but not L2 bandwidth).
//
// Grid = heads x groups x 2 CTAs, sized so that the whole launch is ONE co-resident wave (512 CTAs of 51 KB for the benchmarked
// 64 rows in groups of 4: 4 CTAs per SM fit, 592 slots).  CTA (h, grp, u) streams a sequence of SEGMENTS through one
// shared-memory ring that keeps running across segment boundaries (no start-up bubble per segment); a producer warp issues the
// copies, four consumer warps score the keys, decoupled by full/empty mbarriers (no CTA barrier per tile):
//   segment 0        : half u of the shared prefix [0, shared_end) of the group, G queries per key tile;
//   segment 1 .. G/2 : the own keys [shared_end, pos) of row u*G/2 + j - 1 of the group, one query, + the new position
//                      (appended to the cache here).
// Keys travel as 16-key tiles (K 4 KB + V 4 KB, contiguous runs of the [row][head][position][64] cache) with TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx, L2 evict-first: the cache is touched once per launch), 4 tiles in flight per CTA;
// a half-warp owns one key at a time (16 lanes x float4 = one 256-byte row, conflict-free), dot products finish with 4 warp
// shuffles, online softmax per half-warp, states merged through shared memory.
// The 3 partial softmax states of a (row, head) — prefix half 0, prefix half 1, own keys — are merged by whichever CTA
// delivers the last one (atomic arrival counter, fixed merge order -> bitwise deterministic); no combine kernel.
#include "ar_kernels.cuh"
#include "tc_common.cuh"

namespace sfb {

using namespace tc;

constexpr int AG_TK = 16;                 // keys per tile
constexpr int AG_NS = 5;                  // ring stages (40 KB per CTA; 4 tiles = 32 KB in flight per CTA, 4 CTAs per SM)
constexpr int AG_TILE = AG_TK * 64 * 4;   // 4 KB (K or V)
constexpr int AG_MRG = 68;                // merge row stride (floats)
constexpr int AG_PARTS = 3;               // partial states per (row, head)
constexpr int AG_PSTRIDE = 66;            // 64 acc + m + l

__device__ __forceinline__ float hw_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_ag(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

struct AgState {
    float m, l;
    float4 acc;
};

// Consumes the tiles of one segment (streamed by the producer warp) and updates NQ online-softmax states per half-warp.  The four
// consumer warps never meet at a CTA barrier inside the stream: a warp waits for the tile (full[s]), reads its keys, and hands the
// slot back with one mbarrier arrival (empty[s], 4 arrivals = slot free), so warps drift apart by up to the ring depth.
template <int NQ>
__device__ __forceinline__ void ag_stream(AgState (&st)[NQ], const float4 (&q4)[NQ], int t_beg, int t_end, unsigned char *ring,
                                          uint64_t *full, uint64_t *empty, uint32_t &it) {
    const int tid = threadIdx.x, lane = tid & 31, hw = (tid >> 5) * 2 + (lane >> 4), c = lane & 15;
    const int n_tiles = (t_end - t_beg + AG_TK - 1) / AG_TK;
    for (int tile = 0; tile < n_tiles; ++tile, ++it) {
        const uint32_t s = it % AG_NS;
        mbar_wait(&full[s], (it / AG_NS) & 1);
        const float *ks = reinterpret_cast<const float *>(ring + s * 2 * AG_TILE), *vs = ks + AG_TK * 64;
        const int nk = min(AG_TK, t_end - (t_beg + tile * AG_TK));
        // this half-warp's keys of the tile: hw and hw + 8
        const float4 k0 = ld4(ks + hw * 64 + c * 4), k1 = ld4(ks + (hw + 8) * 64 + c * 4);
        const float4 v0r = ld4(vs + hw * 64 + c * 4), v1r = ld4(vs + (hw + 8) * 64 + c * 4);
        const bool ok0 = hw < nk, ok1 = hw + 8 < nk;
        // rows past the end of a partial tile hold stale shared memory: their scores become -inf (p = 0) and their V rows are
        // zeroed so that 0 * garbage cannot produce a NaN
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v0 = ok0 ? v0r : z4, v1 = ok1 ? v1r : z4;
        float s0[NQ], s1[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float a0 = q4[q].x * k0.x;
            a0 = fmaf(q4[q].y, k0.y, a0); a0 = fmaf(q4[q].z, k0.z, a0); a0 = fmaf(q4[q].w, k0.w, a0);
            float a1 = q4[q].x * k1.x;
            a1 = fmaf(q4[q].y, k1.y, a1); a1 = fmaf(q4[q].z, k1.z, a1); a1 = fmaf(q4[q].w, k1.w, a1);
            s0[q] = a0; s1[q] = a1;
        }
        // the operands are in registers: release the slot before the arithmetic (the refill overlaps it)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
#pragma unroll
        for (int q = 0; q < NQ; ++q) { s0[q] = hw_sum(s0[q]); s1[q] = hw_sum(s1[q]); }     // 2 NQ independent shuffle chains
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float x0 = ok0 ? s0[q] : -INFINITY, x1 = ok1 ? s1[q] : -INFINITY;
            const float mx = fmaxf(st[q].m, fmaxf(x0, x1));
            const float ms = mx == -INFINITY ? 0.f : mx;     // nothing seen yet: every exponential below is exp(-inf) = 0
            const float corr = expf(st[q].m - ms);
            const float p0 = expf(x0 - ms), p1 = expf(x1 - ms);
            st[q].l = fmaf(st[q].l, corr, p0) + p1;
            st[q].acc.x = fmaf(p1, v1.x, fmaf(p0, v0.x, st[q].acc.x * corr));
            st[q].acc.y = fmaf(p1, v1.y, fmaf(p0, v0.y, st[q].acc.y * corr));
            st[q].acc.z = fmaf(p1, v1.z, fmaf(p0, v0.z, st[q].acc.z * corr));
            st[q].acc.w = fmaf(p1, v1.w, fmaf(p0, v0.w, st[q].acc.w * corr));
            st[q].m = mx;
        }
    }
}

// Merge the 8 half-warp states of the CTA for NQ queries (consumer warp w merges queries w, w + 4, ...) into the CTA's
// partial-state table in shared memory: sp[(p0 + q) * AG_PSTRIDE ..] = 64 acc | m | l.   `sm`: NQ x 8 x AG_MRG floats.
template <int NQ>
__device__ __forceinline__ void ag_merge(const AgState (&st)[NQ], float *sm, float *sp, int p0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, hw = warp * 2 + (lane >> 4), c = lane & 15;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float *mine = sm + (q * 8 + hw) * AG_MRG;
        if (c == 0) { mine[64] = st[q].m; mine[65] = st[q].l; }
        st4(mine + c * 4, st[q].acc);
    }
    bar_sync(1, 128);
    for (int q = warp; q < NQ; q += 4) {
        const float *sq = sm + q * 8 * AG_MRG;
        float M = -INFINITY;
#pragma unroll
        for (int g = 0; g < 8; ++g) M = fmaxf(M, sq[g * AG_MRG + 64]);
        float Ls = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float mg = sq[g * AG_MRG + 64];
            const float w = (mg == -INFINITY) ? 0.f : expf(mg - M);
            Ls = fmaf(sq[g * AG_MRG + 65], w, Ls);
            o0 = fmaf(sq[g * AG_MRG + 2 * lane], w, o0);
            o1 = fmaf(sq[g * AG_MRG + 2 * lane + 1], w, o1);
        }
        float *p = sp + (p0 + q) * AG_PSTRIDE;
        *reinterpret_cast<float2 *>(p + 2 * lane) = make_float2(o0, o1);
        if (lane == 0) *reinterpret_cast<float2 *>(p + 64) = make_float2(M, Ls);
    }
    bar_sync(1, 128);     // `sm` is reused by the next segment; sp entries are complete
}

// Publish partial state `sp_e` as part u of (row b, head h); the warp that delivers the last of the AG_PARTS parts combines them
// (fixed order) into out[b][h*64 ..].  One warp per call; the memory fence and the atomic are paid once per part, at the end of
// the CTA, by four warps in parallel.
__device__ __forceinline__ void ag_publish(const float *sp_e, float *part, int *cnt, float *out, float *out_lo, int b, int h, int H, int u) {
    const int lane = threadIdx.x & 31;
    float *p = part + (((size_t)b * H + h) * AG_PARTS + u) * AG_PSTRIDE;
    __stcg(reinterpret_cast<float2 *>(p + 2 * lane), *reinterpret_cast<const float2 *>(sp_e + 2 * lane));
    if (lane == 0) __stcg(reinterpret_cast<float2 *>(p + 64), *reinterpret_cast<const float2 *>(sp_e + 64));
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atomicAdd(cnt + (size_t)b * H + h, 1) == AG_PARTS - 1;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        __threadfence();
        const float *pp = part + ((size_t)b * H + h) * AG_PARTS * AG_PSTRIDE;
        float ms[AG_PARTS], ls[AG_PARTS];
        float2 os[AG_PARTS];
        float MM = -INFINITY;
#pragma unroll
        for (int i = 0; i < AG_PARTS; ++i) {
            const float2 ml = __ldcg(reinterpret_cast<const float2 *>(pp + i * AG_PSTRIDE + 64));
            ms[i] = ml.x; ls[i] = ml.y;
            os[i] = __ldcg(reinterpret_cast<const float2 *>(pp + i * AG_PSTRIDE + 2 * lane));
            MM = fmaxf(MM, ms[i]);
        }
        float L = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < AG_PARTS; ++i) {
            const float w = (ms[i] == -INFINITY) ? 0.f : expf(ms[i] - MM);
            L = fmaf(ls[i], w, L);
            a0 = fmaf(os[i].x, w, a0);
            a1 = fmaf(os[i].y, w, a1);
        }
        const float inv = 1.0f / L;
        const float o0 = a0 * inv, o1 = a1 * inv;
        *reinterpret_cast<float2 *>(out + (size_t)b * H * 64 + h * 64 + 2 * lane) = make_float2(o0, o1);
        // low part of the TF32 operand split, for a tc_big projection GEMM that reads it from memory
        if (out_lo) *reinterpret_cast<float2 *>(out_lo + (size_t)b * H * 64 + h * 64 + 2 * lane) = make_float2(tf32_lo(o0), tf32_lo(o1));
        if (lane == 0) cnt[(size_t)b * H + h] = 0;     // ready for the next launch
    }
}

constexpr int AG_THREADS = 160;     // warps 0-3 consume, warp 4 produces (TMA bulk copies)

template <int G, bool MONO>
__global__ void __launch_bounds__(AG_THREADS) attn_grouped_kernel(const float *qkv, float *kcache, float *vcache, float *out, float *part,
                                                                  int *cnt, int H, int max_len, int pos_arg, const int32_t *st_dev,
                                                                  int lcond_arg, int lcond_delta, float *out_lo) {
    pdl_trigger();
    extern __shared__ __align__(128) unsigned char ag_smem[];
    // MONO (large batches: heads x groups alone fill the GPU several times): ONE CTA per (head, group) streams the whole prefix
    // and the own keys of all G rows, and combines the two states of every row in shared memory — no partials in global memory,
    // no fence, no arrival counter.  Otherwise two CTAs per (head, group) split the work (single co-resident wave at 64 rows).
    constexpr int RPC = MONO ? G : G / 2;                                            // rows whose own keys this CTA streams
    constexpr int NPART = G + RPC;                                                   // partial states this CTA publishes
    unsigned char *ring = ag_smem;                                                   // [NS][K tile | V tile]
    float *sm = reinterpret_cast<float *>(ag_smem + AG_NS * 2 * AG_TILE);            // [G][8][AG_MRG] merge staging
    float *sp = sm + G * 8 * AG_MRG;                                                 // [NPART][AG_PSTRIDE] partial states
    uint64_t *full = reinterpret_cast<uint64_t *>(sp + NPART * AG_PSTRIDE);          // [NS] tile landed (TMA tx bytes)
    uint64_t *empty = full + AG_NS;                                                  // [NS] tile consumed by the 4 warps
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = lane >> 4, c = lane & 15;
    const int h = blockIdx.x, grp = blockIdx.y, u = blockIdx.z;
    if (tid == 0) {
        for (int i = 0; i < AG_NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    const int pos = st_dev ? st_dev[ST_LEN] - 1 : pos_arg;
    const int d = H * 64;
    const int shared_end = min((st_dev ? st_dev[ST_LCOND] : lcond_arg) + lcond_delta, pos);
    const int b0 = grp * G, r0 = b0 + u * RPC;
    int mid = ((shared_end / 2) + AG_TK - 1) / AG_TK * AG_TK;
    if (mid > shared_end) mid = shared_end;
    const int p_beg = (MONO || u == 0) ? 0 : mid, p_end = (!MONO && u == 0) ? mid : shared_end;

    if (warp == 4) {
        // ================================ producer: every tile of every segment, in order ================================
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            uint32_t j = 0;
            for (int sgi = 0; sgi <= RPC; ++sgi) {
                const int row = sgi == 0 ? b0 : r0 + sgi - 1;
                const size_t base = ((size_t)row * H + h) * (size_t)max_len * 64;
                const float *kb = kcache + base, *vb = vcache + base;
                const int t_beg = sgi == 0 ? p_beg : shared_end, t_end = sgi == 0 ? p_end : pos;
                for (int t0 = t_beg; t0 < t_end; t0 += AG_TK, ++j) {
                    const uint32_t s = j % AG_NS;
                    if (j >= AG_NS) mbar_wait(&empty[s], ((j / AG_NS) - 1) & 1);
                    const uint32_t bytes = (uint32_t)min(AG_TK, t_end - t0) * 256u;
                    mbar_expect_tx_ag(&full[s], 2 * bytes);
                    bulk_load_hint(ring + s * 2 * AG_TILE, kb + (size_t)t0 * 64, bytes, &full[s], pol);
                    bulk_load_hint(ring + s * 2 * AG_TILE + AG_TILE, vb + (size_t)t0 * 64, bytes, &full[s], pol);
                }
            }
        }
        return;
    }

    // ================================ consumers ================================
    uint32_t it = 0;
    {
        // ---- shared prefix half u of the group, all G queries; keys come from the group LEADER's cache rows
        float4 q4[G];
        AgState st[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            q4[q] = ld4(qkv + (size_t)(b0 + q) * 3 * d + h * 64 + c * 4);
            q4[q].x *= 0.125f; q4[q].y *= 0.125f; q4[q].z *= 0.125f; q4[q].w *= 0.125f;    // 1/sqrt(64), exact
            st[q].m = -INFINITY; st[q].l = 0.f; st[q].acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ag_stream<G>(st, q4, p_beg, p_end, ring, full, empty, it);
        ag_merge<G>(st, sm, sp, 0);
    }
    for (int j = 0; j < RPC; ++j) {
        // ---- row b's own keys [shared_end, pos) and the new position (appended here)
        const int b = r0 + j;
        const float *qrow = qkv + (size_t)b * 3 * d + h * 64 + c * 4;
        float4 q4[1];
        q4[0] = ld4(qrow);
        q4[0].x *= 0.125f; q4[0].y *= 0.125f; q4[0].z *= 0.125f; q4[0].w *= 0.125f;
        AgState st[1];
        st[0].m = -INFINITY; st[0].l = 0.f; st[0].acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t base = ((size_t)b * H + h) * (size_t)max_len * 64;
        if (warp == 0) {
            // both halves run the shuffle reduction (full-mask shuffles need all 32 lanes); only half 0 keeps the state
            const float4 kn = ld4(qrow + d), vn = ld4(qrow + 2 * d);
            float p = q4[0].x * kn.x;
            p = fmaf(q4[0].y, kn.y, p); p = fmaf(q4[0].z, kn.z, p); p = fmaf(q4[0].w, kn.w, p);
            p = hw_sum(p);
            if (half == 0) {
                st4(kcache + base + (size_t)pos * 64 + c * 4, kn);
                st4(vcache + base + (size_t)pos * 64 + c * 4, vn);
                st[0].m = p; st[0].l = 1.f; st[0].acc = vn;
            }
        }
        ag_stream<1>(st, q4, shared_end, pos, ring, full, empty, it);
        ag_merge<1>(st, sm, sp, G + j);
    }
    if (MONO) {
        // ---- combine in place: row q = prefix state sp[q] + own state sp[G + q] (fixed order -> deterministic)
        for (int q = warp; q < G; q += 4) {
            const float *pa = sp + q * AG_PSTRIDE, *pb = sp + (G + q) * AG_PSTRIDE;
            const float ma = pa[64], mb = pb[64], MM = fmaxf(ma, mb);
            const float wa = (ma == -INFINITY) ? 0.f : expf(ma - MM), wb = (mb == -INFINITY) ? 0.f : expf(mb - MM);
            const float L = fmaf(pb[65], wb, pa[65] * wa);
            const float2 oa = *reinterpret_cast<const float2 *>(pa + 2 * lane), ob = *reinterpret_cast<const float2 *>(pb + 2 * lane);
            const float inv = 1.0f / L;
            const float o0 = fmaf(ob.x, wb, oa.x * wa) * inv, o1 = fmaf(ob.y, wb, oa.y * wa) * inv;
            const size_t off = (size_t)(b0 + q) * H * 64 + h * 64 + 2 * lane;
            *reinterpret_cast<float2 *>(out + off) = make_float2(o0, o1);
            if (out_lo) *reinterpret_cast<float2 *>(out_lo + off) = make_float2(tf32_lo(o0), tf32_lo(o1));
        }
        return;
    }
    // ---- publish: parts 0 .. G-1 = prefix half u of rows b0 + q; parts G + j = own keys (part index 2) of rows r0 + j
    for (int e = warp; e < NPART; e += 4) {
        const bool own = e >= G;
        ag_publish(sp + e * AG_PSTRIDE, part, cnt, out, out_lo, own ? r0 + (e - G) : b0 + e, h, H, own ? 2 : u);
    }
}

template <int G>
constexpr int ag_smem_bytes() { return AG_NS * 2 * AG_TILE + G * 8 * AG_MRG * 4 + 2 * G * AG_PSTRIDE * 4 + 2 * AG_NS * 8; }

template <int G>
static int launch_ag(const float *qkv, float *kc, float *vc, float *out, float *part, int *cnt, int B, int H, int max_len, int pos,
                     const int32_t *st, int lcond, int lcond_delta, cudaStream_t s, float *out_lo) {
    static unsigned long long attr_done = 0;   // bit per device
    if (first_use_on_device(attr_done)) {
        SFB_CUDA_TRY(cudaFuncSetAttribute(attn_grouped_kernel<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ag_smem_bytes<G>()));
        SFB_CUDA_TRY(cudaFuncSetAttribute(attn_grouped_kernel<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ag_smem_bytes<G>()));
    }
    // one CTA per (head, group) once those alone fill the GPU's CTA slots about twice; two CTAs per (head, group) below that
    if (H * (B / G) >= 1024)
        return launch_ex("attn_grouped", attn_grouped_kernel<G, true>, dim3(H, B / G, 1), dim3(AG_THREADS), ag_smem_bytes<G>(), s,
                         dim3(1, 1, 1), qkv, kc, vc, out, part, cnt, H, max_len, pos, st, lcond, lcond_delta, out_lo);
    return launch_ex("attn_grouped", attn_grouped_kernel<G, false>, dim3(H, B / G, 2), dim3(AG_THREADS), ag_smem_bytes<G>(), s,
                     dim3(1, 1, 1), qkv, kc, vc, out, part, cnt, H, max_len, pos, st, lcond, lcond_delta, out_lo);
}

int launch_attn_grouped(const float *qkv, float *kc, float *vc, float *out, float *part, int *cnt, int B, int H, int max_len, int pos,
                        const int32_t *st, int group, int lcond, int lcond_delta, cudaStream_t s, float *out_lo) {
    if (group < 2 || B % group != 0 || !part || !cnt) return SFB200_E_ARG;
    switch (group) {
        case 2: return launch_ag<2>(qkv, kc, vc, out, part, cnt, B, H, max_len, pos, st, lcond, lcond_delta, s, out_lo);
        case 4: return launch_ag<4>(qkv, kc, vc, out, part, cnt, B, H, max_len, pos, st, lcond, lcond_delta, s, out_lo);
        case 6: return launch_ag<6>(qkv, kc, vc, out, part, cnt, B, H, max_len, pos, st, lcond, lcond_delta, s, out_lo);
        case 8: return launch_ag<8>(qkv, kc, vc, out, part, cnt, B, H, max_len, pos, st, lcond, lcond_delta, s, out_lo);
    }
    return SFB200_E_ARG;
}

}  // namespace sfb
