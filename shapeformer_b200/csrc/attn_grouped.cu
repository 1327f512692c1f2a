// Decode attention for batches whose rows come in groups of G identical conditionings (the reference's sample_n expansion,
// shapeformer.py:229): CausalSelfAttention.forward for the newest query of every row (transformer/mingpt.py:74-91) + KV append.
//
// The conditioning-prefix K/V of a group is bit-identical in its G rows, so it is streamed ONCE per (group, head) into shared
// memory and scored against all G queries there — the prefix crosses HBM *and* the L2->SM fabric once per group, not once per
// row (the per-row kernel let siblings hit each other's lines in L2, which saved HBM traffic but not L2 bandwidth).
// Work units, one CTA each (grid = heads x groups x (2 + G)):
//   unit 0, 1      : the two halves of the shared prefix [0, shared_end), G queries per key tile;
//   unit 2 + r     : row r's own keys [shared_end, pos) + the new position (appended to the cache here), one query.
// Every unit streams its keys as 16-key tiles (K 4 KB + V 4 KB, contiguous runs of the [row][head][position][64] cache) with
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx, L2 evict-first: the cache is touched once per launch) through a
// 3-stage shared-memory ring; a half-warp owns one key at a time (16 lanes x float4 = one 256-byte row, conflict-free), dot
// products finish with 4 warp shuffles, online softmax per half-warp, states merged through shared memory.
// The 3 partial softmax states of a (row, head) are merged by whichever unit finishes last (atomic arrival counter, fixed
// merge order -> bitwise deterministic); no combine kernel.
#include "ar_kernels.cuh"
#include "tc_common.cuh"

namespace sfb {

using namespace tc;

constexpr int AG_TK = 16;                 // keys per tile
constexpr int AG_NS = 5;                  // ring stages (40 KB per CTA: 5 CTAs per SM keep ~160 KB of K/V in flight)
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

// Streams keys [t_beg, t_end) of one (row, head) run and updates NQ online-softmax states per half-warp.
template <int NQ>
__device__ __forceinline__ void ag_stream(AgState (&st)[NQ], const float4 (&q4)[NQ], const float *kbase, const float *vbase, int t_beg,
                                          int t_end, unsigned char *ring, uint64_t *full, uint32_t &it, uint64_t pol) {
    const int tid = threadIdx.x, lane = tid & 31, hw = (tid >> 5) * 2 + (lane >> 4), c = lane & 15;
    const int n_tiles = (t_end - t_beg + AG_TK - 1) / AG_TK;
    auto issue = [&](int tile) {     // thread 0 only
        const uint32_t s = (it + tile) % AG_NS;
        const int t0 = t_beg + tile * AG_TK, nk = min(AG_TK, t_end - t0);
        const uint32_t bytes = (uint32_t)nk * 256u;
        mbar_expect_tx_ag(&full[s], 2 * bytes);
        bulk_load_hint(ring + s * 2 * AG_TILE, kbase + (size_t)t0 * 64, bytes, &full[s], pol);
        bulk_load_hint(ring + s * 2 * AG_TILE + AG_TILE, vbase + (size_t)t0 * 64, bytes, &full[s], pol);
    };
    if (tid == 0)
        for (int t = 0; t < AG_NS - 1 && t < n_tiles; ++t) issue(t);
    for (int tile = 0; tile < n_tiles; ++tile) {
        const uint32_t s = (it + tile) % AG_NS;
        // every thread finished tile-1 (whose slot is the one refilled now) at the barrier that closed the previous iteration
        if (tid == 0 && tile + AG_NS - 1 < n_tiles) issue(tile + AG_NS - 1);
        mbar_wait(&full[s], ((it + tile) / AG_NS) & 1);
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
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float s0 = q4[q].x * k0.x;
            s0 = fmaf(q4[q].y, k0.y, s0); s0 = fmaf(q4[q].z, k0.z, s0); s0 = fmaf(q4[q].w, k0.w, s0);
            float s1 = q4[q].x * k1.x;
            s1 = fmaf(q4[q].y, k1.y, s1); s1 = fmaf(q4[q].z, k1.z, s1); s1 = fmaf(q4[q].w, k1.w, s1);
            s0 = hw_sum(s0); s1 = hw_sum(s1);
            s0 = ok0 ? s0 : -INFINITY; s1 = ok1 ? s1 : -INFINITY;
            const float mx = fmaxf(st[q].m, fmaxf(s0, s1));
            if (mx != -INFINITY) {
                const float corr = expf(st[q].m - mx);     // st.m == -inf -> 0
                const float p0 = expf(s0 - mx), p1 = expf(s1 - mx);
                st[q].l = fmaf(st[q].l, corr, p0) + p1;
                st[q].acc.x = fmaf(p1, v1.x, fmaf(p0, v0.x, st[q].acc.x * corr));
                st[q].acc.y = fmaf(p1, v1.y, fmaf(p0, v0.y, st[q].acc.y * corr));
                st[q].acc.z = fmaf(p1, v1.z, fmaf(p0, v0.z, st[q].acc.z * corr));
                st[q].acc.w = fmaf(p1, v1.w, fmaf(p0, v0.w, st[q].acc.w * corr));
                st[q].m = mx;
            }
        }
        __syncthreads();     // the slot may be refilled
    }
    it += n_tiles;
}

// Merge the 8 half-warp states of the CTA for one query, write the partial state of (row b, head h, part u) and, if this
// was the last of the AG_PARTS parts to arrive, combine them (fixed order) into out[b][h*64 ..].
__device__ __forceinline__ void ag_finish(const AgState &st, float *sm, float *part, int *cnt, float *out, int b, int h, int H, int u) {
    const int tid = threadIdx.x, lane = tid & 31, hw = (tid >> 5) * 2 + (lane >> 4), c = lane & 15;
    float *mine = sm + hw * AG_MRG;
    if (c == 0) { mine[64] = st.m; mine[65] = st.l; }
    st4(mine + c * 4, st.acc);
    __syncthreads();
    if (tid < 32) {
        float M = -INFINITY;
#pragma unroll
        for (int g = 0; g < 8; ++g) M = fmaxf(M, sm[g * AG_MRG + 64]);
        float Ls = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float mg = sm[g * AG_MRG + 64];
            const float w = (mg == -INFINITY) ? 0.f : expf(mg - M);
            Ls = fmaf(sm[g * AG_MRG + 65], w, Ls);
            o0 = fmaf(sm[g * AG_MRG + 2 * lane], w, o0);
            o1 = fmaf(sm[g * AG_MRG + 2 * lane + 1], w, o1);
        }
        float *p = part + (((size_t)b * H + h) * AG_PARTS + u) * AG_PSTRIDE;
        __stcg(reinterpret_cast<float2 *>(p + 2 * lane), make_float2(o0, o1));
        if (lane == 0) __stcg(reinterpret_cast<float2 *>(p + 64), make_float2(M, Ls));
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
            *reinterpret_cast<float2 *>(out + (size_t)b * H * 64 + h * 64 + 2 * lane) = make_float2(a0 * inv, a1 * inv);
            if (lane == 0) cnt[(size_t)b * H + h] = 0;     // ready for the next launch
        }
    }
    __syncthreads();     // `sm` is reused by the next query
}

template <int G>
__global__ void __launch_bounds__(128) attn_grouped_kernel(const float *qkv, float *kcache, float *vcache, float *out, float *part,
                                                           int *cnt, int H, int max_len, int pos_arg, const int32_t *st_dev,
                                                           int lcond_arg, int lcond_delta) {
    pdl_trigger();
    __shared__ __align__(128) unsigned char ring[AG_NS * 2 * AG_TILE];
    __shared__ __align__(16) float sm[8 * AG_MRG];
    __shared__ __align__(8) uint64_t full[AG_NS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = lane >> 4, c = lane & 15;
    const int h = blockIdx.x, grp = blockIdx.y, u = blockIdx.z;
    if (tid == 0) {
        for (int i = 0; i < AG_NS; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint64_t pol = policy_evict_first();
    pdl_wait();
    const int pos = st_dev ? st_dev[ST_LEN] - 1 : pos_arg;
    const int d = H * 64;
    const int shared_end = min((st_dev ? st_dev[ST_LCOND] : lcond_arg) + lcond_delta, pos);
    uint32_t it = 0;
    if (u < 2) {
        // ---- shared prefix half u of the group, all G queries; keys come from the group LEADER's cache rows
        int mid = ((shared_end / 2) + AG_TK - 1) / AG_TK * AG_TK;
        if (mid > shared_end) mid = shared_end;
        const int t_beg = u == 0 ? 0 : mid, t_end = u == 0 ? mid : shared_end;
        const int b0 = grp * G;
        float4 q4[G];
        AgState st[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            q4[q] = ld4(qkv + (size_t)(b0 + q) * 3 * d + h * 64 + c * 4);
            q4[q].x *= 0.125f; q4[q].y *= 0.125f; q4[q].z *= 0.125f; q4[q].w *= 0.125f;    // 1/sqrt(64), exact
            st[q].m = -INFINITY; st[q].l = 0.f; st[q].acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const size_t base = ((size_t)b0 * H + h) * (size_t)max_len * 64;
        ag_stream<G>(st, q4, kcache + base, vcache + base, t_beg, t_end, ring, full, it, pol);
#pragma unroll
        for (int q = 0; q < G; ++q) ag_finish(st[q], sm, part, cnt, out, b0 + q, h, H, u);
    } else {
        // ---- row b's own keys [shared_end, pos) and the new position (appended here)
        const int b = grp * G + (u - 2);
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
        ag_stream<1>(st, q4, kcache + base, vcache + base, shared_end, pos, ring, full, it, pol);
        ag_finish(st[0], sm, part, cnt, out, b, h, H, 2);
    }
}

int launch_attn_grouped(const float *qkv, float *kc, float *vc, float *out, float *part, int *cnt, int B, int H, int max_len, int pos,
                        const int32_t *st, int group, int lcond, int lcond_delta, cudaStream_t s) {
    if (group < 2 || B % group != 0 || !part || !cnt) return SFB200_E_ARG;
    const dim3 grid(H, B / group, 2 + group), block(128);
#define SFB_AG_CASE(GG)                                                                                                         \
    case GG:                                                                                                                    \
        return launch_ex("attn_grouped", attn_grouped_kernel<GG>, grid, block, 0, s, dim3(1, 1, 1), qkv, kc, vc, out, part, cnt, H, \
                         max_len, pos, st, lcond, lcond_delta);
    switch (group) {
        SFB_AG_CASE(2)
        SFB_AG_CASE(4)
        SFB_AG_CASE(6)
        SFB_AG_CASE(8)
    }
#undef SFB_AG_CASE
    return SFB200_E_ARG;
}

}  // namespace sfb
