// Causal prefill attention on the 5th-generation tensor cores: CausalSelfAttention.forward over the conditioning positions
// (transformer/mingpt.py:74-91) — the dense QK^T and PV contractions of the path — with fp32-level accuracy (3xTF32), + the K/V
// cache fill.  qkv (B, T, 3d) -> out (B, T, d), caches [row][head][0..T)[64].
//
// CTA = one (row, head, 128-query tile), 128 threads = 128 query rows = the 128 TMEM lanes.  Keys are visited in blocks of 64 up to
// the causal limit, flash-attention style with the softmax state in registers:
//   1. K block and V block -> shared memory as K-major SWIZZLE_128B tiles (V transposed: [dim][key]); hi operand = the raw fp32
//      value (the tensor core truncates it to tf32), lo operand = rna_tf32(v - trunc_tf32(v)), both written by the loading threads;
//   2. S = (Q / 8) K^T : 24 tcgen05.mma.kind::tf32 (2 x 32-dim chunks x 4 k-steps x 3 products), Q from shared memory, D in TMEM;
//   3. thread = query row: tcgen05.ld of its 64 scores, causal mask, online softmax (running max / sum, rescale of the output
//      registers), P -> hi | lo written back into TENSOR MEMORY with tcgen05.st (the A operand of the next GEMM lives in TMEM);
//   4. O_blk = P V : 24 MMAs (8 k-steps of 8 keys x 3 products), accumulated from zero; 5. O += O_blk in fp32 registers — every
//      tensor-core accumulation chain is 24 MMAs long (the TMEM accumulator is not round-to-nearest, see tc_gemm.cu).
// One elected thread issues the MMAs; phases are separated by CTA barriers + one mbarrier for MMA completion (no overlap between
// the phases: prefill is ~1 % of a batch; the point is tensor-core contractions with fp32-grade results, 6x the FFMA kernel).
#include "ar_kernels.cuh"
#include "tc_common.cuh"

namespace sfb {

using namespace tc;

constexpr int PT_THREADS = 128;
constexpr int PT_KB = 64;                               // keys per block
constexpr int PT_OFF_QH = 0;                            // Q hi: 2 chunks x (128 rows x 128 B)
constexpr int PT_OFF_QL = PT_OFF_QH + 2 * 16384;
constexpr int PT_OFF_KH = PT_OFF_QL + 2 * 16384;        // K hi: 2 chunks x (64 keys x 128 B)
constexpr int PT_OFF_KL = PT_OFF_KH + 2 * 8192;
constexpr int PT_OFF_VH = PT_OFF_KL + 2 * 8192;         // V^T hi: 2 key chunks x (64 dims x 128 B = 32 keys)
constexpr int PT_OFF_VL = PT_OFF_VH + 2 * 8192;
constexpr int PT_OFF_BAR = PT_OFF_VL + 2 * 8192;
constexpr int PT_SMEM = PT_OFF_BAR + 64;
constexpr int PT_COL_S = 0, PT_COL_PH = 64, PT_COL_PL = 128, PT_COL_O = 192;   // TMEM columns

// byte offset of element (row r, float j in 0..31) inside a K-major SWIZZLE_128B tile (rows of 128 B, 1024-byte aligned base)
__device__ __forceinline__ int sw128(int r, int j) { return r * 128 + ((((j >> 2) ^ (r & 7)) << 4) | ((j & 3) << 2)); }

__global__ void __launch_bounds__(PT_THREADS, 1) attn_prefill_tc_kernel(const float *qkv, float *kcache, float *vcache, float *out, int H,
                                                                        int T, int max_len, const int32_t *rowmap) {
    extern __shared__ __align__(1024) unsigned char pt_smem[];
    unsigned char *smem = pt_smem;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + PT_OFF_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);

    pdl_trigger();
    const int tid = threadIdx.x, warp = tid >> 5;
    const int h = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * 128;
    const int d = H * 64;
    const size_t rs = (size_t)3 * d;                                  // floats between positions
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
    pdl_wait();
    const int br = rowmap ? rowmap[b] : b;                           // KV-cache row
    const float *rowb = qkv + (size_t)b * T * rs + h * 64;

    // ---- Q tile (scaled by 1/sqrt(64), exact) -> hi / lo tiles; K, V rows of this tile's positions -> cache
    for (int i = 0; i < 16; ++i) {
        const int idx = tid + PT_THREADS * i, r = idx >> 4, c4 = idx & 15;
        const int t = q0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < T) {
            const float *src = rowb + (size_t)t * rs + c4 * 4;
            v = ld4(src);
            const size_t base = (((size_t)br * H + h) * (size_t)max_len + t) * 64 + c4 * 4;
            st4(kcache + base, ld4(src + d));
            st4(vcache + base, ld4(src + 2 * d));
        }
        v.x *= 0.125f; v.y *= 0.125f; v.z *= 0.125f; v.w *= 0.125f;
        const int off = (c4 >> 3) * 16384 + sw128(r, (c4 & 7) * 4);
        *reinterpret_cast<float4 *>(smem + PT_OFF_QH + off) = v;
        *reinterpret_cast<float4 *>(smem + PT_OFF_QL + off) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
    }

    const int q = q0 + tid;                                          // this thread's query position
    float o[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) o[j] = 0.f;
    float m = -INFINITY, l = 0.f;
    uint32_t ph = 0;
    const int k_end = min(T, q0 + 128);
    constexpr uint32_t IDESC = instr_desc(2, 128, PT_KB);
    const uint32_t s0 = smem_u32(smem);

    for (int k0 = 0; k0 < k_end; k0 += PT_KB) {
        // ---- 1. K block [64 keys x 64 dims] and V block transposed [64 dims x 64 keys] -> hi / lo tiles
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + PT_THREADS * i, r = idx >> 4, c4 = idx & 15;
            const int t = k0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < T) v = ld4(rowb + (size_t)t * rs + d + c4 * 4);
            const int off = (c4 >> 3) * 8192 + sw128(r, (c4 & 7) * 4);
            *reinterpret_cast<float4 *>(smem + PT_OFF_KH + off) = v;
            *reinterpret_cast<float4 *>(smem + PT_OFF_KL + off) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
        }
        {
            const int kk = tid & 63, j0 = (tid >> 6) * 32;           // key of the block, first of this thread's 32 dims
            const int t = k0 + kk;
            const int cb = (kk >> 5) * 8192;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < T) v = ld4(rowb + (size_t)t * rs + 2 * d + j0 + j4 * 4);
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int off = cb + sw128(j0 + j4 * 4 + u, kk & 31);
                    *reinterpret_cast<float *>(smem + PT_OFF_VH + off) = e[u];
                    *reinterpret_cast<float *>(smem + PT_OFF_VL + off) = tf32_lo(e[u]);
                }
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- 2. S = Q K^T
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint64_t qh = smem_desc_k128(s0 + PT_OFF_QH + c * 16384), ql = smem_desc_k128(s0 + PT_OFF_QL + c * 16384);
                    const uint64_t kh = smem_desc_k128(s0 + PT_OFF_KH + c * 8192), kl = smem_desc_k128(s0 + PT_OFF_KL + c * 8192);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        mma_tf32_ss(tmem_base + PT_COL_S, ql + 2 * k, kh + 2 * k, IDESC, !(c == 0 && k == 0));
                        mma_tf32_ss(tmem_base + PT_COL_S, qh + 2 * k, kl + 2 * k, IDESC, 1);
                        mma_tf32_ss(tmem_base + PT_COL_S, qh + 2 * k, kh + 2 * k, IDESC, 1);
                    }
                }
                mma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph); ph ^= 1;
        tc_fence_after();
        // ---- 3. scores of this query row -> mask -> online softmax -> P (hi | lo) into tensor memory
        {
            uint32_t sv[64];
            tmem_ld32(tmem_base + lane_off + PT_COL_S, reinterpret_cast<uint32_t(&)[32]>(sv[0]));
            tmem_ld32(tmem_base + lane_off + PT_COL_S + 32, reinterpret_cast<uint32_t(&)[32]>(sv[32]));
            tmem_ld_wait();
            float mx = m;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const int kidx = k0 + j;
                const float s = (kidx <= q && kidx < T) ? __uint_as_float(sv[j]) : -INFINITY;
                sv[j] = __float_as_uint(s);
                mx = fmaxf(mx, s);
            }
            const float corr = expf(m - mx);                         // m == -inf (first block) -> 0; key 0 is visible to every row
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float p = expf(__uint_as_float(sv[j]) - mx);
                sum += p;
                sv[j] = __float_as_uint(p);
            }
            l = fmaf(l, corr, sum);
            m = mx;
#pragma unroll
            for (int j = 0; j < 64; ++j) o[j] *= corr;
            tmem_st32(tmem_base + lane_off + PT_COL_PH, reinterpret_cast<uint32_t(&)[32]>(sv[0]));
            tmem_st32(tmem_base + lane_off + PT_COL_PH + 32, reinterpret_cast<uint32_t(&)[32]>(sv[32]));
#pragma unroll
            for (int j = 0; j < 64; ++j) sv[j] = __float_as_uint(tf32_lo(__uint_as_float(sv[j])));
            tmem_st32(tmem_base + lane_off + PT_COL_PL, reinterpret_cast<uint32_t(&)[32]>(sv[0]));
            tmem_st32(tmem_base + lane_off + PT_COL_PL + 32, reinterpret_cast<uint32_t(&)[32]>(sv[32]));
            tmem_st_wait();
        }
        tc_fence_before();
        __syncthreads();
        // ---- 4. O_blk = P V
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t vh = smem_desc_k128(s0 + PT_OFF_VH + (ks >> 2) * 8192) + 2 * (ks & 3);
                    const uint64_t vl = smem_desc_k128(s0 + PT_OFF_VL + (ks >> 2) * 8192) + 2 * (ks & 3);
                    mma_tf32_ts(tmem_base + PT_COL_O, tmem_base + PT_COL_PL + 8 * ks, vh, IDESC, ks != 0);
                    mma_tf32_ts(tmem_base + PT_COL_O, tmem_base + PT_COL_PH + 8 * ks, vl, IDESC, 1);
                    mma_tf32_ts(tmem_base + PT_COL_O, tmem_base + PT_COL_PH + 8 * ks, vh, IDESC, 1);
                }
                mma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph); ph ^= 1;
        tc_fence_after();
        // ---- 5. O += O_blk
        {
            uint32_t ov[64];
            tmem_ld32(tmem_base + lane_off + PT_COL_O, reinterpret_cast<uint32_t(&)[32]>(ov[0]));
            tmem_ld32(tmem_base + lane_off + PT_COL_O + 32, reinterpret_cast<uint32_t(&)[32]>(ov[32]));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 64; ++j) o[j] += __uint_as_float(ov[j]);
        }
        tc_fence_before();
        __syncthreads();      // the K / V tiles and the S / P / O columns may be overwritten
    }

    if (q < T) {
        const float inv = 1.0f / l;
        float *dst = out + ((size_t)b * T + q) * d + h * 64;
#pragma unroll
        for (int j = 0; j < 64; j += 4) st4(dst + j, make_float4(o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

int launch_attn_prefill_tc(const float *qkv, float *kc, float *vc, float *out, int B, int H, int T, int max_len, cudaStream_t s,
                           const int32_t *rowmap) {
    if (T <= 0) return SFB200_OK;
    static unsigned long long attr_done = 0;   // bit per device
    if (first_use_on_device(attr_done))
        SFB_CUDA_TRY(cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM));
    return launch_ex("attn_prefill_tc", attn_prefill_tc_kernel, dim3(H, B, (T + 127) / 128), dim3(PT_THREADS), PT_SMEM, s, dim3(1, 1, 1), qkv,
                     kc, vc, out, H, T, max_len, rowmap);
}

}  // namespace sfb
