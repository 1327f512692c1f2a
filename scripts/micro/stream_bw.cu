// Development microbenchmark: HBM read bandwidth of many concurrent streams fetched with cp.async.bulk (TMA) in chunks of S bytes,
// as the decode-attention kernel does (2 streams per unit: K and V), versus chunk size / ring depth / CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/stream_bw scripts/micro/stream_bw.cu && build/stream_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol, int hint) {
    if (hint)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// unit u: two streams (A = base + u*stride, B = A + half) of `len` bytes each, fetched in chunks of S bytes alternately.
__global__ void __launch_bounds__(128) stream_kernel(const unsigned char *base, size_t stride, size_t half, int len, int S, int NS, int units_per_cta,
                                                     int hint, float *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm);
    unsigned char *ring = sm + 128;
    const int tid = threadIdx.x;
    if (tid == 0) { for (int i = 0; i < NS; ++i) mbar_init(&full[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const int per = (len + S - 1) / S;              // chunks per stream
    const int total = units_per_cta * per;          // ring slots to consume (each slot = chunk of A + chunk of B)
    auto issue = [&](int j) {
        const int u = blockIdx.x * units_per_cta + j / per, c = j % per, s = j % NS;
        const int bytes = min(S, len - c * S);
        const unsigned char *a = base + (size_t)u * stride + (size_t)c * S;
        mbar_expect(&full[s], 2 * bytes);
        bulk(ring + (size_t)s * 2 * S, a, bytes, &full[s], pol, hint);
        bulk(ring + (size_t)s * 2 * S + S, a + half, bytes, &full[s], pol, hint);
    };
    if (tid == 0) for (int j = 0; j < NS - 1 && j < total; ++j) issue(j);
    float acc = 0.f;
    for (int j = 0; j < total; ++j) {
        if (tid == 0 && j + NS - 1 < total) issue(j + NS - 1);
        mbar_wait(&full[j % NS], (j / NS) & 1);
        const float4 *p = reinterpret_cast<const float4 *>(ring + (size_t)(j % NS) * 2 * S);
        for (int i = tid; i < 2 * S / 16; i += 128) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        __syncthreads();
    }
    if (acc == 1.2345f) sink[0] = acc;
}

// plain LDG streaming of the same streams: thread = 16 B, U independent loads in flight
__global__ void __launch_bounds__(128) ldg_kernel(const unsigned char *base, size_t stride, size_t half, int len, int units_per_cta, float *sink) {
    float acc = 0.f;
    for (int j = 0; j < units_per_cta; ++j) {
        const unsigned char *a = base + (size_t)(blockIdx.x * units_per_cta + j) * stride;
        for (int o = threadIdx.x * 16; o < len; o += 128 * 16 * 4) {
            float4 v[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int oo = o + q * 128 * 16;
                v[q] = oo < len ? __ldcs(reinterpret_cast<const float4 *>(a + oo)) : make_float4(0, 0, 0, 0);
                v[4 + q] = oo < len ? __ldcs(reinterpret_cast<const float4 *>(a + half + oo)) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += v[q].x + v[q].y + v[q].z + v[q].w;
        }
    }
    if (acc == 1.2345f) sink[0] = acc;
}

int main() {
    const int n_units = 1024 * 2;                  // (row, head) pairs x 2 ... each unit = K run + V run
    const size_t stride = 769 * 256;               // bytes between consecutive (row, head) runs
    const size_t half = (size_t)n_units * stride;  // K array -> V array
    unsigned char *buf; float *sink;
    cudaMalloc(&buf, 2 * half + (1 << 20)); cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, 2 * half);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int len = 512 * 256;                     // 512 positions per run (the benchmark's mean context)
    const double bytes = 2.0 * n_units * len;
    printf("streams: %d units x 2 x %d B = %.1f MB per launch\n", n_units, len, bytes / 1e6);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int hint = 0; hint < 2; ++hint)
        for (int S : {1024, 2048, 4096, 8192, 16384})
            for (int NS : {3, 5, 9})
                for (int upc : {1, 2, 4, 8}) {
                    const size_t smem = 128 + (size_t)NS * 2 * S;
                    if (smem > 200 * 1024) continue;
                    const int grid = n_units / upc;
                    float best = 1e9;
                    for (int r = 0; r < 4; ++r) {
                        cudaEventRecord(e0);
                        stream_kernel<<<grid, 128, smem>>>(buf, stride, half, len, S, NS, upc, hint, sink);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
                    }
                    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stream_kernel, 128, smem);
                    printf("tma hint=%d S=%5d NS=%d units/cta=%d grid=%4d cta/sm=%2d inflight/sm=%4zu KB : %7.1f us  %6.0f GB/s\n", hint, S, NS, upc, grid, occ,
                           (size_t)occ * (NS - 1) * 2 * S / 1024, best * 1e3, bytes / best / 1e6);
                }
    for (int upc : {1, 2, 4}) {
        float best = 1e9;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            ldg_kernel<<<n_units / upc, 128>>>(buf, stride, half, len, upc, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
        }
        printf("ldg units/cta=%d : %7.1f us  %6.0f GB/s\n", upc, best * 1e3, bytes / best / 1e6);
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return 0;
}
