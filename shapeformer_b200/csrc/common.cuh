// Shared device/host helpers for libsfb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/sfb200.h"

namespace sfb {

void set_cuda_error(cudaError_t e, const char *where);
void count_launches(long long n);

// Returns SFB200_OK or SFB200_E_CUDA after recording the message.
static inline int check_launch(const char *where) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_cuda_error(e, where);
        return SFB200_E_CUDA;
    }
    count_launches(1);
    return SFB200_OK;
}

#define SFB_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) {                             \
            ::sfb::set_cuda_error(_e, #expr);                \
            return SFB200_E_CUDA;                            \
        }                                                    \
    } while (0)

#define SFB_TRY(expr)                 \
    do {                              \
        int _r = (expr);              \
        if (_r != SFB200_OK) return _r; \
    } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------------
// Every kernel of the AR step starts with pdl_trigger() + pdl_wait(): the NEXT kernel of the stream may be scheduled as soon
// as all CTAs of this one have started (its launch latency and prologue overlap this kernel's execution), and it blocks in
// pdl_wait() until this kernel has completed and its writes are visible.  Both are no-ops without the launch attribute.
// NOTE: activations produced by the previous kernel must be read with coherent loads — no `const __restrict__` / __ldg on
// them (an early-resident CTA could otherwise hit stale lines through the non-coherent path); only weights keep __restrict__.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
bool pdl_enabled();   // on by default; SFB200_PDL=0 in the environment disables the launch attribute

// Launch with optional cluster dimensions and the PDL attribute.
template <typename... KArgs, typename... Args>
static inline int launch_ex(const char *what, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            dim3 cluster, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster.x * cluster.y * cluster.z > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster.x;
        attr[na].val.clusterDim.y = cluster.y;
        attr[na].val.clusterDim.z = cluster.z;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_cuda_error(e, what);
        return SFB200_E_CUDA;
    }
    count_launches(1);
    return SFB200_OK;
}

// Per-device one-time setup (cudaFuncSetAttribute is a per-device property): returns true the first time it is called with
// this `mask` on the CURRENT device.  A benign race between host threads only repeats an idempotent setup.
static inline bool first_use_on_device(unsigned long long &mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
    if ((mask >> dev) & 1ull) return false;
    mask |= 1ull << dev;
    return true;
}

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// streaming (evict-first) 128-bit load for data touched once per launch (KV cache, weights)
__device__ __forceinline__ float4 ld4_stream(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// lo part of the two-term TF32 split whose hi part is the tensor core's own truncation of v (it reads the upper 19 bits):
// lo = rna_tf32(v - trunc_tf32(v)); the subtraction is exact.
__device__ __forceinline__ float tf32_lo(float v) {
    const float r = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    uint32_t o;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(o) : "f"(r));
    return __uint_as_float(o);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace sfb
