// Launch interface of the conv-prologue kernels (conv_tc.cu), used by the C-ABI.
#pragma once
#include "common.cuh"

namespace sfb {

int launch_conv3d_tc(const float *in, const float *in_lo, const float *w, const float *w_lo, const float *bias, float *out, double *stats,
                     int B, int Z, int Y, int X, int Cin, int Cout, int taps, int relu, cudaStream_t stream);
int launch_conv_prep(const float *src0, int C0, int sh0, const double *st0, double n0, const float *src1, int C1, int sh1, const double *st1,
                     double n1, const float *gamma, const float *beta, int groups, float *dst, float *dst_lo, int B, int Z, int Y, int X,
                     cudaStream_t stream);
int launch_pool_stats(const float *src, float *dst, double *stats, int B, int Zo, int Yo, int Xo, int C, int win, cudaStream_t stream);
int launch_gather_codes_cl(const int64_t *idx, const float *codebook, float *out, double *stats, int B, int cells, int C, int n_codes,
                           cudaStream_t stream);

}  // namespace sfb
