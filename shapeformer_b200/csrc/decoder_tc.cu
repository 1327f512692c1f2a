// tcgen05 form of the fused implicit decoder (placeholder until the tensor-core kernel lands: reports "unsupported"
// instead of silently running anything else).
#include "decoder_kernels.cuh"

namespace sfb {
int decoder_set_weights_tc(const float *, cudaStream_t) { return SFB200_OK; }
int launch_decoder_points_tc(const float *, const float *, int64_t, float *, int, int, int64_t, int, cudaStream_t) {
    return SFB200_E_ARG;
}
}  // namespace sfb
