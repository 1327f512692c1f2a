// Launch interface of the iso-surface extraction kernels (mesh_kernels.cu), used by the C-ABI.
#pragma once
#include "common.cuh"

namespace sfb {

int launch_mesh_mark_edges(const float *grid, int R, float thresh, int32_t *flag, cudaStream_t s);
int launch_mesh_emit_vertices(const float *grid, int R, float thresh, const int32_t *flag, const int32_t *vid, float *verts, cudaStream_t s);
int launch_mesh_count_faces(const float *grid, int R, float thresh, int32_t *count, cudaStream_t s);
int launch_mesh_emit_faces(const float *grid, int R, float thresh, const int32_t *vid, const int32_t *foff, const float *verts, int32_t *faces,
                           cudaStream_t s);

}  // namespace sfb
