// mc_kernels.h -- host-side interface of mc.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dgm {
struct McWS {
  unsigned long long *counts, *offsets;
  uint16_t* info;
  char* cub_temp;
  size_t cub_bytes;
  static McWS from(char* base, int G, size_t* bytes = nullptr);
};
cudaError_t launch_mc_count(int G, const float* phi, float iso, void* ws, int32_t* totals, cudaStream_t s);
cudaError_t launch_mc_emit(int G, const float* phi, float iso, void* ws, float* verts, int32_t* faces, cudaStream_t s);
cudaError_t launch_mc_backward(int G, const float* phi, float iso, void* ws, const float* dverts, float* dphi,
                               cudaStream_t s);
}  // namespace dgm
