// knn_kernels.h -- host-side launch interface of knn.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dgm {
struct KnnWS {
  uint32_t *bounds, *codes, *codes_sorted, *idx, *order;
  float *sorted, *boxes;
  char* cub_temp;
  size_t cub_bytes;
  static KnnWS from(char* base, size_t P, size_t cub_bytes, size_t* bytes = nullptr);
};
size_t knn_cub_bytes(int P);
cudaError_t launch_nearest(int Q, const float* q, int R, const float* r, float* dist2, long long* index,
                           cudaStream_t s);
cudaError_t launch_knn(int P, const float* points, float* mean_dist2, void* ws, cudaStream_t s);
}  // namespace dgm
