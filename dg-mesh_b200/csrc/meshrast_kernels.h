// meshrast_kernels.h -- host-side launch interface of meshrast.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace dgm {
cudaError_t launch_mr_rasterize(int V, int F, int W, int H, const float* pos, const int* tri, void* zbuf, float* rast,
                                cudaStream_t s);
cudaError_t launch_mr_rasterize_bwd(int W, int H, const float* rast, const int* tri, const float* pos,
                                    const float* grast, float* gpos, cudaStream_t s);
cudaError_t launch_mr_interpolate(int W, int H, int C, const float* attr, const float* rast, const int* tri,
                                  float* out, cudaStream_t s);
cudaError_t launch_mr_interpolate_bwd(int W, int H, int C, const float* attr, const float* rast, const int* tri,
                                      const float* gout, float* gattr, float* grast, cudaStream_t s);
cudaError_t launch_mr_antialias(int W, int H, int C, const float* color, const float* rast, const float* pos,
                                const int* tri, const int* opp, float* out, cudaStream_t s);
cudaError_t launch_mr_antialias_bwd(int W, int H, int C, const float* color, const float* rast, const float* pos,
                                    const int* tri, const int* opp, const float* gout, float* gcolor, float* gpos,
                                    cudaStream_t s);
}  // namespace dgm
