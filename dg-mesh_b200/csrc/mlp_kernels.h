// mlp_kernels.h -- host-side interface of mlp.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dgm {
struct GemmArgs;
cudaError_t launch_gemm(const GemmArgs& g, cudaStream_t s);
}  // namespace dgm
