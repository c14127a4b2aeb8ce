// nvls_kernels.h -- host-side interface of nvls.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define NVLS_PAD_SKIP 128  // 32-bit words at the start of every signal pad left to the allocator's own barriers

namespace dgm {
cudaError_t launch_nvls_allreduce(float* multicast_ptr, size_t n_floats, uint32_t* const* signal_pads_dev, int rank,
                                  int world, uint32_t epoch, float scale, int blocks, cudaStream_t s);
}  // namespace dgm
