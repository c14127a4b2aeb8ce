// loss_kernels.h -- host-side interface of loss.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace dgm {
size_t loss_workspace_bytes(int H, int W);
cudaError_t launch_loss_forward(int H, int W, const float* img, const float* gt, float lam, int mode, float* out3,
                                void* ws, cudaStream_t s);
cudaError_t launch_loss_backward(int H, int W, const float* img, const float* gt, float lam, int mode,
                                 const float* upstream, float* dL_dimg, void* ws, cudaStream_t s);
size_t laplacian_ws_bytes(int V);
cudaError_t launch_laplacian_forward(int V, int F, const float* verts, const int* tri, float* out, void* ws,
                                     cudaStream_t s);
cudaError_t launch_laplacian_backward(int V, int F, const int* tri, const float* dL_dloss, float* dverts, void* ws,
                                      cudaStream_t s);
}  // namespace dgm
