// mlp_kernels.h -- host-side interface of mlp.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/dgmesh_b200.h"

namespace dgm {
size_t gemm_test_ws_bytes(int M, int N, int K);
cudaError_t launch_gemm_test(int M, int N, int K, const void* A, int lda, const void* B, int ldb, const float* bias,
                             int relu, float* C, int ldc, void* ws, cudaStream_t s);
cudaError_t launch_gemm_tn_test(int P, int Mf, int Nf, const void* X, int ldx, const void* Y, int ldy, float* C,
                                int ldc, int transpose, void* ws, cudaStream_t s);
cudaError_t launch_mlp_forward(const DglNet& n, int P, const float* x, const float* t, float* out, int train,
                               void* ws, cudaStream_t s);
cudaError_t launch_mlp_backward(const DglNet& n, int P, const float* x, const float* out, const float* g_out,
                                void* ws, const DglGrads& gr, float* dx, cudaStream_t s);
size_t mlp_workspace_bytes(int P, int train);
void mlp_pack_sizes(size_t* w_bytes, size_t* b_bytes, size_t* g_bytes);
cudaError_t launch_mlp_pack(const DglRaw& r, void* wbuf, float* bbuf, DglNet* net, cudaStream_t s);
void mlp_grad_pointers(float* gbuf, DglGrads* g);
cudaError_t launch_mlp_unpack_grads(const DglRaw& r, const float* gbuf, const DglRawGrads& o, cudaStream_t s);
}  // namespace dgm
