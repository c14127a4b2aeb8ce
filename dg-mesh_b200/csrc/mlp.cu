// mlp.cu -- deformation / appearance MLPs on the tcgen05 tensor cores.
// (first part: the GEMM building block and its C-ABI test entry; the network follows below)
#include "mlp_gemm.cuh"
#include "mlp_kernels.h"

namespace dgm {

template <int BN>
static cudaError_t launch_gemm_bn(const GemmArgs& g, cudaStream_t s) {
  static bool attr = false;
  constexpr int smem = GEMM_STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2);
  if (!attr) {
    cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  const int splits = (g.K + g.k_split - 1) / g.k_split;
  dim3 grid((g.M + GEMM_BM - 1) / GEMM_BM, (g.N + BN - 1) / BN, splits);
  gemm_tn_kernel<BN><<<grid, 128, smem, s>>>(g);
  return cudaGetLastError();
}

cudaError_t launch_gemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return cudaSuccess;
  if (g.N <= 32) return launch_gemm_bn<32>(g, s);
  if (g.N <= 64) return launch_gemm_bn<64>(g, s);
  if (g.N <= 128) return launch_gemm_bn<128>(g, s);
  return launch_gemm_bn<256>(g, s);
}

}  // namespace dgm
