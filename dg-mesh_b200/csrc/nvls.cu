// nvls.cu -- the data-parallel exchange as our own kernel over NVLink / NVSwitch.
//
// The one collective of the path is an all-reduce (sum) of ONE flat fp32 gradient buffer
// [canonical-Gaussian gradients || MLP gradients] (23.6 MB at C2, 60 MB at C4).  At these sizes NCCL's
// all-reduce is latency-bound (measured 0.12 ms for 23.6 MB on 8 B200s = 340 GB/s bus bandwidth, a fifth of
// what the switch moves).  With the buffer in symmetric memory mapped through an NVSwitch MULTICAST object,
// the reduction is done by the switch itself:
//     rank r owns slice r of the buffer:  v = multimem.ld_reduce.add(slice r)    -- the switch fetches the N
//                                                                                   replicas, adds them, returns one
//                                         multimem.st(slice r, v)                -- the switch writes v to all N
// so every GPU receives the buffer once and sends it once (2 x bytes / N per direction and rank), in ONE
// kernel bracketed by two flag barriers across the ranks (signal pads in the peers' symmetric memory).
// The caller (dg-mesh_b200/dp.py) allocates the buffer with torch's symmetric-memory allocator and passes
// the multicast address and the table of signal-pad addresses; nothing here allocates or synchronises.
#include "common.cuh"
#include "nvls_kernels.h"

namespace dgm {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All ranks' CTA `blockIdx.x` meet: thread t < world raises flag (block, my rank) in rank t's pad and waits
// for flag (block, t) in its own.  Flags carry the epoch, so pads are never reset.
__device__ __forceinline__ void rank_barrier(uint32_t* const* __restrict__ pads, int rank, int world, uint32_t epoch) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    const int t = threadIdx.x;
    st_release_sys(pads[t] + (size_t)blockIdx.x * world + rank, epoch);
    const uint32_t* mine = pads[rank] + (size_t)blockIdx.x * world + t;
    while (ld_acquire_sys(mine) != epoch) {
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) nvls_allreduce_kernel(float* __restrict__ mc, size_t n_vec4,
                                                             uint32_t* const* __restrict__ pads, int rank, int world,
                                                             uint32_t epoch, float scale) {
  // everyone's local contributions are complete (stream order on each rank) and visible (system fence)
  rank_barrier(pads, rank, world, epoch);
  const size_t per = (n_vec4 + world - 1) / world;
  const size_t lo = per * rank, hi = min(n_vec4, lo + per);
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
    float4 v;
    float* p = mc + 4 * i;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    v.x *= scale, v.y *= scale, v.z *= scale, v.w *= scale;
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
  }
  // every slice has been written to every replica before anyone reads the result
  rank_barrier(pads, rank, world, epoch + 1);
}

cudaError_t launch_nvls_allreduce(float* multicast_ptr, size_t n_floats, uint32_t* const* signal_pads_dev, int rank,
                                  int world, uint32_t epoch, float scale, int blocks, cudaStream_t s) {
  nvls_allreduce_kernel<<<blocks, 512, 0, s>>>(multicast_ptr, n_floats / 4, signal_pads_dev, rank, world, epoch, scale);
  return cudaGetLastError();
}

}  // namespace dgm
