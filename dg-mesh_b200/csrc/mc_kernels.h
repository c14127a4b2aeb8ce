// mc_kernels.h -- host-side interface of mc.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dgm {
struct McWS {
  unsigned long long *blk_counts, *blk_offsets;  // per brick of 8 x 32 nodes: (vertices | triangles << 32), exclusive scan
  int32_t* totals;                                // device copy of {V, F} for the resolve pass
  uint32_t* vid;                                  // [n] sparse: first vertex index | owned-edge mask << 29
  uint32_t* vsrc;                                 // [V] node * 3 + axis of every vertex (backward pass)
  char* cub_temp;
  size_t cub_bytes;
  static McWS from(char* base, int G, size_t* bytes = nullptr);
};
size_t mc_num_blocks(int G);
cudaError_t launch_mc_count(int G, const float* phi, float iso, void* ws, int32_t* totals, int32_t* totals_host,
                            cudaEvent_t ev, cudaStream_t s);
cudaError_t launch_mc_emit(int G, const float* phi, float iso, void* ws, float* verts, long long V_cap,
                           int32_t* faces, long long F_cap, cudaStream_t s);
cudaError_t launch_mc_backward(int G, int V, const float* phi, float iso, void* ws, const float* dverts, float* dphi,
                               cudaStream_t s);
}  // namespace dgm
