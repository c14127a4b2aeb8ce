// densify_kernels.h -- host-side launch interface of densify.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/dgmesh_b200.h"

namespace dgm {

#define DENSIFY_ROLE_PLAIN 0
#define DENSIFY_ROLE_XYZ 1
#define DENSIFY_ROLE_SCALING 2

struct DensifyField {
  const float *src, *m1_src, *m2_src;  // parameter [P, width], Adam exp_avg / exp_avg_sq (null: no state yet)
  float *dst, *m1_dst, *m2_dst;        // [n_out, width]
  int width, role;
};
struct DensifyTables {
  DensifyField f[DGD_MAX_FIELDS];
  int n_fields;
  const float* rotation_raw;  // [P, 4]
};

size_t densify_ws_bytes(int P);
cudaError_t launch_densify_plan(int P, const float* grad_accum, const float* denom, const float* scaling_raw,
                                const float* opacity_raw, float max_grad, float min_opacity, float extent,
                                float percent_dense, int size_prune, float max_screen_size, void* ws, int32_t* counts,
                                cudaStream_t s);
cudaError_t launch_densify_stds(int P, const float* scaling_raw, void* ws, float* stds, cudaStream_t s);
cudaError_t launch_densify_apply(int P, const DensifyTables& t, const float* samples, void* ws, cudaStream_t s);

}  // namespace dgm
