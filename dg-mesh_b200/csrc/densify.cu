// densify.cu -- adaptive density control of the canonical Gaussians on the device.
//
// Reference: GaussianModelDPSRDynamicAnchor.densify_and_prune
// (dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:469-551) =
//     densify_and_clone -> densify_and_split (N = 2) -> prune
// each of which rebuilds the seven parameter tensors and both Adam moments with torch.cat / boolean
// masking (cat_tensors_to_optimizer :424-449, _prune_optimizer :383-401): ~120 kernels, ~60
// allocations and three full copies of the model per call.  Here the three steps are composed into
// ONE plan and ONE gather:
//   plan_kernel    per source Gaussian: clone / split selection and the prune verdict of every copy
//                  it would leave behind -> four 0/1 arrays
//   (4 x cub::DeviceScan::ExclusiveSum, count_kernel -> counts on the device; the caller reads them:
//    the one host read, needed to size the new tensors -- the reference reads sizes at every step)
//   stds_kernel    standard deviations of the split samples in the order the reference draws them, so
//                  that the caller's torch.normal() consumes the generator exactly as the reference
//   apply_kernel   every surviving row of the 7 x 3 tensors is written once at its final position:
//                  [kept originals | clones | split children (first copies) | (second copies)]
// The final order, every copied value and the Adam state (zeros for new rows) equal the reference's;
// the two computed quantities (child position R s + mu, child log-scale) follow its formulas.
// HBM-bound: algorithmic bytes = 2 x 744 B per surviving Gaussian (62 floats x 3 tensors, read + write).
#include <cub/cub.cuh>

#include "common.cuh"
#include "densify_kernels.h"

namespace dgm {

struct DensifyWS {
  int32_t *f_keep, *f_clone, *f_child, *f_split;  // flags, then (in place) their exclusive prefix sums
  uint8_t* bits;                                   // bit0 keep-orig, bit1 clone, bit2 children, bit3 split-selected
  int32_t* counts;                                 // [4] n_keep, n_clone, n_split_selected, n_child (per copy)
  void* cub_tmp;
  size_t cub_bytes;
  static DensifyWS from(char* base, size_t P, size_t cub_bytes, size_t* bytes = nullptr) {
    char* p = base;
    DensifyWS w;
    w.f_keep = carve<int32_t>(p, P);
    w.f_clone = carve<int32_t>(p, P);
    w.f_child = carve<int32_t>(p, P);
    w.f_split = carve<int32_t>(p, P);
    w.bits = carve<uint8_t>(p, P);
    w.counts = carve<int32_t>(p, 4);
    w.cub_tmp = carve<char>(p, cub_bytes);
    w.cub_bytes = cub_bytes;
    if (bytes) *bytes = size_t(p - base) + 128;
    return w;
  }
};

static size_t densify_cub_bytes(int P) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, P);
  return b;
}

size_t densify_ws_bytes(int P) {
  size_t bytes;
  DensifyWS::from(nullptr, (size_t)P, densify_cub_bytes(P), &bytes);
  return bytes;
}

// torch.sigmoid / torch.exp on fp32 CUDA tensors evaluate 1 / (1 + exp(-x)) and exp(x) with the
// full-precision single-precision library functions; the same calls here give the same verdicts.
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) densify_plan_kernel(
    int P, const float* __restrict__ grad_accum, const float* __restrict__ denom,
    const float* __restrict__ scaling_raw, const float* __restrict__ opacity_raw, float max_grad, float min_opacity,
    float extent, float percent_dense, int size_prune, float max_screen_size, DensifyWS w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  // grads = xyz_gradient_accum / denom; grads[isnan] = 0   (:541-542)
  float g = grad_accum[i] / denom[i];
  if (isnan(g)) g = 0.0f;
  const float s0 = expf(scaling_raw[3 * i]), s1 = expf(scaling_raw[3 * i + 1]), s2 = expf(scaling_raw[3 * i + 2]);
  const float smax = fmaxf(s0, fmaxf(s1, s2));
  const float thr = percent_dense * extent;
  // densify_and_clone :487-491 (norm of a 1-vector = |g|), densify_and_split :455-461
  const bool clone = (fabsf(g) >= max_grad) && (smax <= thr);
  const bool split = (g >= max_grad) && (smax > thr);
  // prune :524-533.  densification_postfix has reset max_radii2D to zero (:449), so the screen-size test
  // reads zeros: it fires only for a negative max_screen_size.
  const bool low_op = sigmoid_ref(opacity_raw[i]) < min_opacity;
  const bool vs = size_prune && (0.0f > max_screen_size);
  const bool pruned = low_op || vs || (size_prune && smax > 0.1f * extent);
  // children: _scaling = log(scale / (0.8 * 2)) (:467), get_scaling = exp of that
  const float c0 = expf(logf(s0 / 1.6f)), c1 = expf(logf(s1 / 1.6f)), c2 = expf(logf(s2 / 1.6f));
  const bool child_pruned = low_op || vs || (size_prune && fmaxf(c0, fmaxf(c1, c2)) > 0.1f * extent);
  const bool keep = !split && !pruned;            // a split-selected original is removed (:478-480)
  const bool keep_clone = clone && !pruned;       // the clone is an exact copy: same verdict
  const bool keep_child = split && !child_pruned;
  w.f_keep[i] = keep;
  w.f_clone[i] = keep_clone;
  w.f_child[i] = keep_child;
  w.f_split[i] = split;
  w.bits[i] = (uint8_t)((keep ? 1 : 0) | (keep_clone ? 2 : 0) | (keep_child ? 4 : 0) | (split ? 8 : 0));
}

__global__ void densify_count_kernel(int P, DensifyWS w, int32_t* __restrict__ counts_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const uint8_t b = w.bits[P - 1];
    const int32_t c[4] = {w.f_keep[P - 1] + ((b & 1) ? 1 : 0), w.f_clone[P - 1] + ((b & 2) ? 1 : 0),
                          w.f_split[P - 1] + ((b & 8) ? 1 : 0), w.f_child[P - 1] + ((b & 4) ? 1 : 0)};
    for (int k = 0; k < 4; ++k) {
      w.counts[k] = c[k];
      counts_out[k] = c[k];
    }
  }
}

cudaError_t launch_densify_plan(int P, const float* grad_accum, const float* denom, const float* scaling_raw,
                                const float* opacity_raw, float max_grad, float min_opacity, float extent,
                                float percent_dense, int size_prune, float max_screen_size, void* ws,
                                int32_t* counts, cudaStream_t s) {
  const size_t cb = densify_cub_bytes(P);
  DensifyWS w = DensifyWS::from((char*)ws, (size_t)P, cb);
  densify_plan_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, grad_accum, denom, scaling_raw, opacity_raw, max_grad,
                                                      min_opacity, extent, percent_dense, size_prune, max_screen_size,
                                                      w);
  size_t b = cb;
  cub::DeviceScan::ExclusiveSum(w.cub_tmp, b, w.f_keep, w.f_keep, P, s);
  cub::DeviceScan::ExclusiveSum(w.cub_tmp, b, w.f_clone, w.f_clone, P, s);
  cub::DeviceScan::ExclusiveSum(w.cub_tmp, b, w.f_child, w.f_child, P, s);
  cub::DeviceScan::ExclusiveSum(w.cub_tmp, b, w.f_split, w.f_split, P, s);
  densify_count_kernel<<<1, 32, 0, s>>>(P, w, counts);
  return cudaGetLastError();
}

// stds = get_scaling[selected].repeat(2, 1)   (:463): row r and row n_split + r hold exp(_scaling) of the
// r-th split-selected Gaussian
__global__ void __launch_bounds__(256) densify_stds_kernel(int P, const float* __restrict__ scaling_raw, DensifyWS w,
                                                           float* __restrict__ stds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || !(w.bits[i] & 8)) return;
  const int r = w.f_split[i], n = w.counts[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float s = expf(scaling_raw[3 * i + k]);
    stds[3 * (size_t)r + k] = s;
    stds[3 * ((size_t)n + r) + k] = s;
  }
}

cudaError_t launch_densify_stds(int P, const float* scaling_raw, void* ws, float* stds, cudaStream_t s) {
  DensifyWS w = DensifyWS::from((char*)ws, (size_t)P, densify_cub_bytes(P));
  densify_stds_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, scaling_raw, w, stds);
  return cudaGetLastError();
}

// One CTA of 64 threads per source Gaussian; thread c owns one of the 62 parameter columns
// (xyz 3 | f_dc 3 | f_rest 45 | opacity 1 | scaling 3 | rotation 4 | normal 3) of all three variants
// (parameter, exp_avg, exp_avg_sq).
__global__ void __launch_bounds__(64) densify_apply_kernel(int P, DensifyTables t, const float* __restrict__ samples,
                                                           DensifyWS w) {
  const int i = blockIdx.x;
  const uint8_t b = w.bits[i];
  if (!(b & 7)) return;
  const int c = threadIdx.x;
  // column -> field
  int f = 0, col = c;
  while (f < t.n_fields && col >= t.f[f].width) {
    col -= t.f[f].width;
    ++f;
  }
  const bool active = f < t.n_fields;
  const int n_keep = w.counts[0], n_clone = w.counts[1], n_split = w.counts[2], n_child = w.counts[3];
  const long long d_keep = w.f_keep[i];
  const long long d_clone = (long long)n_keep + w.f_clone[i];
  const long long d_c0 = (long long)n_keep + n_clone + w.f_child[i];
  const long long d_c1 = d_c0 + n_child;
  float v = 0.f, m1 = 0.f, m2 = 0.f;
  int wd = 0;
  if (active) {
    const DensifyField& F = t.f[f];
    wd = F.width;
    v = F.src[(size_t)i * wd + col];
    if (F.m1_src) {
      m1 = F.m1_src[(size_t)i * wd + col];
      m2 = F.m2_src[(size_t)i * wd + col];
    }
    if (b & 1) {
      F.dst[d_keep * wd + col] = v;
      if (F.m1_dst) F.m1_dst[d_keep * wd + col] = m1, F.m2_dst[d_keep * wd + col] = m2;
    }
    if (b & 2) {  // clone: the parameters are copied, its Adam moments start at zero (:432-436)
      F.dst[d_clone * wd + col] = v;
      if (F.m1_dst) F.m1_dst[d_clone * wd + col] = 0.f, F.m2_dst[d_clone * wd + col] = 0.f;
    }
  }
  if (b & 4) {
    // children (:462-472): xyz = R(q / |q|) s + mu with s ~ N(0, scale); _scaling = log(scale / 1.6);
    // everything else copied; moments zero
    float c0v = v, c1v = v;
    if (active && t.f[f].role == DENSIFY_ROLE_XYZ) {
      const float* q = t.rotation_raw + 4 * (size_t)i;
      const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);  // build_rotation :general_utils
      const float r = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
      float R[3];
      if (col == 0) R[0] = 1 - 2 * (y * y + z * z), R[1] = 2 * (x * y - r * z), R[2] = 2 * (x * z + r * y);
      else if (col == 1) R[0] = 2 * (x * y + r * z), R[1] = 1 - 2 * (x * x + z * z), R[2] = 2 * (y * z - r * x);
      else R[0] = 2 * (x * z - r * y), R[1] = 2 * (y * z + r * x), R[2] = 1 - 2 * (x * x + y * y);
      const int rk = w.f_split[i];
      const float* sa = samples + 3 * (size_t)rk;
      const float* sb = samples + 3 * ((size_t)n_split + rk);
      c0v = (R[0] * sa[0] + R[1] * sa[1] + R[2] * sa[2]) + v;
      c1v = (R[0] * sb[0] + R[1] * sb[1] + R[2] * sb[2]) + v;
    } else if (active && t.f[f].role == DENSIFY_ROLE_SCALING) {
      c0v = c1v = logf(expf(v) / 1.6f);
    }
    if (active) {
      const DensifyField& F = t.f[f];
      F.dst[d_c0 * wd + col] = c0v;
      F.dst[d_c1 * wd + col] = c1v;
      if (F.m1_dst) {
        F.m1_dst[d_c0 * wd + col] = 0.f, F.m2_dst[d_c0 * wd + col] = 0.f;
        F.m1_dst[d_c1 * wd + col] = 0.f, F.m2_dst[d_c1 * wd + col] = 0.f;
      }
    }
  }
}

cudaError_t launch_densify_apply(int P, const DensifyTables& t, const float* samples, void* ws, cudaStream_t s) {
  DensifyWS w = DensifyWS::from((char*)ws, (size_t)P, densify_cub_bytes(P));
  densify_apply_kernel<<<P, 64, 0, s>>>(P, t, samples, w);
  return cudaGetLastError();
}

}  // namespace dgm
