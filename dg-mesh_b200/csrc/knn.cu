// knn.cu -- mean squared distance to the 3 nearest neighbours of every point
// (simple-knn `distCUDA2`: dgmesh/submodules/simple-knn/simple_knn.cu:185-221, spatial.cu:15-26).
//
// Same result as the reference (exact 3-NN; only the fp32 rounding of the final mean can
// differ in the last bit), different organisation:
//   * no host round trips: the reference reduces min / max with two blocking D2H copies
//     (simple_knn.cu:196-200); here the bounds stay on the device;
//   * boxes of 256 Morton-ordered points (reference: 1024) and one CTA per box: the CTA's
//     points are Morton neighbours, so they need (almost) the same candidate boxes; a
//     candidate box is staged ONCE in shared memory for the whole CTA (block vote) instead
//     of being re-read from global memory by every thread (simple_knn.cu:166-180);
//   * candidate boxes are visited outwards from the CTA's own box in Morton order, so the
//     pruning radius shrinks early.
#include <cfloat>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "knn_kernels.h"

namespace dgm {

#define KNN_BOX 256

// order-preserving float <-> uint encoding (for atomicMin/Max on signed floats)
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

// bounds[0..2] = max of ~ord(x) (i.e. the minimum), bounds[3..5] = max of ord(x); zero-initialised
__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ pts,
                                                         uint32_t* __restrict__ bounds) {
  uint32_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t o = f2ord(pts[3 * i + c]);
      lo[c] = max(lo[c], ~o);
      hi[c] = max(hi[c], o);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      lo[c] = max(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], off));
      hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], off));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMax(&bounds[c], lo[c]);
      atomicMax(&bounds[3 + c], hi[c]);
    }
  }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {  // 10 bits -> every third bit
  x = (x | (x << 16)) & 0x030000FF;
  x = (x | (x << 8)) & 0x0300F00F;
  x = (x | (x << 4)) & 0x030C30C3;
  x = (x | (x << 2)) & 0x09249249;
  return x;
}

// 30-bit Morton code of the point inside the bounding box (simple_knn.cu:54-70)
__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float* __restrict__ pts,
                                                         const uint32_t* __restrict__ bounds,
                                                         uint32_t* __restrict__ codes, uint32_t* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint32_t code = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float mn = ord2f(~bounds[c]), mx = ord2f(bounds[3 + c]);
    const float ext = mx - mn;
    const float u = ext > 0.f ? (pts[3 * i + c] - mn) / ext : 0.f;
    const uint32_t q = (uint32_t)(u * 1023.0f);
    code |= spread10(min(q, 1023u)) << c;
  }
  codes[i] = code;
  idx[i] = (uint32_t)i;
}

struct Aabb {
  float mn[3], mx[3];
};

// gather the points into Morton order and compute the AABB of every run of 256 (simple_knn.cu:78-117)
__global__ void __launch_bounds__(KNN_BOX) knn_boxes_kernel(int P, const float* __restrict__ pts,
                                                            const uint32_t* __restrict__ order,
                                                            float* __restrict__ sorted, Aabb* __restrict__ boxes) {
  __shared__ float s_mn[3][KNN_BOX / 32], s_mx[3][KNN_BOX / 32];
  const int i = blockIdx.x * KNN_BOX + threadIdx.x;
  float p[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, q[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < P) {
    const uint32_t src = order[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = pts[3 * (size_t)src + c];
      sorted[3 * (size_t)i + c] = v;
      p[c] = q[c] = v;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      p[c] = fminf(p[c], __shfl_xor_sync(0xffffffffu, p[c], off));
      q[c] = fmaxf(q[c], __shfl_xor_sync(0xffffffffu, q[c], off));
    }
    if ((threadIdx.x & 31) == 0) {
      s_mn[c][threadIdx.x >> 5] = p[c];
      s_mx[c][threadIdx.x >> 5] = q[c];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Aabb b;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      b.mn[c] = s_mn[c][0];
      b.mx[c] = s_mx[c][0];
      for (int w = 1; w < KNN_BOX / 32; ++w) {
        b.mn[c] = fminf(b.mn[c], s_mn[c][w]);
        b.mx[c] = fmaxf(b.mx[c], s_mx[c][w]);
      }
    }
    boxes[blockIdx.x] = b;
  }
}

// squared distance from a point to a box (0 inside), simple_knn.cu:119-129
__device__ __forceinline__ float box_dist2(const Aabb& b, const float x, const float y, const float z) {
  float dx = 0.f, dy = 0.f, dz = 0.f;
  if (x < b.mn[0] || x > b.mx[0]) dx = fminf(fabsf(x - b.mn[0]), fabsf(x - b.mx[0]));
  if (y < b.mn[1] || y > b.mx[1]) dy = fminf(fabsf(y - b.mn[1]), fabsf(y - b.mx[1]));
  if (z < b.mn[2] || z > b.mx[2]) dz = fminf(fabsf(z - b.mn[2]), fabsf(z - b.mx[2]));
  return dx * dx + dy * dy + dz * dz;
}

// keep the three smallest distances sorted (simple_knn.cu:131-145)
__device__ __forceinline__ void push3(float (&best)[3], float d) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (best[j] > d) {
      const float t = best[j];
      best[j] = d;
      d = t;
    }
  }
}

__global__ void __launch_bounds__(KNN_BOX) knn_kernel(int P, int nboxes, const float* __restrict__ sorted,
                                                      const uint32_t* __restrict__ order,
                                                      const Aabb* __restrict__ boxes, float* __restrict__ out) {
  __shared__ float s_pts[KNN_BOX * 3];
  const int own = blockIdx.x;
  const int i = own * KNN_BOX + threadIdx.x;
  const bool valid = i < P;
  float x = 0.f, y = 0.f, z = 0.f;
  if (valid) {
    x = sorted[3 * (size_t)i];
    y = sorted[3 * (size_t)i + 1];
    z = sorted[3 * (size_t)i + 2];
  }
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  // visit order: own box, then own+1, own-1, own+2, ...
  for (int step = 0; step < 2 * nboxes; ++step) {
    const int k = (step + 1) >> 1;
    const int b = (step & 1) ? own + k : own - k;
    if (b < 0 || b >= nboxes) continue;  // uniform per CTA
    bool need = false;
    if (valid) need = (b == own) || !(box_dist2(boxes[b], x, y, z) > best[2]);
    if (!__syncthreads_or(need)) continue;
    const int base = b * KNN_BOX;
    const int cnt = min(KNN_BOX, P - base);
    for (int t = threadIdx.x; t < cnt * 3; t += KNN_BOX) s_pts[t] = sorted[3 * (size_t)base + t];
    __syncthreads();
    if (need) {
      for (int j = 0; j < cnt; ++j) {
        if (base + j == i) continue;  // only the point itself is excluded (duplicates count)
        const float dx = s_pts[3 * j] - x, dy = s_pts[3 * j + 1] - y, dz = s_pts[3 * j + 2] - z;
        push3(best, dx * dx + dy * dy + dz * dz);
      }
    }
    __syncthreads();
  }
  if (valid) out[order[i]] = (best[0] + best[1] + best[2]) / 3.0f;
}

size_t knn_cub_bytes(int P) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, P, 0, 30);
  return bytes;
}

KnnWS KnnWS::from(char* base, size_t P, size_t cub_bytes, size_t* bytes) {
  char* p = base;
  KnnWS w;
  const size_t nb = (P + KNN_BOX - 1) / KNN_BOX;
  w.bounds = carve<uint32_t>(p, 8);
  w.codes = carve<uint32_t>(p, P);
  w.codes_sorted = carve<uint32_t>(p, P);
  w.idx = carve<uint32_t>(p, P);
  w.order = carve<uint32_t>(p, P);
  w.sorted = carve<float>(p, 3 * P);
  w.boxes = carve<float>(p, 6 * nb);
  w.cub_temp = carve<char>(p, cub_bytes);
  w.cub_bytes = cub_bytes;
  if (bytes) *bytes = size_t(p - base) + 128;
  return w;
}

// Cross-set nearest neighbour (K = 1) for anchor_mesh: every query against every reference point, the references
// streamed through shared memory in tiles; squared distance accumulated x, y, z with fused multiply-adds (the
// operation order of pytorch3d's knn kernel), ties resolved to the lowest reference index.  Off the per-iteration
// path (every anchor_interval iterations): ~2e10 pair evaluations at 200k x 100k, a few ms.
#define NN_TILE 1024
__global__ void __launch_bounds__(256) nearest_kernel(int Q, const float* __restrict__ q, int R,
                                                      const float* __restrict__ r, float* __restrict__ dist2,
                                                      long long* __restrict__ index) {
  __shared__ float4 s_r[NN_TILE];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < Q;
  const float x = live ? q[3 * (size_t)i] : 0.f, y = live ? q[3 * (size_t)i + 1] : 0.f, z = live ? q[3 * (size_t)i + 2] : 0.f;
  float best = __int_as_float(0x7f800000);
  int bi = 0;
  for (int base = 0; base < R; base += NN_TILE) {
    const int n = min(NN_TILE, R - base);
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const float* p = r + 3 * (size_t)(base + j);
      s_r[j] = make_float4(p[0], p[1], p[2], 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const float4 c = s_r[j];
      const float dx = x - c.x, dy = y - c.y, dz = z - c.z;
      const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
      if (d < best) {
        best = d;
        bi = base + j;
      }
    }
    __syncthreads();
  }
  if (live) {
    dist2[i] = best;
    index[i] = bi;
  }
}

cudaError_t launch_nearest(int Q, const float* q, int R, const float* r, float* dist2, long long* index,
                           cudaStream_t s) {
  if (Q == 0) return cudaSuccess;
  nearest_kernel<<<(Q + 255) / 256, 256, 0, s>>>(Q, q, R, r, dist2, index);
  return cudaGetLastError();
}

cudaError_t launch_knn(int P, const float* points, float* mean_dist2, void* ws, cudaStream_t s) {
  if (P == 0) return cudaSuccess;
  const size_t cub_bytes = knn_cub_bytes(P);
  KnnWS w = KnnWS::from((char*)ws, P, cub_bytes);
  const int nboxes = (P + KNN_BOX - 1) / KNN_BOX;
  cudaMemsetAsync(w.bounds, 0, 8 * sizeof(uint32_t), s);
  knn_bounds_kernel<<<min(592, (P + 255) / 256), 256, 0, s>>>(P, points, w.bounds);
  knn_morton_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, w.bounds, w.codes, w.idx);
  size_t tb = w.cub_bytes;
  cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.codes, w.codes_sorted, w.idx, w.order, P, 0, 30, s);
  knn_boxes_kernel<<<nboxes, KNN_BOX, 0, s>>>(P, points, w.order, w.sorted, (Aabb*)w.boxes);
  knn_kernel<<<nboxes, KNN_BOX, 0, s>>>(P, nboxes, w.sorted, w.order, (const Aabb*)w.boxes, mean_dist2);
  return cudaGetLastError();
}

}  // namespace dgm
