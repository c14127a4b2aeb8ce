// mc.cu -- differentiable marching cubes on a G^3 scalar grid (iso-surface extraction + backward).
//
// Stands in for `diso.DiffMC.__call__(grid, deform=None, isovalue=0.0)` as DG-Mesh calls it
// (dgmesh/utils/renderer.py:171, dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:704,851):
//   phi[G,G,G] fp32 -> verts[V,3] in [0,1]^3 (grid index / (G-1)), faces[F,3] int32.
// `diso` is a third-party CUDA package that is not in the reference tree: its
// vertex / face ORDER cannot be reproduced, only the surface.  Case tables are derived by
// tools/gen_mc_tables.py (watertight by construction).
//
// Three passes, each one thread per grid node with the z index fastest (coalesced phi reads):
//   count   per node: which of its 3 owned edges (+x,+y,+z) carry a vertex, and how many
//           triangles its cell emits;  packed (nv | nt << 32) for ONE 64-bit exclusive scan
//   scan    cub::DeviceScan (library primitive)
//   emit    vertices (shared between cells through the owning node -> no duplicates) + faces
// Backward: d verts / d phi through the edge interpolation t = (iso - phi0) / (phi1 - phi0).
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "mc_kernels.h"
#include "mc_tables.h"

namespace dgm {

__constant__ unsigned char c_ntri[256];
__constant__ signed char c_tri[256][MC_MAX_TRI * 3];
__constant__ unsigned char c_edge_lo[12];

__device__ __forceinline__ size_t nidx(int i, int j, int k, int G) { return ((size_t)i * G + j) * G + k; }

// info = case (8 bits) | owned-edge mask (3 bits) << 8
__global__ void __launch_bounds__(256) mc_count_kernel(int G, const float* __restrict__ phi, float iso,
                                                       unsigned long long* __restrict__ counts,
                                                       uint16_t* __restrict__ info) {
  const size_t n = (size_t)G * G * G;
  const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n) return;
  const int k = (int)(id % G), j = (int)((id / G) % G), i = (int)(id / ((size_t)G * G));
  const bool s0 = phi[id] < iso;
  unsigned mask = 0;
  if (i + 1 < G && (phi[nidx(i + 1, j, k, G)] < iso) != s0) mask |= 1;
  if (j + 1 < G && (phi[nidx(i, j + 1, k, G)] < iso) != s0) mask |= 2;
  if (k + 1 < G && (phi[nidx(i, j, k + 1, G)] < iso) != s0) mask |= 4;
  unsigned cs = 0;
  if (i + 1 < G && j + 1 < G && k + 1 < G) {
#pragma unroll
    for (int v = 0; v < 8; ++v)
      if (phi[nidx(i + (v & 1), j + ((v >> 1) & 1), k + (v >> 2), G)] < iso) cs |= 1u << v;
  }
  info[id] = (uint16_t)(cs | (mask << 8));
  counts[id] = (unsigned long long)__popc(mask) | ((unsigned long long)c_ntri[cs] << 32);
}

__global__ void mc_totals_kernel(size_t n, const unsigned long long* __restrict__ counts,
                                 const unsigned long long* __restrict__ offsets, int32_t* __restrict__ totals) {
  const unsigned long long t = offsets[n - 1] + counts[n - 1];
  totals[0] = (int32_t)(t & 0xffffffffull);
  totals[1] = (int32_t)(t >> 32);
}

__device__ __forceinline__ uint32_t vertex_of_edge(int i, int j, int k, int e, int G,
                                                   const unsigned long long* __restrict__ offsets,
                                                   const uint16_t* __restrict__ info) {
  const int c = c_edge_lo[e], axis = e >> 2;
  const size_t owner = nidx(i + (c & 1), j + ((c >> 1) & 1), k + (c >> 2), G);
  const unsigned mask = info[owner] >> 8;
  return (uint32_t)(offsets[owner] & 0xffffffffull) + __popc(mask & ((1u << axis) - 1u));
}

__global__ void __launch_bounds__(256) mc_emit_kernel(int G, const float* __restrict__ phi, float iso,
                                                      const unsigned long long* __restrict__ offsets,
                                                      const uint16_t* __restrict__ info, float* __restrict__ verts,
                                                      int32_t* __restrict__ faces) {
  const size_t n = (size_t)G * G * G;
  const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n) return;
  const unsigned inf = info[id];
  if (inf == 0) return;
  const int k = (int)(id % G), j = (int)((id / G) % G), i = (int)(id / ((size_t)G * G));
  const unsigned mask = inf >> 8, cs = inf & 0xff;
  const unsigned long long off = offsets[id];
  const float gm1 = (float)(G - 1);
  if (mask) {
    const float p0 = phi[id];
    uint32_t v = (uint32_t)(off & 0xffffffffull);
    const float base[3] = {(float)i, (float)j, (float)k};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (mask & (1u << a)) {
        const float p1 = phi[nidx(i + (a == 0), j + (a == 1), k + (a == 2), G)];
        const float t = (iso - p0) / (p1 - p0);
        float pos[3] = {base[0], base[1], base[2]};
        pos[a] += t;
        verts[3 * (size_t)v + 0] = pos[0] / gm1;  // IEEE division: identical to the numpy restatement
        verts[3 * (size_t)v + 1] = pos[1] / gm1;
        verts[3 * (size_t)v + 2] = pos[2] / gm1;
        ++v;
      }
    }
  }
  const int nt = c_ntri[cs];
  if (nt) {
    size_t f = (size_t)(off >> 32);
    for (int t = 0; t < nt; ++t, ++f) {
#pragma unroll
      for (int q = 0; q < 3; ++q)
        faces[3 * f + q] = (int32_t)vertex_of_edge(i, j, k, c_tri[cs][3 * t + q], G, offsets, info);
    }
  }
}

// dL/dphi += dL/dverts . d verts/d phi   (only the coordinate along the edge moves)
__global__ void __launch_bounds__(256) mc_backward_kernel(int G, const float* __restrict__ phi, float iso,
                                                          const unsigned long long* __restrict__ offsets,
                                                          const uint16_t* __restrict__ info,
                                                          const float* __restrict__ dverts,
                                                          float* __restrict__ dphi) {
  const size_t n = (size_t)G * G * G;
  const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n) return;
  const unsigned mask = info[id] >> 8;
  if (!mask) return;
  const int k = (int)(id % G), j = (int)((id / G) % G), i = (int)(id / ((size_t)G * G));
  const float p0 = phi[id];
  const float inv = 1.0f / (float)(G - 1);
  uint32_t v = (uint32_t)(offsets[id] & 0xffffffffull);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (mask & (1u << a)) {
      const size_t nb = nidx(i + (a == 0), j + (a == 1), k + (a == 2), G);
      const float p1 = phi[nb];
      const float d = p1 - p0;
      const float g = dverts[3 * (size_t)v + a] * inv;
      // t = (iso - p0)/(p1 - p0):  dt/dp0 = (iso - p1)/d^2,  dt/dp1 = -(iso - p0)/d^2
      atomicAdd(&dphi[id], g * (iso - p1) / (d * d));
      atomicAdd(&dphi[nb], -g * (iso - p0) / (d * d));
      ++v;
    }
  }
}

static bool g_tables_uploaded = false;
static cudaError_t upload_tables() {
  if (g_tables_uploaded) return cudaSuccess;
  cudaError_t e = cudaMemcpyToSymbol(c_ntri, MC_NTRI, sizeof(MC_NTRI));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_tri, MC_TRI, sizeof(MC_TRI));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_edge_lo, MC_EDGE_LO, sizeof(MC_EDGE_LO));
  g_tables_uploaded = (e == cudaSuccess);
  return e;
}

size_t mc_cub_bytes(size_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                (int)n);
  return bytes;
}

McWS McWS::from(char* base, int G, size_t* bytes) {
  char* p = base;
  McWS w;
  const size_t n = (size_t)G * G * G;
  w.counts = carve<unsigned long long>(p, n);
  w.offsets = carve<unsigned long long>(p, n);
  w.info = carve<uint16_t>(p, n);
  w.cub_bytes = mc_cub_bytes(n);
  w.cub_temp = carve<char>(p, w.cub_bytes);
  if (bytes) *bytes = size_t(p - base) + 128;
  return w;
}

cudaError_t launch_mc_count(int G, const float* phi, float iso, void* ws, int32_t* totals, cudaStream_t s) {
  cudaError_t e = upload_tables();
  if (e != cudaSuccess) return e;
  McWS w = McWS::from((char*)ws, G);
  const size_t n = (size_t)G * G * G;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  mc_count_kernel<<<blocks, 256, 0, s>>>(G, phi, iso, w.counts, w.info);
  size_t tb = w.cub_bytes;
  cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, w.counts, w.offsets, (int)n, s);
  mc_totals_kernel<<<1, 1, 0, s>>>(n, w.counts, w.offsets, totals);
  return cudaGetLastError();
}

cudaError_t launch_mc_emit(int G, const float* phi, float iso, void* ws, float* verts, int32_t* faces,
                           cudaStream_t s) {
  McWS w = McWS::from((char*)ws, G);
  const size_t n = (size_t)G * G * G;
  mc_emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(G, phi, iso, w.offsets, w.info, verts, faces);
  return cudaGetLastError();
}

cudaError_t launch_mc_backward(int G, const float* phi, float iso, void* ws, const float* dverts, float* dphi,
                               cudaStream_t s) {
  McWS w = McWS::from((char*)ws, G);
  const size_t n = (size_t)G * G * G;
  cudaMemsetAsync(dphi, 0, sizeof(float) * n, s);
  mc_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(G, phi, iso, w.offsets, w.info, dverts, dphi);
  return cudaGetLastError();
}

}  // namespace dgm
