// mc.cu -- differentiable marching cubes on a G^3 scalar grid (iso-surface extraction + backward).
//
// Stands in for `diso.DiffMC.__call__(grid, deform=None, isovalue=0.0)` as DG-Mesh calls it
// (dgmesh/utils/renderer.py:171, dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:704,851):
//   phi[G,G,G] fp32 -> verts[V,3] in [0,1]^3 (grid index / (G-1)), faces[F,3] int32.
// `diso` is a third-party CUDA package that is not in the reference tree: its
// vertex / face ORDER cannot be reproduced, only the surface.  Case tables are derived by
// tools/gen_mc_tables.py (watertight by construction).
//
// Warp-cooperative sweep (z fastest: a warp owns a strip of 32 consecutive nodes (i, j, k0..k0+31) and
// loads the four phi rows (i|i+1, j|j+1) of its strip once, coalesced; the k+1 neighbours come from the
// same rows), hierarchical counting instead of a per-node scan:
//   count   per CTA (8 strips = 256 nodes): ONE packed (vertices | triangles << 32) total      [reads phi once]
//   scan    cub::DeviceScan over the ~n/256 CTA totals (library primitive), totals -> device + mapped host
//   emit    only CTAs that own something re-read their strips (a few % of the grid): intra-CTA prefix sums
//           give every vertex its index; vertices are written, the owner table vid[node] = first vertex |
//           edge mask << 29 is filled SPARSELY (only nodes that own vertices are ever looked up), faces are
//           written as encoded owner references (node * 3 + axis), and vsrc[v] remembers each vertex's edge
//   resolve 3F threads turn the owner references into vertex indices through vid[]
// Backward: one thread per VERTEX (vsrc), d verts / d phi through t = (iso - phi0) / (phi1 - phi0).
// Scratch traffic is ~8 B per CTA instead of 26 B per node: the sweep is bound by reading phi once
// (4 G^3 bytes) rather than by its own bookkeeping.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "mc_kernels.h"
#include "mc_tables.h"

namespace dgm {

__constant__ unsigned char c_ntri[256];
__constant__ signed char c_tri[256][MC_MAX_TRI * 3];
__constant__ unsigned char c_edge_lo[12];

#define MC_STRIPS 8  // warps (strips of 32 nodes) per CTA

struct McGeom {
  int G, spr;        // grid size, strips per (i, j) row = ceil(G / 32)
  size_t nstrips;    // G * G * spr
};
__host__ __device__ inline McGeom mc_geom(int G) {
  McGeom g;
  g.G = G;
  g.spr = (G + 31) / 32;
  g.nstrips = (size_t)G * G * g.spr;
  return g;
}
size_t mc_num_blocks(int G) { return (mc_geom(G).nstrips + MC_STRIPS - 1) / MC_STRIPS; }

// what one lane knows about its node after the strip loads
struct McNode {
  bool valid;
  int i, j, k;
  size_t id;
  unsigned mask, cs;  // owned-edge mask (bit a: the edge towards +axis a carries a vertex), cell case
  float p0, p1[3];    // phi at the node and at its +x / +y / +z neighbours
};

__device__ __forceinline__ McNode mc_load(const McGeom& g, size_t strip, unsigned lane, const float* __restrict__ phi,
                                          float iso) {
  McNode n;
  const int G = g.G;
  n.valid = strip < g.nstrips;
  const size_t row = n.valid ? strip / g.spr : 0;
  n.k = (int)((n.valid ? strip % g.spr : 0) * 32 + lane);
  n.j = (int)(row % G);
  n.i = (int)(row / G);
  n.valid = n.valid && n.k < G;
  n.id = ((size_t)n.i * G + n.j) * G + n.k;
  n.mask = n.cs = 0;
  n.p0 = n.p1[0] = n.p1[1] = n.p1[2] = 0.f;
  if (!n.valid) return n;
  const bool xi = n.i + 1 < G, yj = n.j + 1 < G, zk = n.k + 1 < G;
  const float* r00 = phi + n.id;
  const size_t sx = (size_t)G * G, sy = G;
  float v[8];
  v[0] = r00[0];
  v[1] = xi ? r00[sx] : 0.f;
  v[2] = yj ? r00[sy] : 0.f;
  v[3] = (xi && yj) ? r00[sx + sy] : 0.f;
  v[4] = zk ? r00[1] : 0.f;
  v[5] = (xi && zk) ? r00[sx + 1] : 0.f;
  v[6] = (yj && zk) ? r00[sy + 1] : 0.f;
  v[7] = (xi && yj && zk) ? r00[sx + sy + 1] : 0.f;
  n.p0 = v[0], n.p1[0] = v[1], n.p1[1] = v[2], n.p1[2] = v[4];
  const bool s0 = v[0] < iso;
  if (xi && ((v[1] < iso) != s0)) n.mask |= 1;
  if (yj && ((v[2] < iso) != s0)) n.mask |= 2;
  if (zk && ((v[4] < iso) != s0)) n.mask |= 4;
  if (xi && yj && zk) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (v[c] < iso) n.cs |= 1u << c;  // corner c = (i + (c & 1), j + ((c >> 1) & 1), k + (c >> 2))
  }
  return n;
}

__global__ void __launch_bounds__(32 * MC_STRIPS) mc_count_kernel(McGeom g, const float* __restrict__ phi, float iso,
                                                                  unsigned long long* __restrict__ blk_counts) {
  __shared__ unsigned s_part[MC_STRIPS];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const McNode n = mc_load(g, (size_t)blockIdx.x * MC_STRIPS + wid, lane, phi, iso);
  unsigned c = __popc(n.mask) | ((unsigned)c_ntri[n.cs] << 16);  // <= 96 / 160 per warp: 16 bits each suffice
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) s_part[wid] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long nv = 0, nt = 0;
#pragma unroll
    for (int w = 0; w < MC_STRIPS; ++w) nv += s_part[w] & 0xffffu, nt += s_part[w] >> 16;
    blk_counts[blockIdx.x] = nv | (nt << 32);
  }
}

__global__ void mc_totals_kernel(size_t nb, const unsigned long long* __restrict__ counts,
                                 const unsigned long long* __restrict__ offsets, int32_t* __restrict__ totals,
                                 volatile int32_t* __restrict__ totals_host) {
  const unsigned long long t = offsets[nb - 1] + counts[nb - 1];
  totals[0] = (int32_t)(t & 0xffffffffull);
  totals[1] = (int32_t)(t >> 32);
  if (totals_host) {
    totals_host[0] = totals[0];
    totals_host[1] = totals[1];
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(32 * MC_STRIPS) mc_emit_kernel(
    McGeom g, const float* __restrict__ phi, float iso, const unsigned long long* __restrict__ blk_counts,
    const unsigned long long* __restrict__ blk_offsets, uint32_t* __restrict__ vid, uint32_t* __restrict__ vsrc,
    float* __restrict__ verts, long long V_cap, int32_t* __restrict__ faces, long long F_cap) {
  __shared__ unsigned s_part[MC_STRIPS];
  if (blk_counts[blockIdx.x] == 0) return;  // nothing owned here (the vast majority of CTAs)
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const McNode n = mc_load(g, (size_t)blockIdx.x * MC_STRIPS + wid, lane, phi, iso);
  const unsigned mine = __popc(n.mask) | ((unsigned)c_ntri[n.cs] << 16);
  unsigned inc = mine;  // inclusive prefix inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += up;
  }
  if (lane == 31) s_part[wid] = inc;
  __syncthreads();
  unsigned before = 0;
  for (unsigned w = 0; w < wid; ++w) before += s_part[w];
  const unsigned excl = before + inc - mine;
  const unsigned long long base = blk_offsets[blockIdx.x];
  const unsigned long long v0 = (base & 0xffffffffull) + (excl & 0xffffu);
  const unsigned long long f0 = (base >> 32) + (excl >> 16);
  const int G = g.G;
  if (n.mask) {
    vid[n.id] = (uint32_t)v0 | (n.mask << 29);
    const float gm1 = (float)(G - 1);
    const float basep[3] = {(float)n.i, (float)n.j, (float)n.k};
    unsigned long long v = v0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (n.mask & (1u << a)) {
        if ((long long)v < V_cap) {
          const float t = (iso - n.p0) / (n.p1[a] - n.p0);
          float pos[3] = {basep[0], basep[1], basep[2]};
          pos[a] += t;
          verts[3 * v + 0] = pos[0] / gm1;  // IEEE division: identical to the numpy restatement
          verts[3 * v + 1] = pos[1] / gm1;
          verts[3 * v + 2] = pos[2] / gm1;
          vsrc[v] = (uint32_t)(n.id * 3 + a);
        }
        ++v;
      }
    }
  }
  const int nt = c_ntri[n.cs];
  unsigned long long f = f0;
  for (int t = 0; t < nt; ++t, ++f) {
    if ((long long)f >= F_cap) break;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = c_tri[n.cs][3 * t + q];
      const int c = c_edge_lo[e], axis = e >> 2;
      const size_t owner = ((size_t)(n.i + (c & 1)) * G + (n.j + ((c >> 1) & 1))) * G + (n.k + (c >> 2));
      faces[3 * f + q] = (int32_t)(owner * 3 + axis);  // resolved by mc_resolve_kernel
    }
  }
}

__global__ void __launch_bounds__(256) mc_resolve_kernel(const int32_t* __restrict__ totals, long long F_cap,
                                                         const uint32_t* __restrict__ vid,
                                                         int32_t* __restrict__ faces) {
  const long long n = 3 * min((long long)totals[1], F_cap);
  const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const uint32_t e = (uint32_t)faces[x];
  const uint32_t w = vid[e / 3];
  faces[x] = (int32_t)((w & 0x1fffffffu) + __popc((w >> 29) & ((1u << (e % 3)) - 1u)));
}

// dL/dphi += dL/dverts . d verts/d phi   (only the coordinate along the edge moves); one thread per vertex
__global__ void __launch_bounds__(256) mc_backward_kernel(int G, int V, const float* __restrict__ phi, float iso,
                                                          const uint32_t* __restrict__ vsrc,
                                                          const float* __restrict__ dverts,
                                                          float* __restrict__ dphi) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const uint32_t e = vsrc[v];
  const size_t id = e / 3;
  const int a = (int)(e % 3);
  const size_t nb = id + (a == 0 ? (size_t)G * G : (a == 1 ? (size_t)G : 1));
  const float p0 = phi[id], p1 = phi[nb];
  const float d = p1 - p0;
  const float gr = dverts[3 * (size_t)v + a] / (float)(G - 1);
  // t = (iso - p0)/(p1 - p0):  dt/dp0 = (iso - p1)/d^2,  dt/dp1 = -(iso - p0)/d^2
  atomicAdd(&dphi[id], gr * (iso - p1) / (d * d));
  atomicAdd(&dphi[nb], -gr * (iso - p0) / (d * d));
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
  const size_t n = (size_t)G * G * G, nb = mc_num_blocks(G);
  w.blk_counts = carve<unsigned long long>(p, nb);
  w.blk_offsets = carve<unsigned long long>(p, nb);
  w.totals = carve<int32_t>(p, 4);
  w.cub_bytes = mc_cub_bytes(nb);
  w.cub_temp = carve<char>(p, w.cub_bytes);
  w.vid = carve<uint32_t>(p, n);    // sparse: only entries of vertex-owning nodes are ever written or read
  w.vsrc = carve<uint32_t>(p, n);   // [V] used; worst case V <= 3 n is never near: capacity n (>= V for any field
                                    //     that is not pure checkerboard noise); the caller's V_cap <= n is enforced
  if (bytes) *bytes = size_t(p - base) + 128;
  return w;
}

cudaError_t launch_mc_count(int G, const float* phi, float iso, void* ws, int32_t* totals, int32_t* totals_host,
                            cudaEvent_t ev, cudaStream_t s) {
  cudaError_t e = upload_tables();
  if (e != cudaSuccess) return e;
  McWS w = McWS::from((char*)ws, G);
  const McGeom g = mc_geom(G);
  const size_t nb = mc_num_blocks(G);
  mc_count_kernel<<<(unsigned)nb, 32 * MC_STRIPS, 0, s>>>(g, phi, iso, w.blk_counts);
  size_t tb = w.cub_bytes;
  cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, w.blk_counts, w.blk_offsets, (int)nb, s);
  mc_totals_kernel<<<1, 1, 0, s>>>(nb, w.blk_counts, w.blk_offsets, w.totals, nullptr);
  mc_totals_kernel<<<1, 1, 0, s>>>(nb, w.blk_counts, w.blk_offsets, totals, (volatile int32_t*)totals_host);
  if (ev) cudaEventRecord(ev, s);
  return cudaGetLastError();
}

cudaError_t launch_mc_emit(int G, const float* phi, float iso, void* ws, float* verts, long long V_cap,
                           int32_t* faces, long long F_cap, cudaStream_t s) {
  McWS w = McWS::from((char*)ws, G);
  const McGeom g = mc_geom(G);
  const size_t nb = mc_num_blocks(G);
  const size_t n = (size_t)G * G * G;
  if (V_cap > (long long)n) V_cap = (long long)n;
  mc_emit_kernel<<<(unsigned)nb, 32 * MC_STRIPS, 0, s>>>(g, phi, iso, w.blk_counts, w.blk_offsets, w.vid, w.vsrc, verts,
                                                         V_cap, faces, F_cap);
  if (F_cap > 0)
    mc_resolve_kernel<<<(unsigned)((3 * F_cap + 255) / 256), 256, 0, s>>>(w.totals, F_cap, w.vid, faces);
  return cudaGetLastError();
}

cudaError_t launch_mc_backward(int G, int V, const float* phi, float iso, void* ws, const float* dverts, float* dphi,
                               cudaStream_t s) {
  McWS w = McWS::from((char*)ws, G);
  const size_t n = (size_t)G * G * G;
  cudaMemsetAsync(dphi, 0, sizeof(float) * n, s);
  if (V > 0) mc_backward_kernel<<<(V + 255) / 256, 256, 0, s>>>(G, V, phi, iso, w.vsrc, dverts, dphi);
  return cudaGetLastError();
}

}  // namespace dgm
