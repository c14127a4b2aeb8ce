// mc.cu -- differentiable marching cubes on a G^3 scalar grid (iso-surface extraction + backward).
//
// Stands in for `diso.DiffMC.__call__(grid, deform=None, isovalue=0.0)` as DG-Mesh calls it
// (dgmesh/utils/renderer.py:171, dgmesh/scene/gaussian_model_dpsr_dynamic_anchor.py:704,851):
//   phi[G,G,G] fp32 -> verts[V,3] in [0,1]^3 (grid index / (G-1)), faces[F,3] int32.
// `diso` is a third-party CUDA package that is not in the reference tree: its
// vertex / face ORDER cannot be reproduced, only the surface.  Case tables are derived by
// tools/gen_mc_tables.py (watertight by construction).
//
// Warp-per-brick sweep.  A BRICK is 8 (j) x 32 (k) nodes of one i-plane; a warp owns one brick: lane = k, and
// every lane loads the 9 rows (j .. j+8) of planes i and i+1 at its k -- 18 independent, fully coalesced loads
// in flight per lane, 2.25 loads per node instead of 8 -- the k+1 neighbours come from the next lane by shuffle
// (lane 31 fetches its own).  No shared memory, no CTA barrier.  Hierarchical counting instead of a per-node scan:
//   count   per brick: ONE packed (vertices | triangles << 32) total                              [reads phi once]
//   scan    cub::DeviceScan over the n/256 brick totals (library primitive), totals -> device + mapped host
//   emit    only warps whose brick owns something re-read it (a few % of the grid): a warp prefix sum gives
//           every vertex its index; vertices are written, the owner table vid[node] = first vertex |
//           edge mask << 29 is filled SPARSELY (only nodes that own vertices are ever looked up), faces are
//           written as encoded owner references (node * 3 + axis), and vsrc[v] remembers each vertex's edge
//   resolve 3F threads turn the owner references into vertex indices through vid[]
// Backward: one thread per VERTEX (vsrc), d verts / d phi through t = (iso - phi0) / (phi1 - phi0).
// Scratch traffic is 8 B per brick instead of 26 B per node: the sweep is bound by reading phi once
// (4 G^3 bytes) rather than by its own bookkeeping.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "mc_kernels.h"
#include "mc_tables.h"

namespace dgm {

// case tables in GLOBAL memory, read through the read-only path: the case index differs from lane to lane, and a
// __constant__ access with divergent addresses is replayed once per distinct address (measured: the emit pass
// spent most of its time there); L1-cached loads are not
__device__ unsigned char g_ntri[256];
__device__ signed char g_tri[256][MC_MAX_TRI * 3];
__device__ __forceinline__ unsigned mc_ntri(unsigned cs) { return (cs == 0u || cs == 255u) ? 0u : __ldg(&g_ntri[cs]); }
// corner of edge e = 4 * axis + combo with the lower index (MC_EDGE_LO of mc_tables.h, in closed form)
__device__ __forceinline__ int mc_edge_lo(int e) {
  const int axis = e >> 2, k = e & 3;
  return axis == 0 ? (k << 1) : (axis == 1 ? ((k & 1) | ((k & 2) << 1)) : k);
}

#define MC_WARPS 8  // bricks (warps) per CTA
#define MC_BJ 8     // j-rows per brick

struct McGeom {
  int G, spr, njb;   // grid size, k-chunks per row = ceil(G / 32), j-blocks per plane = ceil(G / MC_BJ)
  size_t nbricks;    // G * njb * spr
};
__host__ __device__ inline McGeom mc_geom(int G) {
  McGeom g;
  g.G = G;
  g.spr = (G + 31) / 32;
  g.njb = (G + MC_BJ - 1) / MC_BJ;
  g.nbricks = (size_t)G * g.njb * g.spr;
  return g;
}
size_t mc_num_blocks(int G) { return mc_geom(G).nbricks; }

// one lane's view of its brick: phi at (i + p, j0 + r, k) and at k + 1, r = 0..MC_BJ (out of range = 0)
struct McBrick {
  int i, j0, k;
  bool xi, zk, kin;
  float a[2][MC_BJ + 1], z[2][MC_BJ + 1];
};

__device__ __forceinline__ void mc_load_brick(const McGeom& g, size_t brick, unsigned lane,
                                              const float* __restrict__ phi, McBrick& B) {
  const int G = g.G;
  const int kc = (int)(brick % g.spr);
  const size_t t = brick / g.spr;
  B.j0 = (int)(t % g.njb) * MC_BJ;
  B.i = (int)(t / g.njb);
  B.k = kc * 32 + (int)lane;
  B.kin = B.k < G;
  B.xi = B.i + 1 < G;
  B.zk = B.k + 1 < G;
  const float* base = phi + ((size_t)B.i * G + B.j0) * G + B.k;
  const size_t sx = (size_t)G * G;
#pragma unroll
  for (int r = 0; r <= MC_BJ; ++r) {
    const bool jin = B.j0 + r < G;
    B.a[0][r] = (jin && B.kin) ? base[(size_t)r * G] : 0.f;
    B.a[1][r] = (jin && B.kin && B.xi) ? base[sx + (size_t)r * G] : 0.f;
  }
  // the k + 1 column: lanes 0..30 take their neighbour's value, lane 31 reads the next chunk's first element
#pragma unroll
  for (int r = 0; r <= MC_BJ; ++r) {
    const bool jin = B.j0 + r < G;
    float e0 = 0.f, e1 = 0.f;
    if (lane == 31 && jin && B.zk) {
      e0 = base[(size_t)r * G + 1];
      if (B.xi) e1 = base[sx + (size_t)r * G + 1];
    }
    const float n0 = __shfl_down_sync(0xffffffffu, B.a[0][r], 1), n1 = __shfl_down_sync(0xffffffffu, B.a[1][r], 1);
    B.z[0][r] = lane == 31 ? e0 : n0;
    B.z[1][r] = lane == 31 ? e1 : n1;
  }
}

// node r of the lane: owned-edge mask (bit a: the edge towards +axis a carries a vertex) and cell case
__device__ __forceinline__ void mc_classify(const McGeom& g, const McBrick& B, int r, float iso, unsigned& mask,
                                            unsigned& cs) {
  mask = cs = 0;
  const int j = B.j0 + r;
  if (!B.kin || j >= g.G) return;
  const bool yj = j + 1 < g.G;
  const float v0 = B.a[0][r], v1 = B.a[1][r], v2 = B.a[0][r + 1], v3 = B.a[1][r + 1];
  const float v4 = B.z[0][r], v5 = B.z[1][r], v6 = B.z[0][r + 1], v7 = B.z[1][r + 1];
  const bool s0 = v0 < iso;
  if (B.xi && ((v1 < iso) != s0)) mask |= 1;
  if (yj && ((v2 < iso) != s0)) mask |= 2;
  if (B.zk && ((v4 < iso) != s0)) mask |= 4;
  if (B.xi && yj && B.zk) {  // corner c = (i + (c & 1), j + ((c >> 1) & 1), k + (c >> 2))
    cs = (unsigned)s0 | ((unsigned)(v1 < iso) << 1) | ((unsigned)(v2 < iso) << 2) | ((unsigned)(v3 < iso) << 3) |
         ((unsigned)(v4 < iso) << 4) | ((unsigned)(v5 < iso) << 5) | ((unsigned)(v6 < iso) << 6) |
         ((unsigned)(v7 < iso) << 7);
  }
}

__global__ void __launch_bounds__(32 * MC_WARPS) mc_count_kernel(McGeom g, const float* __restrict__ phi, float iso,
                                                                 unsigned long long* __restrict__ blk_counts) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t brick = (size_t)blockIdx.x * MC_WARPS + wid;
  if (brick >= g.nbricks) return;
  McBrick B;
  mc_load_brick(g, brick, lane, phi, B);
  // the common case (>95 % of a 288^3 grid): every value the brick touches lies on one side of the level set --
  // nothing to classify.  Lanes look at their own valid values (lane 31 also at its k + 1 column).
  {
    bool neg = false, pos = false;
#pragma unroll
    for (int r = 0; r <= MC_BJ; ++r) {
      const bool jin = B.j0 + r < g.G;
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        const bool ok = jin && B.kin && (pl == 0 || B.xi);
        const bool s = B.a[pl][r] < iso;
        neg |= ok && s;
        pos |= ok && !s;
        if (lane == 31) {
          const bool okz = jin && B.zk && (pl == 0 || B.xi);
          const bool sz = B.z[pl][r] < iso;
          neg |= okz && sz;
          pos |= okz && !sz;
        }
      }
    }
    if (!(__any_sync(0xffffffffu, neg) && __any_sync(0xffffffffu, pos))) {
      if (lane == 0) blk_counts[brick] = 0ull;
      return;
    }
  }
  unsigned c = 0;  // vertices | triangles << 16: <= 768 / 1280 per brick
#pragma unroll
  for (int r = 0; r < MC_BJ; ++r) {
    unsigned mask, cs;
    mc_classify(g, B, r, iso, mask, cs);
    c += __popc(mask) | (mc_ntri(cs) << 16);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) blk_counts[brick] = (unsigned long long)(c & 0xffffu) | ((unsigned long long)(c >> 16) << 32);
}

// {V, F} -> the workspace copy (resolve pass), the caller's device copy and its host-mapped mirror
__global__ void mc_totals_kernel(size_t nb, const unsigned long long* __restrict__ counts,
                                 const unsigned long long* __restrict__ offsets, int32_t* __restrict__ ws_totals,
                                 int32_t* __restrict__ totals, volatile int32_t* __restrict__ totals_host) {
  const unsigned long long t = offsets[nb - 1] + counts[nb - 1];
  const int32_t V = (int32_t)(t & 0xffffffffull), F = (int32_t)(t >> 32);
  ws_totals[0] = V;
  ws_totals[1] = F;
  totals[0] = V;
  totals[1] = F;
  if (totals_host) {
    totals_host[0] = V;
    totals_host[1] = F;
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(32 * MC_WARPS) mc_emit_kernel(
    McGeom g, const float* __restrict__ phi, float iso, const unsigned long long* __restrict__ blk_counts,
    const unsigned long long* __restrict__ blk_offsets, uint32_t* __restrict__ vid, uint32_t* __restrict__ vsrc,
    float* __restrict__ verts, long long V_cap, int32_t* __restrict__ faces, long long F_cap) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t brick = (size_t)blockIdx.x * MC_WARPS + wid;
  if (brick >= g.nbricks || blk_counts[brick] == 0) return;  // nothing owned here (the vast majority of bricks)
  McBrick B;
  mc_load_brick(g, brick, lane, phi, B);
  unsigned masks = 0, css[MC_BJ];  // 3 bits of mask per node, packed
  unsigned mine = 0;
#pragma unroll
  for (int r = 0; r < MC_BJ; ++r) {
    unsigned mask;
    mc_classify(g, B, r, iso, mask, css[r]);
    masks |= mask << (3 * r);
    mine += __popc(mask) | (mc_ntri(css[r]) << 16);
  }
  unsigned inc = mine;  // inclusive prefix over the lanes (order inside the brick: lane-major, then r)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned up = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += up;
  }
  if (mine == 0) return;
  const unsigned excl = inc - mine;
  const unsigned long long base = blk_offsets[brick];
  unsigned long long v = (base & 0xffffffffull) + (excl & 0xffffu);
  unsigned long long f = (base >> 32) + (excl >> 16);
  const int G = g.G;
  const float gm1 = (float)(G - 1);
#pragma unroll
  for (int r = 0; r < MC_BJ; ++r) {
    const unsigned mask = (masks >> (3 * r)) & 7u, cs = css[r];
    if (mask == 0 && cs == 0) continue;
    const int j = B.j0 + r;
    const size_t id = ((size_t)B.i * G + j) * G + B.k;
    if (mask) {
      vid[id] = (uint32_t)v | (mask << 29);
      const float p0 = B.a[0][r];
      const float p1[3] = {B.a[1][r], B.a[0][r + 1], B.z[0][r]};
      const float basep[3] = {(float)B.i, (float)j, (float)B.k};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (mask & (1u << a)) {
          if ((long long)v < V_cap) {
            const float t = (iso - p0) / (p1[a] - p0);
            float pos[3] = {basep[0], basep[1], basep[2]};
            pos[a] += t;
            verts[3 * v + 0] = pos[0] / gm1;  // IEEE division: identical to the numpy restatement
            verts[3 * v + 1] = pos[1] / gm1;
            verts[3 * v + 2] = pos[2] / gm1;
            vsrc[v] = (uint32_t)(id * 3 + a);
          }
          ++v;
        }
      }
    }
    const int nt = (int)mc_ntri(cs);
    for (int t = 0; t < nt; ++t, ++f) {
      if ((long long)f >= F_cap) break;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int e = __ldg(&g_tri[cs][3 * t + q]);
        const int c = mc_edge_lo(e), axis = e >> 2;
        const size_t owner = ((size_t)(B.i + (c & 1)) * G + (j + ((c >> 1) & 1))) * G + (B.k + (c >> 2));
        faces[3 * f + q] = (int32_t)(owner * 3 + axis);  // resolved by mc_resolve_kernel
      }
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

static unsigned long long g_tables_uploaded = 0;  // per device
static cudaError_t upload_tables() {
  if (!once_per_device(g_tables_uploaded)) return cudaSuccess;
  cudaError_t e = cudaMemcpyToSymbol(g_ntri, MC_NTRI, sizeof(MC_NTRI));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_tri, MC_TRI, sizeof(MC_TRI));
  for (int i = 0; i < 12 && e == cudaSuccess; ++i) {  // the closed form must be the generated table
    const int axis = i >> 2, k = i & 3;
    const int lo = axis == 0 ? (k << 1) : (axis == 1 ? ((k & 1) | ((k & 2) << 1)) : k);
    if (lo != MC_EDGE_LO[i]) e = cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) g_tables_uploaded = 0;  // retry on the next call
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
  mc_count_kernel<<<(unsigned)((nb + MC_WARPS - 1) / MC_WARPS), 32 * MC_WARPS, 0, s>>>(g, phi, iso, w.blk_counts);
  size_t tb = w.cub_bytes;
  cub::DeviceScan::ExclusiveSum(w.cub_temp, tb, w.blk_counts, w.blk_offsets, (int)nb, s);
  mc_totals_kernel<<<1, 1, 0, s>>>(nb, w.blk_counts, w.blk_offsets, w.totals, totals, (volatile int32_t*)totals_host);
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
  mc_emit_kernel<<<(unsigned)((nb + MC_WARPS - 1) / MC_WARPS), 32 * MC_WARPS, 0, s>>>(
      g, phi, iso, w.blk_counts, w.blk_offsets, w.vid, w.vsrc, verts, V_cap, faces, F_cap);
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
