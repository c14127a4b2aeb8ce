// raster_fwd.cu -- forward pass of the tile-based 3D-Gaussian rasterizer for sm_100a.
//
// Pipeline (one stream, no host synchronisation):
//   memset(tile counters) -> preprocess_kernel -> tile_scan_kernel -> scatter_kernel
//   -> tile_sort_pack_kernel -> render_fwd_kernel
//
// Results are those of the reference forward (dgr/cuda_rasterizer/forward.cu:155-374,
// rasterizer_impl.cu:198-336): identical radii / tiles_touched / per-tile sorted
// lists / n_contrib, image within fp32 rounding.  The design differs:
//   * no global 64-bit radix sort and no prefix sum over Gaussians: instances are
//     counted per tile, each tile segment is sorted on its own in shared memory by
//     (depth bits, gaussian id) which is exactly the order the reference's stable
//     sort on (tile | depth) yields (emission order == gaussian id);
//   * the sorted segment is materialised as packed records that the blend kernels
//     stream with 1-D TMA bulk copies (cp.async.bulk + mbarrier) instead of gathering
//     per-Gaussian attributes from three arrays in both passes;
//   * each warp owns an 8x4 pixel block and first culls the batch against a
//     conservative alpha >= 1/255 extent, so only Gaussians that can contribute to
//     the block enter the blend loop (the skipped ones are exactly those the
//     reference `continue`s over, so results are unchanged).
#include <stdio.h>
#include "common.cuh"
#include "raster_math.cuh"
#include "raster_kernels.h"

namespace dgm {

// =========================================================== preprocess ====
// One thread per Gaussian (forward.cu:155-256).  Besides the reference outputs it
// counts instances per tile (atomics on T counters) and clears the backward
// accumulators.  Rectangles larger than 32 tiles are counted warp-cooperatively.
__global__ void __launch_bounds__(256) preprocess_kernel(
    int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ scales, float scale_modifier,
    const float* __restrict__ rotations, const float* __restrict__ opacities, const float* __restrict__ shs,
    const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp,
    const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix, const float* __restrict__ cam_pos,
    int W, int H, float tan_fovx, float tan_fovy, float focal_x, float focal_y, int* __restrict__ radii_out,
    GeomWS g, unsigned gx, unsigned gy, uint32_t* __restrict__ tile_counts, int prefiltered) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  bool valid = idx < P;
  uint2 rmin = make_uint2(0, 0), rmax = make_uint2(0, 0);
  float3 p_view = make_float3(0, 0, 0);

  if (valid) {
    g.radii[idx] = 0;
    if (radii_out) radii_out[idx] = 0;
    g.tiles_touched[idx] = 0;
    g.grad_acc[3 * idx + 0] = make_float4(0, 0, 0, 0);
    g.grad_acc[3 * idx + 1] = make_float4(0, 0, 0, 0);
    g.grad_acc[3 * idx + 2] = make_float4(0, 0, 0, 0);
  }
  float3 p_orig = make_float3(0, 0, 0);
  float4 p_hom;
  float p_w;
  float3 p_proj = make_float3(0, 0, 0);
  if (valid) {
    p_orig = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    // near culling (auxiliary.h:139-164)
    p_hom = xform4x4(p_orig, projmatrix);
    p_w = 1.0f / (p_hom.w + 0.0000001f);
    p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);
    p_view = xform4x3(p_orig, viewmatrix);
    if (p_view.z <= 0.2f) {
      if (prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
      }
      valid = false;
    }
  }
  float my_radius = 0.f;
  float2 point_image = make_float2(0, 0);
  float3 conic = make_float3(0, 0, 0);
  if (valid) {
    const float* cov3D;
    if (cov3D_precomp != nullptr) {
      cov3D = cov3D_precomp + idx * 6;
    } else {
      const float3 sc = make_float3(scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]);
      const float4 q = make_float4(rotations[4 * idx], rotations[4 * idx + 1], rotations[4 * idx + 2],
                                   rotations[4 * idx + 3]);
      cov3d_from_scale_rot(sc, scale_modifier, q, g.cov3D + idx * 6);
      cov3D = g.cov3D + idx * 6;
    }
    EwaFrame fr;
    ewa_frame(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix, fr);
    const float3 cov = ewa_cov2d(fr);
    // invert (EWA), forward.cu:217-221
    const float det = (cov.x * cov.z - cov.y * cov.y);
    if (det == 0.0f) {
      valid = false;
    } else {
      const float det_inv = 1.f / det;
      conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);
      // screen-space extent from the larger eigenvalue, forward.cu:227-235
      const float mid = 0.5f * (cov.x + cov.z);
      const float lambda1 = mid + sqrtf(max(0.1f, mid * mid - det));
      const float lambda2 = mid - sqrtf(max(0.1f, mid * mid - det));
      my_radius = ceilf(3.f * sqrtf(max(lambda1, lambda2)));
      point_image = make_float2(ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H));
      tile_rect(point_image, my_radius, rmin, rmax, gx, gy);
      if ((rmax.x - rmin.x) * (rmax.y - rmin.y) == 0) valid = false;
    }
  }
  uint32_t ntiles = 0;
  if (valid) {
    if (colors_precomp == nullptr) {
      bool cl[3];
      const float3 cp = make_float3(cam_pos[0], cam_pos[1], cam_pos[2]);
      const float3 c = sh_to_rgb(D, p_orig, cp, shs + (size_t)idx * M * 3, cl);
      g.clamped[3 * idx + 0] = cl[0];
      g.clamped[3 * idx + 1] = cl[1];
      g.clamped[3 * idx + 2] = cl[2];
      g.rgb[3 * idx + 0] = c.x;
      g.rgb[3 * idx + 1] = c.y;
      g.rgb[3 * idx + 2] = c.z;
    }
    g.depths[idx] = p_view.z;
    g.radii[idx] = my_radius;
    if (radii_out) radii_out[idx] = my_radius;
    g.means2D[idx] = point_image;
    g.conic_opacity[idx] = make_float4(conic.x, conic.y, conic.z, opacities[idx]);
    ntiles = (rmax.y - rmin.y) * (rmax.x - rmin.x);
    g.tiles_touched[idx] = ntiles;
  }
  // ---- per-tile instance counts
  const bool big = ntiles > 32;
  if (valid && !big) {
    for (unsigned y = rmin.y; y < rmax.y; ++y)
      for (unsigned x = rmin.x; x < rmax.x; ++x) atomicAdd(&tile_counts[y * gx + x], 1u);
  }
  unsigned bigmask = __ballot_sync(0xffffffffu, big);
  while (bigmask) {
    const int src = __ffs(bigmask) - 1;
    bigmask &= bigmask - 1;
    const unsigned x0 = __shfl_sync(0xffffffffu, rmin.x, src), y0 = __shfl_sync(0xffffffffu, rmin.y, src);
    const unsigned x1 = __shfl_sync(0xffffffffu, rmax.x, src), y1 = __shfl_sync(0xffffffffu, rmax.y, src);
    const unsigned w = x1 - x0, n = w * (y1 - y0);
    for (unsigned k = lane; k < n; k += 32) atomicAdd(&tile_counts[(y0 + k / w) * gx + (x0 + k % w)], 1u);
  }
}

// ============================================================ tile scan ====
// Exclusive scan of the per-tile counts -> ranges (rasterizer_impl.cu:116-138 writes
// the same [start,end) pairs after its global sort).  Single CTA; T is a few thousand.
__global__ void __launch_bounds__(1024) tile_scan_kernel(int T, const uint32_t* __restrict__ tile_counts,
                                                         uint2* __restrict__ ranges, int32_t* __restrict__ status,
                                                         long long R_cap) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  __shared__ uint32_t s_max;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    s_carry = 0;
    s_max = 0;
  }
  __syncthreads();
  uint32_t local_max = 0;
  for (int base = 0; base < T; base += 1024) {
    const int t = base + tid;
    const uint32_t c = (t < T) ? tile_counts[t] : 0u;
    local_max = max(local_max, c);
    uint32_t v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_off = (wid == 0) ? 0u : s_warp[wid - 1];
    const uint32_t incl = s_carry + warp_off + v;
    if (t < T) ranges[t] = c ? make_uint2(incl - c, incl) : make_uint2(0u, 0u);
    __syncthreads();
    if (tid == 1023) s_carry = incl;
    __syncthreads();
  }
  atomicMax(&s_max, local_max);
  __syncthreads();
  if (tid == 0) {
    status[0] = (int32_t)s_carry;
    status[1] = ((long long)s_carry > R_cap) ? 1 : 0;
    status[2] = (int32_t)s_max;
  }
}

// ============================================================== scatter ====
// Emit one (depth bits << 32 | gaussian id) key per (Gaussian, tile) instance into the
// tile's segment (role of duplicateWithKeys, rasterizer_impl.cu:70-111; the tile id is
// implied by the segment, so the 64-bit word carries the gaussian id instead).
__global__ void __launch_bounds__(256) scatter_kernel(int P, GeomWS g, const uint2* __restrict__ ranges,
                                                      uint32_t* __restrict__ tile_fill,
                                                      unsigned long long* __restrict__ keys, unsigned gx, unsigned gy,
                                                      const int32_t* __restrict__ status) {
  if (status[1]) return;  // overflow: nothing may be written
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  uint2 rmin = make_uint2(0, 0), rmax = make_uint2(0, 0);
  unsigned long long key = 0;
  uint32_t ntiles = 0;
  if (idx < P) {
    const int r = g.radii[idx];
    if (r > 0) {
      tile_rect(g.means2D[idx], r, rmin, rmax, gx, gy);
      ntiles = (rmax.y - rmin.y) * (rmax.x - rmin.x);
      key = ((unsigned long long)__float_as_uint(g.depths[idx]) << 32) | (unsigned)idx;
    }
  }
  const bool big = ntiles > 32;
  if (ntiles && !big) {
    for (unsigned y = rmin.y; y < rmax.y; ++y)
      for (unsigned x = rmin.x; x < rmax.x; ++x) {
        const unsigned t = y * gx + x;
        const uint32_t slot = atomicAdd(&tile_fill[t], 1u);
        keys[ranges[t].x + slot] = key;
      }
  }
  unsigned bigmask = __ballot_sync(0xffffffffu, big);
  while (bigmask) {
    const int src = __ffs(bigmask) - 1;
    bigmask &= bigmask - 1;
    const unsigned x0 = __shfl_sync(0xffffffffu, rmin.x, src), y0 = __shfl_sync(0xffffffffu, rmin.y, src);
    const unsigned x1 = __shfl_sync(0xffffffffu, rmax.x, src), y1 = __shfl_sync(0xffffffffu, rmax.y, src);
    const unsigned long long k64 = __shfl_sync(0xffffffffu, key, src);
    const unsigned w = x1 - x0, n = w * (y1 - y0);
    for (unsigned k = lane; k < n; k += 32) {
      const unsigned t = (y0 + k / w) * gx + (x0 + k % w);
      const uint32_t slot = atomicAdd(&tile_fill[t], 1u);
      keys[ranges[t].x + slot] = k64;
    }
  }
}

// ======================================================= tile sort+pack ====
// One CTA per tile: sort the segment by (depth bits, gaussian id) with a normalised
// bitonic network (all comparators ascending, so virtual +inf padding never moves and
// any length works), then write point_list and the packed blend records.
#define SORT_SMEM_KEYS 6144  // 48 KB static shared memory

__device__ __forceinline__ void bitonic_sort_any(unsigned long long* a, uint32_t n, uint32_t tid, uint32_t nthreads) {
  uint32_t m = 1;
  while (m < n) m <<= 1;
  for (uint32_t k = 2; k <= m; k <<= 1) {
    const uint32_t hk = k >> 1;
    // mirror step
    for (uint32_t i = tid; i < (m >> 1); i += nthreads) {
      const uint32_t blk = i / hk, off = i - blk * hk;
      const uint32_t lo = blk * k + off, hi = blk * k + k - 1 - off;
      if (hi < n) {
        const unsigned long long x = a[lo], y = a[hi];
        if (x > y) {
          a[lo] = y;
          a[hi] = x;
        }
      }
    }
    __syncthreads();
    for (uint32_t j = hk >> 1; j >= 1; j >>= 1) {
      for (uint32_t i = tid; i < (m >> 1); i += nthreads) {
        const uint32_t lo = (i / j) * (j << 1) + (i % j), hi = lo + j;
        if (hi < n) {
          const unsigned long long x = a[lo], y = a[hi];
          if (x > y) {
            a[lo] = y;
            a[hi] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) tile_sort_pack_kernel(const uint2* __restrict__ ranges,
                                                             unsigned long long* __restrict__ keys, GeomWS g,
                                                             const float* __restrict__ colors_precomp, BinWS b,
                                                             const int32_t* __restrict__ status) {
  __shared__ unsigned long long s_keys[SORT_SMEM_KEYS];
  if (status[1]) return;
  const uint2 range = ranges[blockIdx.x];
  const uint32_t n = range.y - range.x;
  if (n == 0) return;
  const uint32_t tid = threadIdx.x;
  unsigned long long* seg = keys + range.x;
  unsigned long long* a;
  if (n <= SORT_SMEM_KEYS) {
    for (uint32_t i = tid; i < n; i += 256) s_keys[i] = seg[i];
    a = s_keys;
  } else {
    a = seg;  // rare: sort the segment in place in global memory (L2-resident)
  }
  __syncthreads();
  bitonic_sort_any(a, n, tid, 256);
  const float* colors = colors_precomp ? colors_precomp : g.rgb;
  for (uint32_t i = tid; i < n; i += 256) {
    const unsigned long long k = a[i];
    const uint32_t id = (uint32_t)(k & 0xffffffffull);
    const float2 xy = g.means2D[id];
    const float4 co = g.conic_opacity[id];
    // conservative half extents of { alpha >= 1/255 }: power >= -tau, tau = ln(255 o).
    // Margins (0.02 on tau, 1e-4 relative + 0.01 px) dwarf fp32 rounding of the blend.
    float hx, hy;
    const float tau = __logf(255.0f * co.w) + 0.02f;
    const float det = co.x * co.z - co.y * co.y;
    if (!(tau > 0.0f)) {
      hx = hy = -1.0f;  // opacity < 1/255: can never pass the alpha test
    } else if (!(det > 0.0f) || !isfinite(det) || !isfinite(tau)) {
      hx = hy = 3.0e38f;  // degenerate conic: never cull
    } else {
      hx = sqrtf(2.0f * tau * co.z / det) * 1.0001f + 0.01f;
      hy = sqrtf(2.0f * tau * co.x / det) * 1.0001f + 0.01f;
      if (!isfinite(hx) || !isfinite(hy)) hx = hy = 3.0e38f;
    }
    const size_t o = (size_t)range.x + i;
    b.point_list[o] = id;
    b.inst_geo[o] = make_float4(xy.x, xy.y, hx, hy);
    b.inst_attr[2 * o] = co;
    b.inst_attr[2 * o + 1] = make_float4(colors[3 * id], colors[3 * id + 1], colors[3 * id + 2], __uint_as_float(id));
  }
}

// =============================================================== render ====
// One CTA per 16x16 tile, one thread per pixel, warp = 8x4 pixel block
// (semantics of renderCUDA, forward.cu:261-374).
#define RB 256  // records per batch (== reference BLOCK_SIZE staging granularity)

__global__ void __launch_bounds__(256) render_fwd_kernel(const uint2* __restrict__ ranges,
                                                         const float4* __restrict__ inst_geo,
                                                         const float4* __restrict__ inst_attr, int W, int H,
                                                         const float* __restrict__ bg_color,
                                                         float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                                                         float* __restrict__ out_color,
                                                         const int32_t* __restrict__ status) {
  __shared__ __align__(128) float4 s_geo[2][RB];
  __shared__ __align__(128) float4 s_attr[2][2 * RB];
  __shared__ __align__(8) uint64_t s_bar[2];

  const unsigned gx = (W + TILE_X - 1) / TILE_X;
  const unsigned tile = blockIdx.x;
  const unsigned tx = tile % gx, ty = tile / gx;
  const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // warp -> 8x4 block inside the tile; lane -> pixel inside the block
  const int wx0 = tx * TILE_X + (wid & 1) * 8, wy0 = ty * TILE_Y + (wid >> 1) * 4;
  const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const uint32_t pix_id = W * py + px;
  const float pixfx = (float)px, pixfy = (float)py;
  const float bx0 = (float)wx0, bx1 = (float)(wx0 + 7), by0 = (float)wy0, by1 = (float)(wy0 + 3);

  uint2 range = ranges[tile];
  if (status[1]) range = make_uint2(0, 0);
  const int total = range.y - range.x;
  const int rounds = (total + RB - 1) / RB;

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0 && rounds > 0) {
    const uint32_t nb = min(RB, total);
    mbar_expect_tx(&s_bar[0], nb * 48u);
    tma_load_1d(&s_geo[0][0], inst_geo + range.x, nb * 16u, &s_bar[0]);
    tma_load_1d(&s_attr[0][0], inst_attr + 2 * (size_t)range.x, nb * 32u, &s_bar[0]);
  }

  bool done = !inside;
  float T = 1.0f;
  uint32_t last_contributor = 0;
  float C0 = 0.f, C1 = 0.f, C2 = 0.f;

  for (int i = 0; i < rounds; ++i) {
    const int st = i & 1;
    mbar_wait(&s_bar[st], (i >> 1) & 1);
    // block-wide exit vote (forward.cu:309-311); also fences reuse of the other stage
    const int num_done = __syncthreads_count(done);
    if (num_done == TILE_PIX) break;
    const int base = i * RB;
    const int nb = min(RB, total - base);
    if (tid == 0 && i + 1 < rounds) {
      const int nbase = base + RB;
      const uint32_t nn = min(RB, total - nbase);
      mbar_expect_tx(&s_bar[st ^ 1], nn * 48u);
      tma_load_1d(&s_geo[st ^ 1][0], inst_geo + range.x + nbase, nn * 16u, &s_bar[st ^ 1]);
      tma_load_1d(&s_attr[st ^ 1][0], inst_attr + 2 * ((size_t)range.x + nbase), nn * 32u, &s_bar[st ^ 1]);
    }
    if (__all_sync(0xffffffffu, done)) continue;  // whole 8x4 block finished

    // ---- cull this batch against the warp's pixel block
    unsigned keep[RB / 32];
#pragma unroll
    for (int k = 0; k < RB / 32; ++k) {
      const int r = k * 32 + lane;
      bool kp = false;
      if (r < nb) {
        const float4 ge = s_geo[st][r];
        const float ddx = fmaxf(fmaxf(bx0 - ge.x, ge.x - bx1), 0.0f);
        const float ddy = fmaxf(fmaxf(by0 - ge.y, ge.y - by1), 0.0f);
        kp = !(ddx > ge.z || ddy > ge.w);
      }
      keep[k] = __ballot_sync(0xffffffffu, kp);
    }
    // ---- blend the survivors front to back
#pragma unroll
    for (int k = 0; k < RB / 32; ++k) {
      unsigned mask = keep[k];
      if (mask == 0) continue;
      if (__all_sync(0xffffffffu, done)) break;
      while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        const int j = k * 32 + bit;
        const float4 ge = s_geo[st][j];
        const float4 con_o = s_attr[st][2 * j];
        const float4 col = s_attr[st][2 * j + 1];
        if (!done) {
          const float dx = ge.x - pixfx, dy = ge.y - pixfy;
          const float power = -0.5f * (con_o.x * dx * dx + con_o.z * dy * dy) - con_o.y * dx * dy;
          if (!(power > 0.0f)) {
            const float alpha = min(0.99f, con_o.w * expf(power));
            if (!(alpha < 1.0f / 255.0f)) {
              const float test_T = T * (1 - alpha);
              if (test_T < 0.0001f) {
                done = true;
              } else {
                C0 += col.x * alpha * T;
                C1 += col.y * alpha * T;
                C2 += col.z * alpha * T;
                T = test_T;
                last_contributor = base + j + 1;
              }
            }
          }
        }
      }
    }
  }
  if (inside) {
    final_T[pix_id] = T;
    n_contrib[pix_id] = last_contributor;
    const size_t HW = (size_t)H * W;
    out_color[0 * HW + pix_id] = C0 + T * bg_color[0];
    out_color[1 * HW + pix_id] = C1 + T * bg_color[1];
    out_color[2 * HW + pix_id] = C2 + T * bg_color[2];
  }
}

// frustum test only (rasterizer_impl.cu:54-66)
__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                           const float* __restrict__ viewmatrix,
                                                           const float* __restrict__ projmatrix,
                                                           uint8_t* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
  const float3 pv = xform4x3(p, viewmatrix);
  present[idx] = (pv.z <= 0.2f) ? 0 : 1;
}

// ------------------------------------------------------------ launchers ----
cudaError_t launch_forward(const FwdArgs& a, cudaStream_t s) {
  const unsigned gx = (a.W + TILE_X - 1) / TILE_X, gy = (a.H + TILE_Y - 1) / TILE_Y;
  const int T = gx * gy;
  GeomWS g = GeomWS::from((char*)a.geom_ws, a.P);
  ImgWS im = ImgWS::from((char*)a.img_ws, (size_t)a.W * a.H, T);
  BinWS b = BinWS::from((char*)a.binning_ws, (size_t)a.R_cap);
  const float focal_y = a.H / (2.0f * a.tan_fovy);
  const float focal_x = a.W / (2.0f * a.tan_fovx);
  cudaMemsetAsync(im.tile_counts, 0, im.zero_bytes, s);
  if (a.P > 0) {
    g_prof.begin(0, s);
    preprocess_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(
        a.P, a.D, a.M, a.means3D, a.scales, a.scale_modifier, a.rotations, a.opacities, a.shs, a.cov3D_precomp,
        a.colors_precomp, a.viewmatrix, a.projmatrix, a.cam_pos, a.W, a.H, a.tan_fovx, a.tan_fovy, focal_x, focal_y,
        a.radii, g, gx, gy, im.tile_counts, a.prefiltered);
    g_prof.end(0, s);
  }
  g_prof.begin(1, s);
  tile_scan_kernel<<<1, 1024, 0, s>>>(T, im.tile_counts, im.ranges, a.status, (long long)a.R_cap);
  g_prof.end(1, s);
  if (a.P > 0) {
    g_prof.begin(2, s);
    scatter_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(a.P, g, im.ranges, im.tile_fill, b.keys, gx, gy, a.status);
    g_prof.end(2, s);
    g_prof.begin(3, s);
    tile_sort_pack_kernel<<<T, 256, 0, s>>>(im.ranges, b.keys, g, a.colors_precomp, b, a.status);
    g_prof.end(3, s);
  }
  g_prof.begin(4, s);
  render_fwd_kernel<<<T, 256, 0, s>>>(im.ranges, b.inst_geo, b.inst_attr, a.W, a.H, a.background, im.final_T,
                                      im.n_contrib, a.out_color, a.status);
  g_prof.end(4, s);
  return cudaGetLastError();
}

cudaError_t launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                                cudaStream_t s) {
  if (P > 0) mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, proj, present);
  return cudaGetLastError();
}

}  // namespace dgm
