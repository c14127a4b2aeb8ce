// mlp_gemm.cuh -- the two tensor-core kernels of the deformation / appearance MLPs (tcgen05.mma, bf16
// operands, fp32 accumulators in TMEM), built around ONE global data layout.
//
// Blocked ("K-blocked") bf16 layout.  A matrix with F feature columns (F % 8 == 0) is stored in row
// tiles of R rows:  element (row, f) lives at
//     tile(row / R) * tile_stride  +  ((f / 8) * R + row % R) * 8 + f % 8            [elements]
// i.e. per tile, per 8-feature block ("kb"), R units of 16 bytes.  Activations use R = 128 (one tile =
// one 128-point M tile), weight matrices a single tile with R = number of output rows.
// Why: (1) a K-chunk of a tile (8 kbs) is ONE contiguous piece of memory that is already in the UMMA
// no-swizzle core-matrix order, so an operand stage is filled by a single 1-D bulk copy
// (cp.async.bulk + mbarrier complete_tx) issued by one thread -- no per-thread address arithmetic,
// no tensor maps; (2) the epilogue's stores are fully coalesced (32 lanes = 32 consecutive rows = 512
// contiguous bytes per instruction); (3) the SAME stored activations serve as MN-major operands of
// the weight-gradient GEMM (contraction over the points), so no transposed copies are ever written.
//
//   layer_gemm_kernel   D[rows, N] = A[rows, K] . B[N, K]^T          (forward layers, dZ back-propagation)
//       persistent CTAs (one per SM) over the 128-row tiles; warp-specialised:
//         warp 0    producer: per K-chunk one bulk copy for A and one for B into a 4-stage ring
//         warp 1    issues tcgen05.mma (M = 128, N <= 256, K = 16) into one of TWO TMEM accumulators
//         warps 2-9 epilogue of the previous tile (tcgen05.ld -> bias / ReLU / ReLU bit-mask -> blocked
//                   bf16 store, fp32 store or fp32 reduction) while the next tile is being multiplied
//   dw_gemm_kernel      C[Mf, Nf] += X[rows, Mf]^T . Y[rows, Nf]     (weight gradients)
//       both operands MN-major straight from the blocked activations; each CTA accumulates its share
//       of the row tiles in TMEM and adds its partial result to C once (red.global.add).
#pragma once
#include "umma.cuh"

namespace dgm {

#define ACT_R 128             // rows per activation tile
#define KB_ELEMS (ACT_R * 8)  // elements of one 8-feature block of an activation tile (2 KB)

// view of a blocked activation matrix: pointer to tile 0, elements between tiles, first kb used
struct BlkView {
  const __nv_bfloat16* p;
  size_t tile_stride;
  int kb0;
};

struct LayerArgs {
  BlkView A;                  // [tiles*128, K]
  const __nv_bfloat16* B;     // blocked single tile, R = N rows: [K/8][N][8]
  int K, N;                   // K % 16 == 0, N % 16 == 0, 16 <= N <= 256
  int tiles;                  // 128-row tiles
  const int* tiles_dev;       // optional: device-side cap on `tiles` (decided by an earlier kernel of the stream)
  int rows_valid;             // rows < rows_valid are written to the fp32 output
  const float* bias;          // [N] or null
  int relu;
  const uint32_t* mask_bits;  // null: none; else ReLU mask of a previous output: bit j of word
                              // [(tile*8 + c/32)*128 + row] keeps column c = 32*(c/32) + j
  uint32_t* mask_out;         // null: none; else write this output's (value > 0) bits in that layout
  __nv_bfloat16* out;         // blocked output (N columns from out_kb0) or null
  size_t out_tile_stride;
  int out_kb0;
  float* out_f32;             // row-major [rows, ld_f32] or null; first n_f32 columns
  int ld_f32, n_f32, atomic;  // atomic: accumulate with red.global.add
  // split precision ("bf16x3"): operands are pairs hi + lo of bf16 numbers (lo = the rounding
  // residual of hi), the product is A_hi B_hi + A_lo B_hi + A_hi B_lo accumulated in fp32 -- three
  // passes over K into the same TMEM accumulator, ~16 mantissa bits per operand.  A_lo / B_lo have
  // the geometry of A / B; out_lo (optional) receives the residual of the bf16 output.
  const __nv_bfloat16* A_lo;  // null: single pass
  const __nv_bfloat16* B_lo;
  __nv_bfloat16* out_lo;
};

#define LG_STAGES 4
#define LG_A_BYTES (ACT_R * 64 * 2)  // 16 KB: 128 rows x 64 K
#define LG_B_BYTES (256 * 64 * 2)    // 32 KB: up to 256 rows x 64 K
#define LG_THREADS 320  // producer warp + MMA warp + 8 epilogue warps
#define LG_SMEM (LG_STAGES * (LG_A_BYTES + LG_B_BYTES) + 1024)
// RESIDENT variant (RES = true; split-precision layers whose weight matrix fits 128 KB, i.e. every 256x256 layer):
// measured on the streaming variant, such a layer moved 704 KB per 128-row tile through L2 (the weights
// re-fetched by every tile and pass) -- ~11 TB/s, the L2 slices' throughput limit, while HBM sat at 39 %.  Here
// B_hi is loaded into shared memory ONCE per CTA; the ring carries only what changes per tile, in 16 KB stages:
// one 16-wide K step of A_hi, A_lo and B_lo -- the three products of a K step are issued back to back
// (A_hi B_hi, A_lo B_hi, A_hi B_lo), so A_hi is fetched once, not twice.  L2 traffic per tile: 704 -> 384 KB.
// (Single-pass launches gain nothing from residency -- measured -- and keep the streaming variant.)
#define LGR_B_BYTES 131072
#define LGR_STAGES 5
#define LGR_STAGE_BYTES 16384
#define LGR_SMEM (LGR_B_BYTES + LGR_STAGES * LGR_STAGE_BYTES + 1024)
#define LG_MAX_STAGES 5

template <bool RES>
__global__ void __launch_bounds__(LG_THREADS, 1) layer_gemm_kernel(const LayerArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[LG_MAX_STAGES], s_empty[LG_MAX_STAGES], s_acc_full[2], s_acc_empty[2], s_bres;
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[256];
  constexpr int NST = RES ? LGR_STAGES : LG_STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 1 KB aligned operand ring
  const uint32_t sA = ring, sB = ring + LG_STAGES * LG_A_BYTES;  // streaming variant
  const uint32_t sBres = ring, sStage = ring + LGR_B_BYTES;      // resident variant

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    mbar_init(&s_bres, 1);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_acc_full[b], 1);
      mbar_init(&s_acc_empty[b], 256);
    }
    mbar_fence_init();
  }
  for (int i = tid; i < 256; i += LG_THREADS) s_bias[i] = (g.bias && i < g.N) ? g.bias[i] : 0.f;
  if (warp == 1) umma::tmem_alloc(&s_tmem, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;

  const int nkb = g.K >> 3;            // 8-wide k blocks
  const int nchunks = (nkb + 7) >> 3;  // K chunks of (up to) 64
  const int npasses = g.A_lo ? 3 : 1;
  const int n_tiles = g.tiles_dev ? min(g.tiles, *g.tiles_dev) : g.tiles;
  const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    if (RES && lane == 0) {
      if (my_tiles > 0) {  // the resident weights, once
        const uint32_t bytes = (uint32_t)nkb * g.N * 16;
        mbar_expect_tx(&s_bres, bytes);
        for (uint32_t off = 0; off < bytes; off += 32768u)
          tma_load_1d_u32(sBres + off, reinterpret_cast<const uint8_t*>(g.B) + off, min(32768u, bytes - off), &s_bres);
      }
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const size_t a_off = (size_t)tile * g.A.tile_stride + (size_t)g.A.kb0 * KB_ELEMS;
        {
          const uint32_t bl_bytes = (uint32_t)g.N * 32;  // two k blocks of B_lo
          for (int c = 0; c < (nkb >> 1); ++c, ++it) {
            const int st = it % NST;
            mbar_wait(&s_empty[st], ((it / NST) & 1) ^ 1);
            mbar_expect_tx(&s_full[st], 2 * 4096u + bl_bytes);
            const uint32_t dst = sStage + st * LGR_STAGE_BYTES;
            tma_load_1d_u32(dst, g.A.p + a_off + (size_t)c * 2 * KB_ELEMS, 4096u, &s_full[st]);
            tma_load_1d_u32(dst + 4096u, g.A_lo + a_off + (size_t)c * 2 * KB_ELEMS, 4096u, &s_full[st]);
            tma_load_1d_u32(dst + 8192u, g.B_lo + (size_t)c * 2 * g.N * 8, bl_bytes, &s_full[st]);
          }
        }
      }
    } else if (!RES && lane == 0) {
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const size_t a_off = (size_t)tile * g.A.tile_stride + (size_t)g.A.kb0 * KB_ELEMS;
        for (int pass = 0; pass < npasses; ++pass) {
          // pass 0: A_hi B_hi, pass 1: A_lo B_hi, pass 2: A_hi B_lo
          const __nv_bfloat16* a_tile = ((pass == 1) ? g.A_lo : g.A.p) + a_off;
          const __nv_bfloat16* b_mat = (pass == 2) ? g.B_lo : g.B;
          for (int c = 0; c < nchunks; ++c, ++it) {
            const int st = it % LG_STAGES;
            const int kbs = min(8, nkb - 8 * c);
            mbar_wait(&s_empty[st], ((it / LG_STAGES) & 1) ^ 1);
            const uint32_t a_bytes = kbs * (ACT_R * 16), b_bytes = kbs * g.N * 16;
            mbar_expect_tx(&s_full[st], a_bytes + b_bytes);
            tma_load_1d_u32(sA + st * LG_A_BYTES, a_tile + (size_t)c * 8 * KB_ELEMS, a_bytes, &s_full[st]);
            tma_load_1d_u32(sB + st * LG_B_BYTES, b_mat + (size_t)c * 8 * g.N * 8, b_bytes, &s_full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma::instr_desc_bf16(128, g.N);
      const uint32_t lbo_b = g.N * 16;
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int buf = i & 1;
        mbar_wait(&s_acc_empty[buf], ((i >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        umma::fence_after_sync();
        const uint32_t d_tmem = tmem + buf * 256;
        if (RES) {
          if (i == 0) mbar_wait(&s_bres, 0);
          {
            for (int c = 0; c < (nkb >> 1); ++c, ++it) {
              const int st = it % NST;
              mbar_wait(&s_full[st], (it / NST) & 1);
              umma::fence_after_sync();
              const uint32_t base = sStage + st * LGR_STAGE_BYTES;
              const uint64_t da_hi = umma::smem_desc(base, ACT_R * 16, 128);
              const uint64_t da_lo = umma::smem_desc(base + 4096u, ACT_R * 16, 128);
              const uint64_t db_hi = umma::smem_desc(sBres + (uint32_t)(2 * c) * lbo_b, lbo_b, 128);
              const uint64_t db_lo = umma::smem_desc(base + 8192u, lbo_b, 128);
              umma::mma_bf16(d_tmem, da_hi, db_hi, idesc, c != 0);
              umma::mma_bf16(d_tmem, da_lo, db_hi, idesc, true);
              umma::mma_bf16(d_tmem, da_hi, db_lo, idesc, true);
              umma::commit(&s_empty[st]);
            }
          }
          umma::commit(&s_acc_full[buf]);
          continue;
        }
        for (int pc = 0; pc < npasses * nchunks; ++pc, ++it) {
          const int c = pc % nchunks;
          const int st = it % LG_STAGES;
          const int kbs = min(8, nkb - 8 * c);
          mbar_wait(&s_full[st], (it / LG_STAGES) & 1);
          umma::fence_after_sync();
          for (int ks = 0; ks < (kbs >> 1); ++ks) {
            const uint64_t da = umma::smem_desc(sA + st * LG_A_BYTES + ks * 2 * (ACT_R * 16), ACT_R * 16, 128);
            const uint64_t db = umma::smem_desc(sB + st * LG_B_BYTES + ks * 2 * lbo_b, lbo_b, 128);
            umma::mma_bf16(d_tmem, da, db, idesc, (pc | ks) != 0);
          }
          umma::commit(&s_empty[st]);  // the stage is reusable once these MMAs have read it
        }
        umma::commit(&s_acc_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // two warps per TMEM lane group (a warp may only touch lanes 32*(warp%4)..+31): warps 2-5 take
    // the 32-column slices 0-3, warps 6-9 the slices 4-7, so two tcgen05.ld / store streams overlap
    const int lg = warp & 3;       // TMEM lane group this warp may access
    const int r = lg * 32 + lane;  // row inside the tile
    const int q0 = (warp >= 6) ? 4 : 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int buf = i & 1;
      // the ReLU-mask words of this row are fetched while the tile is still being multiplied
      uint32_t mb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        mb[q] = (g.mask_bits && (q0 + q) * 32 < g.N) ? g.mask_bits[((size_t)tile * 8 + q0 + q) * ACT_R + r]
                                                     : 0xffffffffu;
      mbar_wait(&s_acc_full[buf], (i >> 1) & 1);
      umma::fence_after_sync();
      const size_t row = (size_t)tile * ACT_R + r;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const int q = q0 + qq;
        const int c0 = q * 32;
        if (c0 < g.N) {  // uniform per warp
          uint32_t raw[32];
          umma::tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 256 + c0), raw);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(raw[j]) + s_bias[c0 + j];
            if (g.relu) x = fmaxf(x, 0.0f);
            v[j] = x;
          }
          if (g.mask_bits) {  // uniform
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((mb[qq] >> j) & 1u) ? v[j] : 0.0f;
          }
          if (g.mask_out) {   // uniform
            uint32_t pos = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) pos |= (v[j] > 0.0f ? 1u : 0u) << j;
            g.mask_out[((size_t)tile * 8 + q) * ACT_R + r] = pos;
          }
          const int nunits = min(4, (g.N - c0) >> 3);  // 8-column units of this 32-column slice
          if (g.out) {
            __nv_bfloat16* ot = g.out + (size_t)tile * g.out_tile_stride;
            __nv_bfloat16* ol = g.out_lo ? g.out_lo + (size_t)tile * g.out_tile_stride : nullptr;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (u < nunits) {
                uint4 pk;
                __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int k = 0; k < 4; ++k) p2[k] = __floats2bfloat162_rn(v[8 * u + 2 * k], v[8 * u + 2 * k + 1]);
                const size_t off = ((size_t)(g.out_kb0 + (c0 >> 3) + u) * ACT_R + r) * 8;
                *reinterpret_cast<uint4*>(ot + off) = pk;
                if (ol) {  // uniform: the rounding residual as a second bf16 number
                  uint4 pl;
                  __nv_bfloat162* l2 = reinterpret_cast<__nv_bfloat162*>(&pl);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 hi = __bfloat1622float2(p2[k]);
                    l2[k] = __floats2bfloat162_rn(v[8 * u + 2 * k] - hi.x, v[8 * u + 2 * k + 1] - hi.y);
                  }
                  *reinterpret_cast<uint4*>(ol + off) = pl;
                }
              }
            }
          }
          if (g.out_f32 && row < (size_t)g.rows_valid) {
            float* dst = g.out_f32 + row * g.ld_f32 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (c0 + j < g.n_f32) {
                if (g.atomic) red_add_v4(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                else *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
            }
          }
        }
      }
      umma::fence_before_sync();
      mbar_arrive(&s_acc_empty[buf]);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
struct DwArgs {
  BlkView X;      // rows x 256 features: CTA (blockIdx.y) takes features [X.kb0*8 + 128*blockIdx.y, +128)
  BlkView Y;      // rows x N features from Y.kb0
  int N;          // multiple of 16, <= 256
  int tiles;
  const int* tiles_dev;  // optional device-side cap on `tiles`
  float* C;       // fp32, leading dimension ld
  int ld;
  int transpose;  // 0: C[m, n] (m = X feature, n = Y feature);  1: C[n, m]
  int m_valid;    // X features >= m_valid are not written
  int n_valid;
  // optional bias gradients from the operand tiles already staged in shared memory (the four epilogue warps
  // are idle during the main loop): colsum_x[f] += sum over rows of X[:, f] (f relative to the view) for this
  // CTA's 128 features, colsum_y[f] += sum over rows of Y[:, f] (N <= 128; CTAs with blockIdx.y == 0 only)
  float* colsum_x;
  float* colsum_y;
};

#define DW_STAGES 2
#define DW_A_BYTES (128 * ACT_R * 2)  // 32 KB: 128 features x 128 rows
#define DW_B_BYTES (256 * ACT_R * 2)  // 64 KB
#define DW_THREADS 192
#define DW_SMEM (DW_STAGES * (DW_A_BYTES + DW_B_BYTES) + 1024)

__global__ void __launch_bounds__(DW_THREADS, 1) dw_gemm_kernel(const DwArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[DW_STAGES], s_empty[DW_STAGES], s_acc_full;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = ring, sB = ring + DW_STAGES * DW_A_BYTES;
  const bool sum_x = g.colsum_x != nullptr, sum_y = g.colsum_y != nullptr && blockIdx.y == 0 && g.N <= 128;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < DW_STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      // a stage is free once the MMAs have consumed it AND (column sums) the 128 epilogue threads have read it
      mbar_init(&s_empty[s], (sum_x || sum_y) ? 1 + 128 : 1);
    }
    mbar_init(&s_acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) umma::tmem_alloc(&s_tmem, 256);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int n_tiles = g.tiles_dev ? min(g.tiles, *g.tiles_dev) : g.tiles;
  const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int m0 = blockIdx.y * 128;  // first X feature of this CTA

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int st = i % DW_STAGES;
        mbar_wait(&s_empty[st], ((i / DW_STAGES) & 1) ^ 1);
        const uint32_t a_bytes = 16 * (ACT_R * 16), b_bytes = (g.N >> 3) * (ACT_R * 16);
        mbar_expect_tx(&s_full[st], a_bytes + b_bytes);
        tma_load_1d_u32(sA + st * DW_A_BYTES,
                        g.X.p + (size_t)tile * g.X.tile_stride + (size_t)(g.X.kb0 + (m0 >> 3)) * KB_ELEMS, a_bytes,
                        &s_full[st]);
        tma_load_1d_u32(sB + st * DW_B_BYTES, g.Y.p + (size_t)tile * g.Y.tile_stride + (size_t)g.Y.kb0 * KB_ELEMS,
                        b_bytes, &s_full[st]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // both operands MN-major: core matrix = 8 rows (K) x 8 features; K groups 128 B apart (LBO),
      // feature groups one kb = 2 KB apart (SBO)
      const uint32_t idesc = umma::instr_desc_bf16(128, g.N, 1, 1);
      for (int i = 0; i < my_tiles; ++i) {
        const int st = i % DW_STAGES;
        mbar_wait(&s_full[st], (i / DW_STAGES) & 1);
        umma::fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < ACT_R / 16; ++ks) {
          const uint64_t da = umma::smem_desc(sA + st * DW_A_BYTES + ks * 256, 128, ACT_R * 16);
          const uint64_t db = umma::smem_desc(sB + st * DW_B_BYTES + ks * 256, 128, ACT_R * 16);
          umma::mma_bf16(tmem, da, db, idesc, (i | ks) != 0);
        }
        umma::commit(&s_empty[st]);
      }
      if (my_tiles > 0) umma::commit(&s_acc_full);
    }
  } else if (my_tiles > 0) {
    if (sum_x || sum_y) {
      // ---- column sums of the staged operand tiles while the tensor core multiplies them
      const int e = tid - 64;                // 0..127
      const int kb = e >> 3, sub = e & 7;    // 8 threads per 8-feature block, 16 rows each (conflict-free LDS.128)
      float ax[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ay[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const bool do_y = sum_y && kb < (g.N >> 3);
      const uint8_t* ring_p = reinterpret_cast<const uint8_t*>(smem_raw) + (ring - smem_u32(smem_raw));
      for (int i = 0; i < my_tiles; ++i) {
        const int st = i % DW_STAGES;
        mbar_wait(&s_full[st], (i / DW_STAGES) & 1);
        const uint8_t* pa = ring_p + st * DW_A_BYTES + kb * (ACT_R * 16);
        const uint8_t* pb = ring_p + DW_STAGES * DW_A_BYTES + st * DW_B_BYTES + kb * (ACT_R * 16);
#pragma unroll 4
        for (int r = sub; r < ACT_R; r += 8) {
          if (sum_x) {
            const uint4 v = *reinterpret_cast<const uint4*>(pa + r * 16);
            const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __bfloat1622float2(p2[q]);
              ax[2 * q] += f.x;
              ax[2 * q + 1] += f.y;
            }
          }
          if (do_y) {
            const uint4 v = *reinterpret_cast<const uint4*>(pb + r * 16);
            const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __bfloat1622float2(p2[q]);
              ay[2 * q] += f.x;
              ay[2 * q + 1] += f.y;
            }
          }
        }
        mbar_arrive(&s_empty[st]);
      }
      // the 8 threads of a feature block are consecutive lanes
#pragma unroll
      for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          ax[q] += __shfl_xor_sync(0xffffffffu, ax[q], o);
          ay[q] += __shfl_xor_sync(0xffffffffu, ay[q], o);
        }
      }
      if (sub == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (sum_x) atomicAdd(&g.colsum_x[m0 + kb * 8 + q], ax[q]);
          if (do_y) atomicAdd(&g.colsum_y[kb * 8 + q], ay[q]);
        }
      }
    }
    const int lg = warp & 3;
    const int m = m0 + lg * 32 + lane;  // X feature = accumulator row
    mbar_wait(&s_acc_full, 0);
    umma::fence_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < g.N; c0 += 32) {
      uint32_t raw[32];
      umma::tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, raw);
      if (m < g.m_valid) {
        if (!g.transpose) {
          float* dst = g.C + (size_t)m * g.ld + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + j < g.n_valid)
              red_add_v4(dst + j, __uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]),
                         __uint_as_float(raw[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < g.n_valid) atomicAdd(g.C + (size_t)(c0 + j) * g.ld + m, __uint_as_float(raw[j]));
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc(tmem, 256);
  }
}

}  // namespace dgm
