// mlp_gemm.cuh -- bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulator in
// TMEM), the building block of the deformation / appearance MLPs.
//
//   D[M,N] = A[M,K] * B[N,K]^T          A, B row-major bf16 (K contiguous), fp32 accumulate
//
// One CTA (4 warps) owns a 128 x BN output tile: operands are staged global -> shared with 16-byte
// cp.async copies straight into the UMMA core-matrix layout (umma.cuh), 2-stage ring; ONE thread
// issues the tcgen05.mma instructions (M=128, N=BN, K=16) and commits each stage to an mbarrier;
// after the last commit the four warps read their 32 TMEM lanes (tcgen05.ld 32x32b.x32) and run the
// fused epilogue: + bias, ReLU, ReLU-mask (for dZ = dA * [act > 0]), then any of: fp32 store,
// fp32 split-K reduction (red.global.add.v4), bf16 store, transposed bf16 store.
#pragma once
#include "umma.cuh"

namespace dgm {

struct GemmArgs {
  const __nv_bfloat16* A;
  const __nv_bfloat16* B;
  int lda, ldb;        // elements, multiples of 8
  int M, N, K;         // K multiple of 8
  int k_split;         // K elements per blockIdx.z slice (multiple of 64), == K when not split
  const float* bias;   // [N] or null
  int relu;            // apply max(0, .)
  const __nv_bfloat16* mask;  // [M, ld_mask] or null: multiply by (mask > 0)
  int ld_mask;
  float* out_f32;      // [M, ld_f32] or null
  int ld_f32;
  int atomic;          // accumulate into out_f32 with reductions (split-K)
  __nv_bfloat16* out_bf16;    // [M, ld_bf16] or null
  int ld_bf16;
  __nv_bfloat16* out_bf16_t;  // [N, ld_t] or null (transposed copy)
  int ld_t;
};

#define GEMM_BM 128
#define GEMM_BK 64
#define GEMM_STAGES 2  // 2 x (16 + 32) KB at BN = 256: two CTAs per SM, one's epilogue overlaps the other's main loop

template <int BN>
__global__ void __launch_bounds__(128) gemm_tn_kernel(const GemmArgs g) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB per stage
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr int LBO_A = GEMM_BM * 16, LBO_B = BN * 16;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  __shared__ __align__(8) uint64_t s_bar[GEMM_STAGES];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * g.k_split, kend = min(g.K, kbeg + g.k_split);
  const int nk = (kend - kbeg + GEMM_BK - 1) / GEMM_BK;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GEMM_STAGES; ++s) mbar_init(&s_bar[s], 1);
    mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&s_tmem, TMEM_COLS);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t sA = smem_u32(smem), sB = sA + GEMM_STAGES * A_BYTES;

  auto load_chunk = [&](int c) {
    const int st = c % GEMM_STAGES;
    const int k0 = kbeg + c * GEMM_BK;
    // A: 128 rows x 8 k-blocks.  A warp-iteration covers 8 rows x 4 k-blocks (64 B per row from
    // global, 4 shared-memory wavefronts -- the minimum for 512 B).
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int w = it * 4 + warp;
      const int row = (w >> 1) * 8 + (lane & 7), kb = (w & 1) * 4 + (lane >> 3);
      const int gr = m0 + row, gk = k0 + kb * 8;
      const bool ok = gr < g.M && gk < kend;
      const __nv_bfloat16* src = ok ? g.A + (size_t)gr * g.lda + gk : g.A;
      umma::cp_async16(sA + st * A_BYTES + kb * LBO_A + (row >> 3) * 128 + (row & 7) * 16, src, ok ? 16u : 0u);
    }
#pragma unroll
    for (int it = 0; it < BN / 16; ++it) {
      const int w = it * 4 + warp;
      const int row = (w >> 1) * 8 + (lane & 7), kb = (w & 1) * 4 + (lane >> 3);
      const int gr = n0 + row, gk = k0 + kb * 8;
      const bool ok = gr < g.N && gk < kend;
      const __nv_bfloat16* src = ok ? g.B + (size_t)gr * g.ldb + gk : g.B;
      umma::cp_async16(sB + st * B_BYTES + kb * LBO_B + (row >> 3) * 128 + (row & 7) * 16, src, ok ? 16u : 0u);
    }
  };

  constexpr uint32_t IDESC = umma::instr_desc_bf16(GEMM_BM, BN < 16 ? 16 : BN);
  // prologue: chunks 0 .. STAGES-2 in flight (one commit group per chunk, empty groups keep the count uniform)
#pragma unroll
  for (int c = 0; c < GEMM_STAGES - 1; ++c) {
    if (c < nk) load_chunk(c);
    umma::cp_async_commit();
  }
  for (int c = 0; c < nk; ++c) {
    const int cn = c + GEMM_STAGES - 1;  // chunk to prefetch; its stage was last read by the MMAs of chunk c-1
    if (cn < nk) {
      if (c >= 1) mbar_wait(&s_bar[(c - 1) % GEMM_STAGES], ((c - 1) / GEMM_STAGES) & 1);
      load_chunk(cn);
    }
    umma::cp_async_commit();
    umma::cp_async_wait<GEMM_STAGES - 1>();  // chunk c has landed (at most STAGES-1 younger groups in flight)
    umma::fence_smem_to_async();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      const int st = c % GEMM_STAGES;
#pragma unroll
      for (int ks = 0; ks < GEMM_BK / 16; ++ks) {
        const uint64_t da = umma::smem_desc(sA + st * A_BYTES + ks * 2 * LBO_A, LBO_A, 128);
        const uint64_t db = umma::smem_desc(sB + st * B_BYTES + ks * 2 * LBO_B, LBO_B, 128);
        umma::mma_bf16(tmem, da, db, IDESC, (c | ks) != 0);
      }
      umma::commit(&s_bar[st]);
    }
  }
  // ---- epilogue: wait for the last commit (it covers every earlier MMA)
  if (nk > 0) mbar_wait(&s_bar[(nk - 1) % GEMM_STAGES], ((nk - 1) / GEMM_STAGES) & 1);
  umma::fence_after_sync();
  const int row = m0 + warp * 32 + lane;
  const bool row_ok = row < g.M;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    if (n0 + c0 >= g.N) break;  // uniform
    uint32_t r[32];
    if (nk > 0) {
      umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = 0u;
    }
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int n = n0 + c0 + i;
      float x = __uint_as_float(r[i]);
      if (g.bias && n < g.N) x += g.bias[n];
      if (g.relu) x = fmaxf(x, 0.0f);
      v[i] = x;
    }
    if (g.mask && row_ok) {
      const __nv_bfloat16* mrow = g.mask + (size_t)row * g.ld_mask + n0 + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        if (n0 + c0 + i < g.N) {
          const uint4 mm = *reinterpret_cast<const uint4*>(mrow + i);
          const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&mm);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[i + q] = (__bfloat162float(mb[q]) > 0.0f) ? v[i + q] : 0.0f;
        }
      }
    }
    if (g.out_f32 && row_ok) {
      float* dst = g.out_f32 + (size_t)row * g.ld_f32 + n0 + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        if (n0 + c0 + i < g.N) {
          if (g.atomic) red_add_v4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
          else *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
    if (g.out_bf16 && row_ok) {
      __nv_bfloat16* dst = g.out_bf16 + (size_t)row * g.ld_bf16 + n0 + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        if (n0 + c0 + i < g.N) {
          uint4 pk;
          __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int q = 0; q < 4; ++q) p2[q] = __floats2bfloat162_rn(v[i + 2 * q], v[i + 2 * q + 1]);
          *reinterpret_cast<uint4*>(dst + i) = pk;
        }
      }
    }
    if (g.out_bf16_t && row_ok) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int n = n0 + c0 + i;
        if (n < g.N) g.out_bf16_t[(size_t)n * g.ld_t + row] = __float2bfloat16_rn(v[i]);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace dgm
