// mlp.cu -- deformation / appearance MLPs on the tcgen05 tensor cores.
// (first part: the GEMM building block and its C-ABI test entry; the network follows below)
#include "mlp_gemm.cuh"
#include "mlp_kernels.h"

namespace dgm {

template <int BN>
static cudaError_t launch_gemm_bn(const GemmArgs& g, cudaStream_t s) {
  static bool attr = false;
  constexpr int smem = GEMM_STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2);
  if (!attr) {
    cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  const int splits = (g.K + g.k_split - 1) / g.k_split;
  dim3 grid((g.M + GEMM_BM - 1) / GEMM_BM, (g.N + BN - 1) / BN, splits);
  gemm_tn_kernel<BN><<<grid, 128, smem, s>>>(g);
  return cudaGetLastError();
}

cudaError_t launch_gemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return cudaSuccess;
  if (g.N <= 32) return launch_gemm_bn<32>(g, s);
  if (g.N <= 64) return launch_gemm_bn<64>(g, s);
  if (g.N <= 128) return launch_gemm_bn<128>(g, s);
  return launch_gemm_bn<256>(g, s);
}

}  // namespace dgm

namespace dgm {

// =====================================================================================
// The networks (dgmesh/utils/time_utils.py): DeformNetwork / DeformNetworkNormal /
// DeformNetworkNormalSep / AppearanceNetwork share one trunk --
//   x_emb = pe(x, 10) [63], t_emb = pe(t, 6) [13] -> timenet 13->256->30 (is_blender)
//                            or pe(t, 10) [21]                       (otherwise)
//   h = [x_emb, t_emb] -> 8 x (Linear + ReLU), width 256, [x_emb, t_emb] re-concatenated after
//   layer 4 (time_utils.py:178-188) -> linear heads (13 / 3 / 10 outputs, sigmoid for colour)
// -- and run here as a chain of tcgen05 GEMMs with fused bias/ReLU epilogues over bf16
// activations (fp32 accumulation).  Activation layout (Pp = P rounded up to 128):
//   A5 [Pp, 352] : cols 0..62 x_emb | 63: 0 | 64..64+in_t-1 t features | ..95: 0 | 96..351 h4
//   (layer 0 reads A5[:, :96]; the skip layer reads all 352 columns: no concat copy)
// For training every activation is also stored TRANSPOSED ([features, Pp]) because the weight
// gradient dW = dZ^T . H is a GEMM whose contraction runs over the points.
// =====================================================================================

#define XE 63      // x embedding width
#define TCOL 64    // first column of the time features
#define K0 96      // padded width of [x_emb, t features]
#define K5 352     // skip layer input width
#define WID 256

struct MlpBufs {
  __nv_bfloat16 *A5, *T0, *T1, *H[8];        // H[4] aliases A5 + 96 (ld 352)
  __nv_bfloat16 *A5T, *T0T, *T1T, *HT[8];    // transposed copies (training only); HT[4] = A5T + 96*Pp
  __nv_bfloat16 *dZ[2], *dZT[2], *dZh, *dZhT, *dZt1, *dZt1T;
  float* dE;                                  // [Pp, 96] gradient w.r.t. the embedded inputs
  int Pp;
  static MlpBufs carve_all(char* base, int P, int train, size_t* bytes) {
    char* p = base;
    MlpBufs b;
    const size_t Pp = ((size_t)P + 127) / 128 * 128;
    b.Pp = (int)Pp;
    b.A5 = carve<__nv_bfloat16>(p, Pp * K5);
    b.T0 = carve<__nv_bfloat16>(p, Pp * 16);
    b.T1 = carve<__nv_bfloat16>(p, Pp * WID);
    if (train) {
      for (int l = 0; l < 8; ++l) b.H[l] = (l == 4) ? b.A5 + K0 : carve<__nv_bfloat16>(p, Pp * WID);
      b.A5T = carve<__nv_bfloat16>(p, Pp * K5);
      b.T0T = carve<__nv_bfloat16>(p, Pp * 16);
      b.T1T = carve<__nv_bfloat16>(p, Pp * WID);
      for (int l = 0; l < 8; ++l) b.HT[l] = (l == 4) ? b.A5T + (size_t)K0 * Pp : carve<__nv_bfloat16>(p, Pp * WID);
      for (int i = 0; i < 2; ++i) {
        b.dZ[i] = carve<__nv_bfloat16>(p, Pp * WID);
        b.dZT[i] = carve<__nv_bfloat16>(p, Pp * WID);
      }
      b.dZh = carve<__nv_bfloat16>(p, Pp * 16);
      b.dZhT = carve<__nv_bfloat16>(p, Pp * 16);
      b.dZt1 = carve<__nv_bfloat16>(p, Pp * 32);
      b.dZt1T = carve<__nv_bfloat16>(p, Pp * 32);
      b.dE = carve<float>(p, Pp * K0);
    } else {
      __nv_bfloat16* ping = carve<__nv_bfloat16>(p, Pp * WID);
      __nv_bfloat16* pong = carve<__nv_bfloat16>(p, Pp * WID);
      for (int l = 0; l < 8; ++l) b.H[l] = (l == 4) ? b.A5 + K0 : ((l & 1) ? pong : ping);
      b.A5T = b.T0T = b.T1T = nullptr;
      for (int l = 0; l < 8; ++l) b.HT[l] = nullptr;
      b.dZ[0] = b.dZ[1] = b.dZT[0] = b.dZT[1] = b.dZh = b.dZhT = b.dZt1 = b.dZt1T = nullptr;
      b.dE = nullptr;
    }
    if (bytes) *bytes = size_t(p - base) + 128;
    return b;
  }
};

// positional encodings (time_utils.py:8-55): [v, sin(v 2^0), cos(v 2^0), ..., sin(v 2^(L-1)), cos(v 2^(L-1))]
// One thread per point; writes bf16 rows of A5 (x part + direct time features) and T0 (timenet input).
__global__ void __launch_bounds__(256) pe_kernel(int P, int Pp, const float* __restrict__ x,
                                                 const float* __restrict__ t, int has_timenet, int t_freqs,
                                                 __nv_bfloat16* __restrict__ A5, __nv_bfloat16* __restrict__ T0,
                                                 __nv_bfloat16* __restrict__ A5T, __nv_bfloat16* __restrict__ T0T) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Pp) return;
  const bool ok = p < P;
  float e[K0];
#pragma unroll
  for (int i = 0; i < K0; ++i) e[i] = 0.f;
  float te[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) te[i] = 0.f;
  if (ok) {
    const float v[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
#pragma unroll
    for (int c = 0; c < 3; ++c) e[c] = v[c];
#pragma unroll
    for (int f = 0; f < 10; ++f) {
      const float fr = (float)(1 << f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s, co;
        sincosf(v[c] * fr, &s, &co);
        e[3 + 6 * f + c] = s;
        e[3 + 6 * f + 3 + c] = co;
      }
    }
    const float tv = t[p];
    te[0] = tv;
    for (int f = 0; f < t_freqs; ++f) {
      float s, co;
      sincosf(tv * (float)(1 << f), &s, &co);
      te[1 + 2 * f] = s;
      te[2 + 2 * f] = co;
    }
    if (!has_timenet) {
#pragma unroll
      for (int i = 0; i < 21; ++i) e[TCOL + i] = te[i];
    }
  }
  __nv_bfloat16* row = A5 + (size_t)p * K5;
#pragma unroll
  for (int i = 0; i < K0; i += 8) {
    if (has_timenet && i >= TCOL) break;  // the time-feature columns are written by the timenet GEMM
    uint4 pk;
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int q = 0; q < 4; ++q) p2[q] = __floats2bfloat162_rn(e[i + 2 * q], e[i + 2 * q + 1]);
    *reinterpret_cast<uint4*>(row + i) = pk;
  }
  if (has_timenet) {
    uint4 pk[2];
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
    for (int q = 0; q < 8; ++q) p2[q] = __floats2bfloat162_rn(te[2 * q], te[2 * q + 1]);
    reinterpret_cast<uint4*>(T0 + (size_t)p * 16)[0] = pk[0];
    reinterpret_cast<uint4*>(T0 + (size_t)p * 16)[1] = pk[1];
  }
  if (A5T) {
    const int n_x = has_timenet ? TCOL : K0;
    for (int i = 0; i < n_x; ++i) A5T[(size_t)i * Pp + p] = __float2bfloat16_rn(e[i]);
    A5T[(size_t)94 * Pp + p] = __float2bfloat16_rn(0.f);  // pad rows the timenet GEMM never writes
    A5T[(size_t)95 * Pp + p] = __float2bfloat16_rn(0.f);
    if (has_timenet)
      for (int i = 0; i < 16; ++i) T0T[(size_t)i * Pp + p] = __float2bfloat16_rn(te[i]);
  }
}

// gradient of the heads' pre-activation: dZh = g (* y (1 - y) for the sigmoid colour head), bf16 + transposed
__global__ void __launch_bounds__(256) head_grad_kernel(int P, int Pp, int n_out, int sigmoid,
                                                        const float* __restrict__ g, const float* __restrict__ y,
                                                        __nv_bfloat16* __restrict__ dZh,
                                                        __nv_bfloat16* __restrict__ dZhT) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Pp) return;
  for (int o = 0; o < 16; ++o) {
    float v = 0.f;
    if (p < P && o < n_out) {
      v = g[(size_t)p * 16 + o];
      if (sigmoid) {
        const float yy = y[(size_t)p * 16 + o];
        v *= yy * (1.f - yy);
      }
    }
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    dZh[(size_t)p * 16 + o] = b;
    dZhT[(size_t)o * Pp + p] = b;
  }
}

__global__ void __launch_bounds__(256) sigmoid_kernel(int n, float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = 1.f / (1.f + expf(-y[i]));
}

// column sums of a bf16 matrix [rows, ld] -> db[ncols] (fp32 atomics).  CTA = 256-row slab; a thread
// owns 8 adjacent columns (one 16-byte load per row) and every (256 / column-groups)-th row; partial
// sums meet in shared memory, one atomic per column per CTA.  ncols multiple of 8, <= 256.
#define COLSUM_ROWS 256
__global__ void __launch_bounds__(256) colsum_kernel(int rows, int ncols, int ld, const __nv_bfloat16* __restrict__ Z,
                                                     float* __restrict__ db) {
  __shared__ float s_sum[256];
  const int groups = ncols >> 3;                 // column groups of 8
  const int lanes = 256 / groups;                // row lanes per column group
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  s_sum[threadIdx.x] = 0.f;
  __syncthreads();
  if (rl < lanes) {
    const int r0 = blockIdx.x * COLSUM_ROWS, r1 = min(rows, r0 + COLSUM_ROWS);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = r0 + rl; r < r1; r += lanes) {
      const uint4 v = *reinterpret_cast<const uint4*>(Z + (size_t)r * ld + cg * 8);
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(p2[q]);
        acc[2 * q] += f.x;
        acc[2 * q + 1] += f.y;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(&s_sum[cg * 8 + q], acc[q]);
  }
  __syncthreads();
  if ((int)threadIdx.x < ncols) atomicAdd(&db[threadIdx.x], s_sum[threadIdx.x]);
}

// dE[:, 64:64+30] (fp32) -> dZt1 [Pp, 32] bf16 (+ transposed)
__global__ void __launch_bounds__(256) tfeat_grad_kernel(int Pp, int n_t, const float* __restrict__ dE,
                                                         __nv_bfloat16* __restrict__ dZ, __nv_bfloat16* __restrict__ dZT) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Pp) return;
  for (int o = 0; o < 32; ++o) {
    const __nv_bfloat16 b = __float2bfloat16_rn(o < n_t ? dE[(size_t)p * K0 + TCOL + o] : 0.f);
    dZ[(size_t)p * 32 + o] = b;
    dZT[(size_t)o * Pp + p] = b;
  }
}

// dx = dE . d pe(x) / dx
__global__ void __launch_bounds__(256) pe_backward_kernel(int P, const float* __restrict__ x,
                                                          const float* __restrict__ dE, float* __restrict__ dx) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* g = dE + (size_t)p * K0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = x[3 * p + c];
    float acc = g[c];
#pragma unroll
    for (int f = 0; f < 10; ++f) {
      const float fr = (float)(1 << f);
      float s, co;
      sincosf(v * fr, &s, &co);
      acc += fr * (g[3 + 6 * f + c] * co - g[3 + 6 * f + 3 + c] * s);
    }
    dx[3 * p + c] = acc;
  }
}

static GemmArgs gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K) {
  GemmArgs g = {};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.k_split = K;
  return g;
}

#define CK(call)                        \
  do {                                  \
    cudaError_t e_ = (call);            \
    if (e_ != cudaSuccess) return e_;   \
  } while (0)

cudaError_t launch_mlp_forward(const DglNet& n, int P, const float* x, const float* t, float* out, int train,
                               void* ws, cudaStream_t s) {
  MlpBufs b = MlpBufs::carve_all((char*)ws, P, train, nullptr);
  const int Pp = b.Pp;
  typedef const __nv_bfloat16* CB;
  pe_kernel<<<(Pp + 255) / 256, 256, 0, s>>>(P, Pp, x, t, n.has_timenet, n.has_timenet ? 6 : 10, b.A5, b.T0,
                                             train ? b.A5T : nullptr, train ? b.T0T : nullptr);
  if (n.has_timenet) {
    GemmArgs g = gemm(b.T0, 16, (CB)n.Wt0, 16, Pp, WID, 16);
    g.bias = n.bt0; g.relu = 1; g.out_bf16 = b.T1; g.ld_bf16 = WID; g.out_bf16_t = b.T1T; g.ld_t = Pp;
    CK(launch_gemm(g, s));
    g = gemm(b.T1, WID, (CB)n.Wt1, WID, Pp, n.in_t, WID);
    g.bias = n.bt1; g.out_bf16 = b.A5 + TCOL; g.ld_bf16 = K5;
    g.out_bf16_t = train ? b.A5T + (size_t)TCOL * Pp : nullptr; g.ld_t = Pp;
    CK(launch_gemm(g, s));
  }
  for (int l = 0; l < 8; ++l) {
    const __nv_bfloat16* A = (l == 0 || l == 5) ? b.A5 : b.H[l - 1];
    const int lda = (l == 0 || l == 5) ? K5 : WID, K = (l == 0) ? K0 : (l == 5 ? K5 : WID);
    GemmArgs g = gemm(A, lda, (CB)n.W[l], K, Pp, WID, K);
    g.bias = n.b[l]; g.relu = 1;
    g.out_bf16 = b.H[l]; g.ld_bf16 = (l == 4) ? K5 : WID;
    g.out_bf16_t = b.HT[l]; g.ld_t = Pp;
    CK(launch_gemm(g, s));
  }
  GemmArgs g = gemm(b.H[7], WID, (CB)n.Wh, WID, P, n.n_out, WID);
  g.bias = n.bh; g.out_f32 = out; g.ld_f32 = 16;
  CK(launch_gemm(g, s));
  if (n.sigmoid_out) sigmoid_kernel<<<(P * 16 + 255) / 256, 256, 0, s>>>(P * 16, out);
  return cudaGetLastError();
}

// weight gradient: dW[out, Kin] += dZT[out, Pp] . HT[Kin, Pp]^T   (split over the points)
static cudaError_t dw_gemm(const __nv_bfloat16* dZT, int M, const __nv_bfloat16* HT, int N, int Pp, float* dW, int ldw,
                           cudaStream_t s) {
  GemmArgs g = gemm(dZT, Pp, HT, Pp, M, N, Pp);
  const int tiles = ((M + 127) / 128) * ((N + 255) / 256);
  int splits = max(1, 296 / tiles);
  g.k_split = max(64, ((Pp + splits - 1) / splits + 63) / 64 * 64);
  g.out_f32 = dW; g.ld_f32 = ldw; g.atomic = 1;
  return launch_gemm(g, s);
}

cudaError_t launch_mlp_backward(const DglNet& n, int P, const float* x, const float* out, const float* g_out,
                                void* ws, const DglGrads& gr, float* dx, cudaStream_t s) {
  MlpBufs b = MlpBufs::carve_all((char*)ws, P, 1, nullptr);
  const int Pp = b.Pp;
  typedef const __nv_bfloat16* CB;
  const int rb = (Pp + COLSUM_ROWS - 1) / COLSUM_ROWS;
  cudaMemsetAsync(b.dE, 0, sizeof(float) * (size_t)Pp * K0, s);
  for (int l = 0; l < 8; ++l) {
    const int K = (l == 0) ? K0 : (l == 5 ? K5 : WID);
    cudaMemsetAsync(gr.dW[l], 0, sizeof(float) * WID * K, s);
    cudaMemsetAsync(gr.db[l], 0, sizeof(float) * WID, s);
  }
  cudaMemsetAsync(gr.dWh, 0, sizeof(float) * 16 * WID, s);
  cudaMemsetAsync(gr.dbh, 0, sizeof(float) * 16, s);
  if (n.has_timenet) {
    cudaMemsetAsync(gr.dWt0, 0, sizeof(float) * WID * 16, s);
    cudaMemsetAsync(gr.dbt0, 0, sizeof(float) * WID, s);
    cudaMemsetAsync(gr.dWt1, 0, sizeof(float) * 32 * WID, s);
    cudaMemsetAsync(gr.dbt1, 0, sizeof(float) * 32, s);
  }
  // heads
  head_grad_kernel<<<(Pp + 255) / 256, 256, 0, s>>>(P, Pp, n.n_out, n.sigmoid_out, g_out, out, b.dZh, b.dZhT);
  CK(dw_gemm(b.dZhT, 16, b.HT[7], WID, Pp, gr.dWh, WID, s));
  colsum_kernel<<<rb, 256, 0, s>>>(Pp, 16, 16, b.dZh, gr.dbh);
  int cur = 0;
  {
    GemmArgs g = gemm(b.dZh, 16, (CB)n.WhT, 16, Pp, WID, 16);  // dH7 = dZh . Wh, masked by ReLU'(H7)
    g.mask = b.H[7]; g.ld_mask = WID;
    g.out_bf16 = b.dZ[cur]; g.ld_bf16 = WID; g.out_bf16_t = b.dZT[cur]; g.ld_t = Pp;
    CK(launch_gemm(g, s));
  }
  for (int l = 7; l >= 0; --l) {
    // dZ[cur] = dL/d(pre-activation of layer l)
    const int K = (l == 0) ? K0 : (l == 5 ? K5 : WID);
    const __nv_bfloat16* HinT = (l == 0 || l == 5) ? b.A5T : b.HT[l - 1];
    CK(dw_gemm(b.dZT[cur], WID, HinT, K, Pp, gr.dW[l], K, s));
    colsum_kernel<<<rb, 256, 0, s>>>(Pp, WID, WID, b.dZ[cur], gr.db[l]);
    if (l == 0 || l == 5) {  // gradient w.r.t. the embedded inputs (no ReLU in front of them)
      GemmArgs g = gemm(b.dZ[cur], WID, (CB)n.WT[l], WID, Pp, K0, WID);
      g.out_f32 = b.dE; g.ld_f32 = K0; g.atomic = 1;
      CK(launch_gemm(g, s));
    }
    if (l > 0) {
      const int off = (l == 5) ? K0 : 0;  // rows of W5^T that belong to h4
      GemmArgs g = gemm(b.dZ[cur], WID, (CB)n.WT[l] + (size_t)off * WID, WID, Pp, WID, WID);
      g.mask = b.H[l - 1]; g.ld_mask = (l - 1 == 4) ? K5 : WID;
      g.out_bf16 = b.dZ[cur ^ 1]; g.ld_bf16 = WID; g.out_bf16_t = b.dZT[cur ^ 1]; g.ld_t = Pp;
      CK(launch_gemm(g, s));
      cur ^= 1;
    }
  }
  if (n.has_timenet) {
    tfeat_grad_kernel<<<(Pp + 255) / 256, 256, 0, s>>>(Pp, n.in_t, b.dE, b.dZt1, b.dZt1T);
    CK(dw_gemm(b.dZt1T, 32, b.T1T, WID, Pp, gr.dWt1, WID, s));
    colsum_kernel<<<rb, 256, 0, s>>>(Pp, 32, 32, b.dZt1, gr.dbt1);
    GemmArgs g = gemm(b.dZt1, 32, (CB)n.Wt1T, 32, Pp, WID, 32);
    g.mask = b.T1; g.ld_mask = WID;
    g.out_bf16 = b.dZ[cur ^ 1]; g.ld_bf16 = WID; g.out_bf16_t = b.dZT[cur ^ 1]; g.ld_t = Pp;
    CK(launch_gemm(g, s));
    CK(dw_gemm(b.dZT[cur ^ 1], WID, b.T0T, 16, Pp, gr.dWt0, 16, s));
    colsum_kernel<<<rb, 256, 0, s>>>(Pp, WID, WID, b.dZ[cur ^ 1], gr.dbt0);
  }
  if (dx) pe_backward_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, x, b.dE, dx);
  return cudaGetLastError();
}

size_t mlp_workspace_bytes(int P, int train) {
  size_t bytes;
  MlpBufs::carve_all(nullptr, P, train, &bytes);
  return bytes;
}

}  // namespace dgm

// =====================================================================================
// Parameter packing: reference-shaped fp32 parameters (nn.Linear weights [out, in]) -> the padded
// bf16 matrices + transposes of DglNet, and packed fp32 gradients -> reference-shaped gradients.
// Column map of an input-facing matrix: c < 63 -> c ; 63 <= c < 63+in_t -> c + 1 ; beyond (the
// hidden part of the skip layer) -> 96 + (c - 63 - in_t).
// =====================================================================================
namespace dgm {

__device__ __forceinline__ int map_col(int c, int in_t, int mapped) {
  if (!mapped) return c;
  if (c < XE) return c;
  if (c < XE + in_t) return c + 1;
  return K0 + (c - XE - in_t);
}

// src fp32 [rows, kin] -> dst bf16 [rows_pad?, kpad] at row offset r0 (+ transposed [kpad, ldt])
__global__ void pack_w_kernel(const float* __restrict__ src, int rows, int kin, int in_t, int mapped, int r0,
                              __nv_bfloat16* __restrict__ dst, int kpad, __nv_bfloat16* __restrict__ dstT, int ldt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * kin) return;
  const int r = i / kin, c = i % kin;
  const int cc = map_col(c, in_t, mapped);
  const __nv_bfloat16 v = __float2bfloat16_rn(src[i]);
  dst[(size_t)(r0 + r) * kpad + cc] = v;
  if (dstT) dstT[(size_t)cc * ldt + r0 + r] = v;
}
// packed fp32 grad [*, kpad] (row offset r0) -> reference-shaped [rows, kin]
__global__ void unpack_w_kernel(const float* __restrict__ src, int kpad, int r0, int rows, int kin, int in_t,
                                int mapped, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * kin) return;
  const int r = i / kin, c = i % kin;
  dst[i] = src[(size_t)(r0 + r) * kpad + map_col(c, in_t, mapped)];
}
__global__ void copy_f32_kernel(const float* __restrict__ src, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// fixed layout of the packed buffers (elements)
struct PackLayout {
  size_t W[8], WT[8], Wh, WhT, Wt0, Wt1, Wt1T, w_total;  // bf16 elements
  size_t b[8], bh, bt0, bt1, b_total;                     // fp32 elements
  // gradient buffer (fp32 elements): same matrices, no transposes
  size_t gW[8], gb[8], gWh, gbh, gWt0, gbt0, gWt1, gbt1, g_total;
  PackLayout() {
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    for (int l = 0; l < 8; ++l) {
      const size_t K = (l == 0) ? K0 : (l == 5 ? K5 : WID);
      W[l] = take(WID * K);
      WT[l] = take(WID * K);
    }
    Wh = take(16 * WID); WhT = take(16 * WID);
    Wt0 = take(WID * 16); Wt1 = take(32 * WID); Wt1T = take(32 * WID);
    w_total = o;
    o = 0;
    for (int l = 0; l < 8; ++l) b[l] = take(WID);
    bh = take(16); bt0 = take(WID); bt1 = take(32);
    b_total = o;
    o = 0;
    for (int l = 0; l < 8; ++l) {
      const size_t K = (l == 0) ? K0 : (l == 5 ? K5 : WID);
      gW[l] = take(WID * K);
      gb[l] = take(WID);
    }
    gWh = take(16 * WID); gbh = take(16); gWt0 = take(WID * 16); gbt0 = take(WID); gWt1 = take(32 * WID);
    gbt1 = take(32);
    g_total = o;
  }
};
static const PackLayout& layout() {
  static PackLayout L;
  return L;
}

void mlp_pack_sizes(size_t* w_bytes, size_t* b_bytes, size_t* g_bytes) {
  const PackLayout& L = layout();
  if (w_bytes) *w_bytes = L.w_total * 2;
  if (b_bytes) *b_bytes = L.b_total * 4;
  if (g_bytes) *g_bytes = L.g_total * 4;
}

static int in_width(const DglRaw& r) { return XE + r.in_t; }

cudaError_t launch_mlp_pack(const DglRaw& r, void* wbuf, float* bbuf, DglNet* net, cudaStream_t s) {
  const PackLayout& L = layout();
  __nv_bfloat16* w = (__nv_bfloat16*)wbuf;
  cudaMemsetAsync(wbuf, 0, L.w_total * 2, s);
  cudaMemsetAsync(bbuf, 0, L.b_total * 4, s);
  net->has_timenet = r.has_timenet; net->in_t = r.in_t; net->sigmoid_out = r.sigmoid_out;
  auto pack = [&](const float* src, int rows, int kin, int mapped, int r0, size_t dst, int kpad, size_t dstT,
                  int ldt) {
    const int n = rows * kin;
    pack_w_kernel<<<(n + 255) / 256, 256, 0, s>>>(src, rows, kin, r.in_t, mapped, r0, w + dst, kpad,
                                                  dstT != (size_t)-1 ? w + dstT : nullptr, ldt);
  };
  for (int l = 0; l < 8; ++l) {
    const int kin = (l == 0) ? in_width(r) : (l == 5 ? in_width(r) + WID : WID);
    const int kpad = (l == 0) ? K0 : (l == 5 ? K5 : WID);
    pack(r.W[l], WID, kin, l == 0 || l == 5, 0, L.W[l], kpad, L.WT[l], WID);
    copy_f32_kernel<<<1, 256, 0, s>>>(r.b[l], WID, bbuf + L.b[l]);
    net->W[l] = w + L.W[l]; net->WT[l] = w + L.WT[l]; net->b[l] = bbuf + L.b[l];
  }
  int r0 = 0;
  for (int h = 0; h < r.n_heads; ++h) {
    pack(r.Wh[h], r.head_rows[h], WID, 0, r0, L.Wh, WID, L.WhT, 16);
    copy_f32_kernel<<<1, 256, 0, s>>>(r.bh[h], r.head_rows[h], bbuf + L.bh + r0);
    r0 += r.head_rows[h];
  }
  net->n_out = r0;
  net->Wh = w + L.Wh; net->WhT = w + L.WhT; net->bh = bbuf + L.bh;
  if (r.has_timenet) {
    pack(r.Wt0, WID, 13, 0, 0, L.Wt0, 16, (size_t)-1, 0);
    pack(r.Wt1, r.in_t, WID, 0, 0, L.Wt1, WID, L.Wt1T, 32);
    copy_f32_kernel<<<1, 256, 0, s>>>(r.bt0, WID, bbuf + L.bt0);
    copy_f32_kernel<<<1, 256, 0, s>>>(r.bt1, r.in_t, bbuf + L.bt1);
    net->Wt0 = w + L.Wt0; net->Wt1 = w + L.Wt1; net->Wt1T = w + L.Wt1T;
    net->bt0 = bbuf + L.bt0; net->bt1 = bbuf + L.bt1;
  } else {
    net->Wt0 = net->Wt1 = net->Wt1T = nullptr;
    net->bt0 = net->bt1 = nullptr;
  }
  return cudaGetLastError();
}

void mlp_grad_pointers(float* gbuf, DglGrads* g) {
  const PackLayout& L = layout();
  for (int l = 0; l < 8; ++l) {
    g->dW[l] = gbuf + L.gW[l];
    g->db[l] = gbuf + L.gb[l];
  }
  g->dWh = gbuf + L.gWh; g->dbh = gbuf + L.gbh;
  g->dWt0 = gbuf + L.gWt0; g->dbt0 = gbuf + L.gbt0; g->dWt1 = gbuf + L.gWt1; g->dbt1 = gbuf + L.gbt1;
}

// packed gradients -> tensors shaped like the reference parameters (DglRawGrads mirrors DglRaw)
cudaError_t launch_mlp_unpack_grads(const DglRaw& r, const float* gbuf, const DglRawGrads& o, cudaStream_t s) {
  const PackLayout& L = layout();
  auto unpack = [&](size_t src, int kpad, int r0, int rows, int kin, int mapped, float* dst) {
    const int n = rows * kin;
    unpack_w_kernel<<<(n + 255) / 256, 256, 0, s>>>(gbuf + src, kpad, r0, rows, kin, r.in_t, mapped, dst);
  };
  for (int l = 0; l < 8; ++l) {
    const int kin = (l == 0) ? in_width(r) : (l == 5 ? in_width(r) + WID : WID);
    const int kpad = (l == 0) ? K0 : (l == 5 ? K5 : WID);
    unpack(L.gW[l], kpad, 0, WID, kin, l == 0 || l == 5, o.W[l]);
    copy_f32_kernel<<<1, 256, 0, s>>>(gbuf + L.gb[l], WID, o.b[l]);
  }
  int r0 = 0;
  for (int h = 0; h < r.n_heads; ++h) {
    unpack(L.gWh, WID, r0, r.head_rows[h], WID, 0, o.Wh[h]);
    copy_f32_kernel<<<1, 256, 0, s>>>(gbuf + L.gbh + r0, r.head_rows[h], o.bh[h]);
    r0 += r.head_rows[h];
  }
  if (r.has_timenet) {
    unpack(L.gWt0, 16, 0, WID, 13, 0, o.Wt0);
    unpack(L.gWt1, WID, 0, r.in_t, WID, 0, o.Wt1);
    copy_f32_kernel<<<1, 256, 0, s>>>(gbuf + L.gbt0, WID, o.bt0);
    copy_f32_kernel<<<1, 256, 0, s>>>(gbuf + L.gbt1, r.in_t, o.bt1);
  }
  return cudaGetLastError();
}

}  // namespace dgm
