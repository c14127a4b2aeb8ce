// loss.cu -- image loss of the training step: (1 - lambda) * L1 + lambda * (1 - SSIM), forward + backward.
//
// Reference: dgmesh/utils/loss_utils.py:18-19 (l1_loss), :33-76 (gaussian window 11 / sigma 1.5, ssim,
// _ssim) as composed in dgmesh/train.py:308-311 (and :270-273 for the mesh image).  The reference runs
// five 11x11 depthwise conv2d (zero padding 5) + ~15 elementwise kernels forward and the same again
// backward, each a full pass over [3,H,W] maps.  Here:
//   ssim_stats_kernel   one pass: per 16x16 tile the five blurred maps (E[x], E[y], E[x^2], E[y^2],
//                       E[xy]) are formed separably in shared memory, the SSIM value and |x - y| are
//                       block-reduced into two double accumulators, and the three partial derivatives
//                       dS/dE[x], dS/dE[x^2], dS/dE[xy] are stored for the backward pass
//   loss_finalize_kernel  loss, l1, ssim from the two sums (no host read)
//   ssim_grad_kernel    one pass: blur of the three partial-derivative maps (the window is symmetric,
//                       so the adjoint of the blur is the blur) combined with x, y and sign(x - y):
//                       dL/dx = (1-l)/n sign(x-y) - l/n (blur(A) + 2 x blur(B) + y blur(C))
// HBM-bound: algorithmic bytes 2*12 n (read x, y) + 3*4 n (partials) forward, 5*4 n + 4 n backward,
// n = 3 H W.
#include "common.cuh"
#include "loss_kernels.h"

namespace dgm {

#define LT 16            // tile edge
#define LR 5             // window radius
#define LP (LT + 2 * LR)  // patch edge (26)

__constant__ float c_win[11];

static unsigned long long g_win_uploaded = 0;  // per device
static cudaError_t upload_window() {
  if (!once_per_device(g_win_uploaded)) return cudaSuccess;
  // loss_utils.gaussian(11, 1.5): exp(-(x - 5)^2 / (2 * 1.5^2)) normalised, float32 like torch.Tensor
  float w[11], sum = 0.f;
  for (int i = 0; i < 11; ++i) {
    w[i] = (float)exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
    sum += w[i];
  }
  for (int i = 0; i < 11; ++i) w[i] /= sum;
  cudaError_t e = cudaMemcpyToSymbol(c_win, w, sizeof(w));
  if (e != cudaSuccess) g_win_uploaded = 0;  // retry on the next call
  return e;
}

// separable blur of NM maps held as a zero-padded LP x LP patch per map: horizontal pass into s_h
// (LP rows x LT columns), then this thread's vertical sum
template <int NM>
__device__ __forceinline__ void blur_patch(const float (&s_p)[NM][LP][LP + 1], float (&s_h)[NM][LP][LT], int tx, int ty,
                                           float (&out)[NM]) {
  const int tid = ty * LT + tx;
  for (int i = tid; i < LP * LT; i += LT * LT) {
    const int r = i / LT, c = i % LT;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) a = fmaf(c_win[k], s_p[m][r][c + k], a);
      s_h[m][r][c] = a;
    }
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) a = fmaf(c_win[k], s_h[m][ty + k][tx], a);
    out[m] = a;
  }
}

__global__ void __launch_bounds__(LT * LT) ssim_stats_kernel(int H, int W, const float* __restrict__ img,
                                                             const float* __restrict__ gt,
                                                             float* __restrict__ partials, double* __restrict__ sums) {
  __shared__ float s_p[5][LP][LP + 1];
  __shared__ float s_h[5][LP][LT];
  __shared__ double s_red[2][LT * LT / 32];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * LT + tx;
  const int ch = blockIdx.z;
  const int x0 = blockIdx.x * LT - LR, y0 = blockIdx.y * LT - LR;
  const size_t plane = (size_t)H * W;
  const float* xi = img + ch * plane;
  const float* yi = gt + ch * plane;
  for (int i = tid; i < LP * LP; i += LT * LT) {
    const int r = i / LP, c = i % LP;
    const int gx = x0 + c, gy = y0 + r;
    float a = 0.f, b = 0.f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
      a = xi[(size_t)gy * W + gx];
      b = yi[(size_t)gy * W + gx];
    }
    s_p[0][r][c] = a;
    s_p[1][r][c] = b;
    s_p[2][r][c] = a * a;
    s_p[3][r][c] = b * b;
    s_p[4][r][c] = a * b;
  }
  __syncthreads();
  float e[5];
  blur_patch<5>(s_p, s_h, tx, ty, e);
  const int px = blockIdx.x * LT + tx, py = blockIdx.y * LT + ty;
  double my_s = 0.0, my_l1 = 0.0;
  if (px < W && py < H) {
    const float m1 = e[0], m2 = e[1];
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float s1 = e[2] - m1 * m1, s2 = e[3] - m2 * m2, s12 = e[4] - m1 * m2;
    const float N1 = 2.f * m1 * m2 + C1, N2 = 2.f * s12 + C2;
    const float D1 = m1 * m1 + m2 * m2 + C1, D2 = s1 + s2 + C2;
    const float inv = 1.0f / (D1 * D2);
    const float S = N1 * N2 * inv;
    const size_t o = ch * plane + (size_t)py * W + px;
    // dS/dE[x], dS/dE[x^2], dS/dE[xy] (everything else fixed)
    partials[o] = (2.f * m2 * N2 - 2.f * m2 * N1) * inv - S * (2.f * m1 / D1 - 2.f * m1 / D2);
    partials[3 * plane + o] = -S / D2;
    partials[6 * plane + o] = 2.f * N1 * inv;
    my_s = (double)S;
    my_l1 = (double)fabsf(s_p[0][ty + LR][tx + LR] - s_p[1][ty + LR][tx + LR]);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    my_s += __shfl_xor_sync(0xffffffffu, my_s, o);
    my_l1 += __shfl_xor_sync(0xffffffffu, my_l1, o);
  }
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = my_s;
    s_red[1][tid >> 5] = my_l1;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < LT * LT / 32; ++w) a += s_red[0][w], b += s_red[1][w];
    atomicAdd(&sums[0], a);
    atomicAdd(&sums[1], b);
  }
}

// out = {loss, l1, ssim};  mode 0: (1-l) l1 + l (1 - ssim);  mode 1: the SSIM value itself
__global__ void loss_finalize_kernel(double n, float lam, int mode, const double* __restrict__ sums,
                                     float* __restrict__ out) {
  const double ssim = sums[0] / n, l1 = sums[1] / n;
  out[0] = mode ? (float)ssim : (float)((1.0 - (double)lam) * l1 + (double)lam * (1.0 - ssim));
  out[1] = (float)l1;
  out[2] = (float)ssim;
}

__global__ void __launch_bounds__(LT * LT) ssim_grad_kernel(int H, int W, const float* __restrict__ img,
                                                            const float* __restrict__ gt,
                                                            const float* __restrict__ partials, float lam, int mode,
                                                            const float* __restrict__ upstream,
                                                            float* __restrict__ dL_dimg) {
  __shared__ float s_p[3][LP][LP + 1];
  __shared__ float s_h[3][LP][LT];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * LT + tx;
  const int ch = blockIdx.z;
  const int x0 = blockIdx.x * LT - LR, y0 = blockIdx.y * LT - LR;
  const size_t plane = (size_t)H * W;
  for (int i = tid; i < LP * LP; i += LT * LT) {
    const int r = i / LP, c = i % LP;
    const int gx = x0 + c, gy = y0 + r;
    const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
    const size_t o = ch * plane + (size_t)(in ? gy : 0) * W + (in ? gx : 0);
    s_p[0][r][c] = in ? partials[o] : 0.f;
    s_p[1][r][c] = in ? partials[3 * plane + o] : 0.f;
    s_p[2][r][c] = in ? partials[6 * plane + o] : 0.f;
  }
  __syncthreads();
  float b[3];
  blur_patch<3>(s_p, s_h, tx, ty, b);
  const int px = blockIdx.x * LT + tx, py = blockIdx.y * LT + ty;
  if (px >= W || py >= H) return;
  const size_t o = ch * plane + (size_t)py * W + px;
  const float x = img[o], y = gt[o];
  const float inv_n = 1.0f / (3.0f * (float)plane);
  const float dssim = (b[0] + 2.f * x * b[1] + y * b[2]) * inv_n;
  float g;
  if (mode) {
    g = dssim;
  } else {
    const float d = x - y;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);  // torch: d|x|/dx = sign(x), 0 at 0
    g = (1.f - lam) * sgn * inv_n - lam * dssim;
  }
  dL_dimg[o] = g * (upstream ? upstream[0] : 1.0f);
}

size_t loss_workspace_bytes(int H, int W) { return (size_t)9 * H * W * sizeof(float) + 128 + 64; }

struct LossWS {
  float* partials;
  double* sums;
  static LossWS from(char* base, int H, int W) {
    char* p = base;
    LossWS w;
    w.partials = carve<float>(p, (size_t)9 * H * W);
    w.sums = carve<double>(p, 4);
    return w;
  }
};

cudaError_t launch_loss_forward(int H, int W, const float* img, const float* gt, float lam, int mode, float* out3,
                                void* ws, cudaStream_t s) {
  cudaError_t e = upload_window();
  if (e != cudaSuccess) return e;
  LossWS w = LossWS::from((char*)ws, H, W);
  cudaMemsetAsync(w.sums, 0, 4 * sizeof(double), s);
  dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, 3), block(LT, LT);
  ssim_stats_kernel<<<grid, block, 0, s>>>(H, W, img, gt, w.partials, w.sums);
  loss_finalize_kernel<<<1, 1, 0, s>>>(3.0 * H * W, lam, mode, w.sums, out3);
  return cudaGetLastError();
}

cudaError_t launch_loss_backward(int H, int W, const float* img, const float* gt, float lam, int mode,
                                 const float* upstream, float* dL_dimg, void* ws, cudaStream_t s) {
  cudaError_t e = upload_window();
  if (e != cudaSuccess) return e;
  LossWS w = LossWS::from((char*)ws, H, W);
  dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, 3), block(LT, LT);
  ssim_grad_kernel<<<grid, block, 0, s>>>(H, W, img, gt, w.partials, lam, mode, upstream, dL_dimg);
  return cudaGetLastError();
}

}  // namespace dgm

// =====================================================================================
// Laplacian mesh regulariser (umbrella operator), dgmesh/nvdiffrast_utils/regularizer.py:40-59 as
// train.py:277-283 uses it:  L = mean_{v,c} ( T[v,c] / max(n_v, 1) )^2 with
//   T[v] = sum over faces containing v of (the two other vertices - 2 v),  n_v = 2 * (faces containing v).
// The reference is 3 gathers + 6 scatter_adds + ~8 elementwise kernels forward and autograd's mirror
// image backward; here: one face pass + one vertex pass each way.
// =====================================================================================
namespace dgm {

__global__ void __launch_bounds__(256) lap_face_fwd_kernel(int F, const float* __restrict__ v,
                                                           const int* __restrict__ tri, float* __restrict__ term,
                                                           float* __restrict__ norm) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int i[3] = {tri[3 * f], tri[3 * f + 1], tri[3 * f + 2]};
  float p[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) p[k][c] = v[3 * (size_t)i[k] + c];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int a = (k + 1) % 3, b = (k + 2) % 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&term[3 * (size_t)i[k] + c], (p[a][c] - p[k][c]) + (p[b][c] - p[k][c]));
    atomicAdd(&norm[i[k]], 2.0f);
  }
}

// per vertex: t = term / max(norm, 1); accumulates sum t^2 (double) and leaves g = 2 t / max(norm, 1) in term
__global__ void __launch_bounds__(256) lap_vertex_kernel(int V, float* __restrict__ term, const float* __restrict__ norm,
                                                         double* __restrict__ sum) {
  __shared__ double s_red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < V) {
    const float n = fmaxf(norm[i], 1.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = term[3 * (size_t)i + c] / n;
      acc += (double)t * (double)t;
      term[3 * (size_t)i + c] = 2.0f * t / n;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += s_red[k];
    atomicAdd(sum, t);
  }
}

__global__ void lap_finalize_kernel(int V, const double* __restrict__ sum, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(sum[0] / (3.0 * (double)V));
}

// dL/dv from g[v] = dL/dT[v] * (3 V) (left in `term` by the forward pass), scaled by dL/dloss / (3 V)
__global__ void __launch_bounds__(256) lap_face_bwd_kernel(int F, int V, const int* __restrict__ tri,
                                                           const float* __restrict__ g, const float* __restrict__ gl,
                                                           float* __restrict__ gv) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float s = (gl ? gl[0] : 1.0f) / (3.0f * (float)V);
  const int i[3] = {tri[3 * f], tri[3 * f + 1], tri[3 * f + 2]};
  float q[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) q[k][c] = g[3 * (size_t)i[k] + c];
  // T[i_k] += p_a + p_b - 2 p_k  ->  gv[i_k] += q_a + q_b - 2 q_k
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int a = (k + 1) % 3, b = (k + 2) % 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&gv[3 * (size_t)i[k] + c], s * (q[a][c] + q[b][c] - 2.0f * q[k][c]));
  }
}

cudaError_t launch_laplacian_forward(int V, int F, const float* verts, const int* tri, float* out, void* ws,
                                     cudaStream_t s) {
  // ws: term[3V] | norm[V] | sum (double)
  float* term = (float*)ws;
  float* norm = term + 3 * (size_t)V;
  double* sum = (double*)(((uintptr_t)(norm + V) + 15) & ~(uintptr_t)15);
  cudaMemsetAsync(ws, 0, (size_t)((char*)(sum + 1) - (char*)ws), s);
  if (F > 0) lap_face_fwd_kernel<<<(F + 255) / 256, 256, 0, s>>>(F, verts, tri, term, norm);
  if (V > 0) lap_vertex_kernel<<<(V + 255) / 256, 256, 0, s>>>(V, term, norm, sum);
  lap_finalize_kernel<<<1, 32, 0, s>>>(V > 0 ? V : 1, sum, out);
  return cudaGetLastError();
}

cudaError_t launch_laplacian_backward(int V, int F, const int* tri, const float* dL_dloss, float* dverts, void* ws,
                                      cudaStream_t s) {
  cudaMemsetAsync(dverts, 0, sizeof(float) * 3 * (size_t)V, s);
  if (F > 0) lap_face_bwd_kernel<<<(F + 255) / 256, 256, 0, s>>>(F, V, tri, (const float*)ws, dL_dloss, dverts);
  return cudaGetLastError();
}

size_t laplacian_ws_bytes(int V) { return sizeof(float) * 4 * (size_t)V + 64; }

}  // namespace dgm
