// Small host-side-of-the-step helpers: the Adam update of a list of small fp32 tensors in one launch, and the pose inverse.
// The training drivers step torch.optim.Adam over ~80 tensors of 64 .. 73,728 elements (reference models/uorf_nogan_model.py:59-61,
// 229-245).  torch's fused multi-tensor kernel gives every 65,536-element chunk to ONE block: 38 blocks, each a chain of 32
// dependent load / store rounds (~70 us for 430 K parameters).  Here a block takes 1,024 elements; the chunk table (pointers of the
// parameter, its gradient, both moments and the step counter) is built once by the host side and lives on the device.
// Same update rule as torch (torch/optim/adam.py, ATen fused_adam_utils.cuh: no amsgrad, not maximizing).
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kAdamChunk = UORF_ADAM_CHUNK;   // elements per block (the host side cuts the tensors)
constexpr int kAdamThreads = 256;

__global__ void __launch_bounds__(kAdamThreads) adam_step_kernel(const uorf_adam_chunk* __restrict__ chunks, const float* __restrict__ lr_dev,
                                                                 float lr_host, float beta1, float beta2, float eps, float weight_decay) {
  const uorf_adam_chunk c = chunks[blockIdx.x];
  const float lr = lr_dev ? __ldg(lr_dev) : lr_host;
  const float step = __ldg(c.step);                    // already incremented for this update
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  const float step_size = lr / bc1;
  const float bc2_sqrt = sqrtf(bc2);
  for (int i = threadIdx.x; i < c.n && i < kAdamChunk; i += kAdamThreads) {
    float g = c.grad[i];
    const float w = c.param[i];
    if (weight_decay != 0.f) g = fmaf(weight_decay, w, g);
    float m = c.exp_avg[i], v = c.exp_avg_sq[i];
    m = m + (1.f - beta1) * (g - m);
    v = beta2 * v + (1.f - beta2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    c.exp_avg[i] = m;
    c.exp_avg_sq[i] = v;
    c.param[i] = w - step_size * m / denom;
  }
}

int launch_adam_step(const uorf_adam_chunk* chunks, int n_chunks, const float* lr_dev, float lr_host, float beta1, float beta2, float eps,
                     float weight_decay, cudaStream_t stream) {
  if (n_chunks <= 0) return UORF_OK;
  adam_step_kernel<<<n_chunks, kAdamThreads, 0, stream>>>(chunks, lr_dev, lr_host, beta1, beta2, eps, weight_decay);
  return finish_launch(1);
}

// Inverse of small pose matrices on the device: the drivers invert one 3x3 azimuth rotation or 4x4 camera pose per scene
// (reference models/uorf_nogan_model.py:117-126, uorf_eval_model.py:98-107: `.inverse()`).  The library route is an LU
// factorisation + two triangular solves + pivot bookkeeping = 13 launches for 16 numbers; here one thread does Gauss-Jordan with
// partial pivoting in registers (same pivoting rule as LU with row exchanges; results agree to fp32 rounding).
template <int N>
__global__ void invert_small_kernel(const float* __restrict__ src, float* __restrict__ dst, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  float a[N][2 * N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) { a[i][j] = src[(b * N + i) * N + j]; a[i][N + j] = i == j ? 1.f : 0.f; }
#pragma unroll
  for (int col = 0; col < N; ++col) {
    int piv = col;
    float best = fabsf(a[col][col]);
#pragma unroll
    for (int r = col + 1; r < N; ++r)
      if (fabsf(a[r][col]) > best) { best = fabsf(a[r][col]); piv = r; }
#pragma unroll
    for (int r = col + 1; r < N; ++r)           // row exchange with compile-time indices (registers)
      if (piv == r) {
#pragma unroll
        for (int j = 0; j < 2 * N; ++j) { const float t = a[col][j]; a[col][j] = a[r][j]; a[r][j] = t; }
      }
    const float inv = 1.f / a[col][col];
#pragma unroll
    for (int j = 0; j < 2 * N; ++j) a[col][j] *= inv;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (r == col) continue;
      const float f = a[r][col];
#pragma unroll
      for (int j = 0; j < 2 * N; ++j) a[r][j] = fmaf(-f, a[col][j], a[r][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) dst[(b * N + i) * N + j] = a[i][N + j];
}

int launch_invert_small(const float* src, float* dst, int batch, int dim, cudaStream_t stream) {
  if (batch <= 0) return UORF_OK;
  const unsigned blocks = (unsigned)((batch + 63) / 64);
  if (dim == 3) invert_small_kernel<3><<<blocks, 64, 0, stream>>>(src, dst, batch);
  else if (dim == 4) invert_small_kernel<4><<<blocks, 64, 0, stream>>>(src, dst, batch);
  else return UORF_ERR_UNSUPPORTED;
  return finish_launch(1);
}

}  // namespace uorf
