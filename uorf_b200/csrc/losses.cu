// Image-side glue of the training step as two pairs of launches (reference models/uorf_nogan_model.py:168-183):
//   rendered = rgb_map as N x 3 x H x W;  x_recon = rendered * 2 - 1;  loss_recon = mse(x_recon, x);
//   rendered_norm = (rendered - mean) / std,  x_norm = ((x + 1) / 2 - mean) / std      (vgg_normalizer)
//   loss_perc = mse(features[:N], features[N:])
// torch runs these as ~25 elementwise / reduce / fill launches per step (forward and backward); the arithmetic is kept operation
// for operation (same roundings), the two means are block partial sums added in a fixed order by the last block to finish.
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kLossBlocks = 64;
constexpr int kLossThreads = 256;

// block sum (fixed tree) -> partials[blockIdx.x]; the last block adds the partials in index order and writes *out = sum * scale.
// `ticket` is a zero-initialised counter that the last block resets: graph replays need no memset.
__device__ __forceinline__ void finish_mean(float v, float* partials, unsigned int* ticket, float scale, float* out) {
  __shared__ float s_red[kLossThreads / 32];
  __shared__ bool s_last;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) b += s_red[w];
    partials[blockIdx.x] = b;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float tot = 0.f;
    for (unsigned int b = 0; b < gridDim.x; ++b) tot += reinterpret_cast<volatile float*>(partials)[b];
    *out = tot * scale;
    *ticket = 0u;
  }
}

// rgb [N*HW][3] (compositing output) and x [N][3][HW] -> x_recon, rendered, rendered_norm, x_norm (all [N][3][HW]) and loss_recon
__global__ void __launch_bounds__(kLossThreads) image_loss_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ x,
                                                                      const float* __restrict__ mean, const float* __restrict__ stdv, int n, int hw,
                                                                      float* __restrict__ x_recon, float* __restrict__ rendered,
                                                                      float* __restrict__ rendered_norm, float* __restrict__ x_norm,
                                                                      float* partials, unsigned int* ticket, float* __restrict__ loss) {
  const int total = n * 3 * hw;
  float acc = 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int pix = e % hw, c = (e / hw) % 3, img = e / (3 * hw);
    const float r = __ldg(rgb + ((size_t)img * hw + pix) * 3 + c);
    const float xv = __ldg(x + e);
    const float m = __ldg(mean + c), sd = __ldg(stdv + c);
    const float xr = __fsub_rn(__fmul_rn(r, 2.f), 1.f);                       // rendered * 2 - 1
    const float d = __fsub_rn(xr, xv);
    acc = __fadd_rn(acc, __fmul_rn(d, d));
    x_recon[e] = xr;
    rendered[e] = r;
    rendered_norm[e] = __fdiv_rn(__fsub_rn(r, m), sd);                        // (rendered - mean) / std
    x_norm[e] = __fdiv_rn(__fsub_rn(__fdiv_rn(__fadd_rn(xv, 1.f), 2.f), m), sd);   // ((x + 1) / 2 - mean) / std
  }
  finish_mean(acc, partials, ticket, 1.f / (float)total, loss);
}

// grad_rgb [N*HW][3] = 2 * (g_x_recon + g_loss * 2 (x_recon - x) / numel) + g_rendered_norm / std     (any of the three may be null)
__global__ void __launch_bounds__(kLossThreads) image_loss_bwd_kernel(const float* __restrict__ x_recon, const float* __restrict__ x,
                                                                      const float* __restrict__ stdv, const float* __restrict__ g_x_recon,
                                                                      const float* __restrict__ g_rendered_norm, const float* __restrict__ g_loss,
                                                                      int n, int hw, float* __restrict__ grad_rgb) {
  const int total = n * 3 * hw;
  const float gl = g_loss ? __ldg(g_loss) * (2.f / (float)total) : 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int pix = e % hw, c = (e / hw) % 3, img = e / (3 * hw);
    float g = g_loss ? gl * (x_recon[e] - __ldg(x + e)) : 0.f;
    if (g_x_recon) g += __ldg(g_x_recon + e);
    g *= 2.f;
    if (g_rendered_norm) g += __ldg(g_rendered_norm + e) / __ldg(stdv + c);
    grad_rgb[((size_t)img * hw + pix) * 3 + c] = g;
  }
}

// features [2][m] (rendered half, ground-truth half, same layout) -> loss = mean((a - b)^2)
__global__ void __launch_bounds__(kLossThreads) feat_mse_fwd_kernel(const float* __restrict__ f, long long m, float* partials, unsigned int* ticket,
                                                                    float* __restrict__ loss) {
  float acc = 0.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (long long)gridDim.x * blockDim.x) {
    const float d = __fsub_rn(__ldg(f + e), __ldg(f + m + e));
    acc = __fadd_rn(acc, __fmul_rn(d, d));
  }
  finish_mean(acc, partials, ticket, 1.f / (float)m, loss);
}

// grad [2][m]: first half g_loss * 2 (a - b) / m, second half zero (the ground-truth features are constants)
__global__ void __launch_bounds__(kLossThreads) feat_mse_bwd_kernel(const float* __restrict__ f, const float* __restrict__ g_loss, long long m,
                                                                    float* __restrict__ grad) {
  const float gl = __ldg(g_loss) * (2.f / (float)m);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (long long)gridDim.x * blockDim.x) {
    grad[e] = gl * (__ldg(f + e) - __ldg(f + m + e));
    grad[m + e] = 0.f;
  }
}

int launch_image_loss_fwd(const float* rgb, const float* x, const float* mean, const float* stdv, int n, int hw, float* x_recon, float* rendered,
                          float* rendered_norm, float* x_norm, float* scratch, float* loss, cudaStream_t stream) {
  if ((long long)n * 3 * hw >= (1LL << 31)) return UORF_ERR_UNSUPPORTED;
  image_loss_fwd_kernel<<<kLossBlocks, kLossThreads, 0, stream>>>(rgb, x, mean, stdv, n, hw, x_recon, rendered, rendered_norm, x_norm, scratch,
                                                                   reinterpret_cast<unsigned int*>(scratch + kLossBlocks), loss);
  return finish_launch(1);
}

int launch_image_loss_bwd(const float* x_recon, const float* x, const float* stdv, const float* g_x_recon, const float* g_rendered_norm,
                          const float* g_loss, int n, int hw, float* grad_rgb, cudaStream_t stream) {
  if ((long long)n * 3 * hw >= (1LL << 31)) return UORF_ERR_UNSUPPORTED;
  const long long total = (long long)n * 3 * hw;
  long long blocks = (total + kLossThreads - 1) / kLossThreads;
  if (blocks > 148 * 4) blocks = 148 * 4;
  image_loss_bwd_kernel<<<(unsigned)blocks, kLossThreads, 0, stream>>>(x_recon, x, stdv, g_x_recon, g_rendered_norm, g_loss, n, hw, grad_rgb);
  return finish_launch(1);
}

int launch_feat_mse_fwd(const float* f, long long m, float* scratch, float* loss, cudaStream_t stream) {
  feat_mse_fwd_kernel<<<kLossBlocks, kLossThreads, 0, stream>>>(f, m, scratch, reinterpret_cast<unsigned int*>(scratch + kLossBlocks), loss);
  return finish_launch(1);
}

int launch_feat_mse_bwd(const float* f, const float* g_loss, long long m, float* grad, cudaStream_t stream) {
  long long blocks = (m + kLossThreads - 1) / kLossThreads;
  if (blocks > 148 * 4) blocks = 148 * 4;
  feat_mse_bwd_kernel<<<(unsigned)blocks, kLossThreads, 0, stream>>>(f, g_loss, m, grad);
  return finish_launch(1);
}

}  // namespace uorf
