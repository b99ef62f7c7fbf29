// Encoder convolutions for the inference path: 3x3 conv (pad 1, stride 1 or 2) + bias + ReLU in fp32 FFMA, with the
// reference's data movement folded into the input read:
//   * the input is the channel concatenation of two sources (torch.cat of model.py:44-49,58-60 is never materialised);
//   * source A may be read through a x2 bilinear upsample (nn.Upsample(scale_factor=2, align_corners=False), model.py:27-32),
//     so enc_up_3 / enc_up_2 store their low-resolution conv output and the consumer interpolates on the fly.
// Replaces the six (seven with `bottom`) cuDNN launches + 2 upsample + 3 cat launches of Encoder.forward
// (reference models/model.py:36-61) when no gradient is needed; training keeps torch autograd / cuDNN.
// The layers are tiny (0.9 GFLOP for the whole U-Net at 64x64): latency-bound library kernels take ~60 us each, a direct
// shared-memory-tiled kernel ~10 us.  CTA = 8x8 output pixels x 32 output channels, 128 threads, thread = 4 pixels x 4 channels.
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kEncTile = 8;      // output pixels per CTA side
constexpr int kEncOcb = 32;      // output channels per CTA
constexpr int kEncIcb = 8;       // input channels per shared-memory chunk
constexpr int kEncThreads = 128;

struct EncConvParams {
  const float* src_a; int ca; int up_a;     // up_a: src_a is [ca][H/2][W/2], read through the x2 bilinear upsample
  const float* src_b; int cb;               // [cb][H][W] or null
  const float* weight_t;                    // [ca+cb][9][c_out]  (transposed once on the host side)
  const float* bias;                        // [c_out]
  int c_out, h_in, w_in, h_out, w_out;
  float* out;                               // [c_out][h_out][w_out]
};

// logical input value (zero padding outside the image); PyTorch's upsample_bilinear2d arithmetic for the upsampled source
__device__ __forceinline__ float enc_fetch(const EncConvParams& p, int c, int y, int x) {
  if (y < 0 || y >= p.h_in || x < 0 || x >= p.w_in) return 0.f;
  if (c < p.ca) {
    if (!p.up_a) return __ldg(p.src_a + ((long long)c * p.h_in + y) * p.w_in + x);
    const int hs = p.h_in >> 1, ws = p.w_in >> 1;
    const float sy = fmaxf(0.5f * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (x + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float* s = p.src_a + (long long)c * hs * ws;
    const float a00 = __ldg(s + y0 * ws + x0), a01 = __ldg(s + y0 * ws + x1), a10 = __ldg(s + y1 * ws + x0), a11 = __ldg(s + y1 * ws + x1);
    return (1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11);
  }
  return __ldg(p.src_b + ((long long)(c - p.ca) * p.h_in + y) * p.w_in + x);
}

template <int STRIDE>
__global__ void __launch_bounds__(kEncThreads) encoder_conv3x3_kernel(const EncConvParams p) {
  constexpr int PH = (kEncTile - 1) * STRIDE + 3;       // input patch side: 10 (stride 1) / 17 (stride 2)
  constexpr int PW = PH + 1;                            // padded row
  constexpr int NCOL = 3 * STRIDE + 3;                  // input columns one thread touches per row: 6 / 9
  __shared__ float s_in[kEncIcb][PH][PW];
  __shared__ __align__(16) float s_w[kEncIcb][9][kEncOcb];
  const int tiles_x = (p.w_out + kEncTile - 1) / kEncTile;
  const int ty0 = (blockIdx.x / tiles_x) * kEncTile, tx0 = (blockIdx.x % tiles_x) * kEncTile;
  const int oc_base = blockIdx.y * kEncOcb;
  const int tid = threadIdx.x;
  const int og = tid & 7;                  // output-channel group: channels oc_base + 4*og .. +3
  const int pg = tid >> 3;                 // pixel group 0..15: row pg >> 1, columns 4*(pg & 1) .. +3
  const int py = pg >> 1, px0 = (pg & 1) * 4;
  const int cin = p.ca + p.cb;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[j][o] = 0.f;

  for (int ic0 = 0; ic0 < cin; ic0 += kEncIcb) {
    __syncthreads();
    for (int e = tid; e < kEncIcb * PH * PH; e += kEncThreads) {
      const int i = e / (PH * PH), r = (e / PH) % PH, c = e % PH;
      const int ic = ic0 + i;
      s_in[i][r][c] = ic < cin ? enc_fetch(p, ic, ty0 * STRIDE - 1 + r, tx0 * STRIDE - 1 + c) : 0.f;
    }
    for (int e = tid; e < kEncIcb * 9 * kEncOcb; e += kEncThreads) {
      const int i = e / (9 * kEncOcb), k = (e / kEncOcb) % 9, o = e % kEncOcb;
      const int ic = ic0 + i, oc = oc_base + o;
      s_w[i][k][o] = (ic < cin && oc < p.c_out) ? __ldg(p.weight_t + ((long long)ic * 9 + k) * p.c_out + oc) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < kEncIcb; ++i) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float a[NCOL];
        const float* row = &s_in[i][py * STRIDE + ky][px0 * STRIDE];
#pragma unroll
        for (int c = 0; c < NCOL; ++c) a[c] = row[c];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 w = *reinterpret_cast<const float4*>(&s_w[i][ky * 3 + kx][4 * og]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float v = a[j * STRIDE + kx];
            acc[j][0] = fmaf(v, w.x, acc[j][0]);
            acc[j][1] = fmaf(v, w.y, acc[j][1]);
            acc[j][2] = fmaf(v, w.z, acc[j][2]);
            acc[j][3] = fmaf(v, w.w, acc[j][3]);
          }
        }
      }
    }
  }
  const int y = ty0 + py;
  if (y >= p.h_out) return;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int oc = oc_base + 4 * og + o;
    if (oc >= p.c_out) continue;
    const float b = __ldg(p.bias + oc);
    float* dst = p.out + ((long long)oc * p.h_out + y) * p.w_out + tx0 + px0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (tx0 + px0 + j < p.w_out) dst[j] = fmaxf(acc[j][o] + b, 0.f);
  }
}

int launch_encoder_conv3x3(const float* src_a, int ca, int up_a, const float* src_b, int cb, const float* weight_t, const float* bias,
                           int c_out, int h_in, int w_in, int stride, float* out, cudaStream_t stream) {
  EncConvParams p;
  p.src_a = src_a; p.ca = ca; p.up_a = up_a;
  p.src_b = src_b; p.cb = cb;
  p.weight_t = weight_t; p.bias = bias;
  p.c_out = c_out; p.h_in = h_in; p.w_in = w_in;
  p.h_out = (h_in - 1) / stride + 1;
  p.w_out = (w_in - 1) / stride + 1;
  p.out = out;
  dim3 grid(((p.h_out + kEncTile - 1) / kEncTile) * ((p.w_out + kEncTile - 1) / kEncTile), (c_out + kEncOcb - 1) / kEncOcb);
  if (stride == 1) encoder_conv3x3_kernel<1><<<grid, kEncThreads, 0, stream>>>(p);
  else encoder_conv3x3_kernel<2><<<grid, kEncThreads, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

}  // namespace uorf
