// Encoder convolutions for the inference path: 3x3 conv (pad 1, stride 1 or 2) + bias + ReLU in fp32 FFMA, with the
// reference's data movement folded into the input read:
//   * the input is the channel concatenation of two sources (torch.cat of model.py:44-49,58-60 is never materialised);
//   * source A may be read through a x2 bilinear upsample (nn.Upsample(scale_factor=2, align_corners=False), model.py:27-32),
//     so enc_up_3 / enc_up_2 store their low-resolution conv output and the consumer interpolates on the fly.
// Replaces the six (seven with `bottom`) cuDNN launches + 2 upsample + 3 cat launches of Encoder.forward
// (reference models/model.py:36-61) when no gradient is needed; training keeps torch autograd / cuDNN.
// The layers are tiny (0.9 GFLOP for the whole U-Net at 64x64): latency-bound library kernels take ~60 us each, a direct
// shared-memory-tiled kernel ~10 us.  CTA = 8x8 output pixels x 32 output channels; 4 warp groups of 128 threads split the input channels of every chunk
// (one warp per scheduler left every FFMA / LDS latency exposed: 30 % issue utilisation), thread = 4 pixels x 4 channels.
#include "common.cuh"
#include <cooperative_groups.h>
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kEncTile = 8;      // output pixels per CTA side
constexpr int kEncOcb = 32;      // output channels per CTA
constexpr int kEncIcb = 16;      // input channels per shared-memory chunk
constexpr int kEncGroups = 4;    // warp groups per CTA; group g accumulates input channels 4g..4g+3 of every chunk
constexpr int kEncGroupThreads = 128;
constexpr int kEncThreads = kEncGroups * kEncGroupThreads;
constexpr int kEncIcPerGroup = kEncIcb / kEncGroups;

struct EncConvParams {
  const float* src_a; int ca; int up_a;     // up_a: src_a is [ca][H/2][W/2], read through the x2 bilinear upsample
  const float* src_b; int cb;               // [cb][H][W] or null
  const float* weight_t;                    // [ca+cb][9][c_out]  (transposed once on the host side)
  const float* bias;                        // [c_out]
  int c_out, h_in, w_in, h_out, w_out;
  float* out;                               // [c_out][h_out][w_out]
  int split;                                // CTAs per cluster sharing one output tile (chunk c goes to cluster rank c % split)
  const float* gate_a;                      // data-gradient modes: src_a is a gradient, taken where gate_a > 0 (the ReLU of the kept output)
  int h_a, w_a;                             // data-gradient modes: stored size of src_a / gate_a (h_in / 2 when zero-inserted)
  int accumulate;                           // data-gradient modes: out += result (the skip connection's second gradient)
};

// One staged input element: the raw global loads (four taps when read through the upsample) stay in registers while the
// FFMA loop of the previous chunk runs; the interpolation arithmetic happens when the element is written to shared memory,
// so nothing waits on the loads before the loop.  Everything about an element that does not depend on the chunk (patch
// position, tap offsets, interpolation weights) is computed once per thread before the channel loop.
// Slot modes: 0 plain, 1 source A through the bilinear upsample, 2 / 3 data gradient of a stride-1 / stride-2 layer: the "input" is
// the output gradient behind its ReLU gate, for stride 2 spread over the even positions of a zero-filled image twice the size
// (a transposed convolution is a stride-1 convolution of that image with the flipped kernel).
constexpr int kSlotPlain = 0, kSlotUp = 1, kSlotGrad = 2, kSlotGradZi = 3;
template <int MODE> struct EncSlot;
template <> struct EncSlot<kSlotPlain> {
  int pk;                                    // (channel within chunk << 16) | y * w_in + x, or -1: padding / beyond the patch
  float v;
  __device__ __forceinline__ void setup(const EncConvParams& p, int i, int y, int x, bool in_patch) {
    pk = (in_patch && y >= 0 && y < p.h_in && x >= 0 && x < p.w_in) ? ((i << 16) | (y * p.w_in + x)) : -1;
  }
  __device__ __forceinline__ void load(const EncConvParams& p, int ic0, int cin) {
    const int ic = ic0 + (pk >> 16), yx = pk & 0xffff, hw = p.h_in * p.w_in;
    v = 0.f;
    if (pk >= 0 && ic < cin) v = ic < p.ca ? __ldg(p.src_a + ic * hw + yx) : __ldg(p.src_b + (ic - p.ca) * hw + yx);
  }
  __device__ __forceinline__ float value() const { return v; }
};
template <int MODE> struct EncSlot {      // kSlotGrad / kSlotGradZi
  int pk;                                    // (channel within chunk << 16) | offset inside one stored channel, or -1
  float v;
  __device__ __forceinline__ void setup(const EncConvParams& p, int i, int y, int x, bool in_patch) {
    bool ok = in_patch && y >= 0 && y < p.h_in && x >= 0 && x < p.w_in;
    if (MODE == kSlotGradZi) {
      ok = ok && !((y | x) & 1);
      y >>= 1; x >>= 1;
      ok = ok && y < p.h_a && x < p.w_a;
    }
    pk = ok ? ((i << 16) | (y * p.w_a + x)) : -1;
  }
  __device__ __forceinline__ void load(const EncConvParams& p, int ic0, int cin) {
    const int ic = ic0 + (pk >> 16), off = ic * (p.h_a * p.w_a) + (pk & 0xffff);
    v = 0.f;
    if (pk >= 0 && ic < cin) v = __ldg(p.gate_a + off) > 0.f ? __ldg(p.src_a + off) : 0.f;
  }
  __device__ __forceinline__ float value() const { return v; }
};
template <> struct EncSlot<kSlotUp> {
  int pk;                                    // as above (full-resolution offset, used for src_b channels)
  unsigned o0, o1;                           // low-resolution tap offsets of src_a: (y0,x0) | (y0,x1) << 16, (y1,x0) | (y1,x1) << 16
  float ly, lx;
  float a00, a01, a10, a11;
  __device__ __forceinline__ void setup(const EncConvParams& p, int i, int y, int x, bool in_patch) {
    pk = (in_patch && y >= 0 && y < p.h_in && x >= 0 && x < p.w_in) ? ((i << 16) | (y * p.w_in + x)) : -1;
    // PyTorch's upsample_bilinear2d source index (align_corners=False, scale 2)
    const int hs = p.h_in >> 1, ws = p.w_in >> 1;
    const float sy = fmaxf(0.5f * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (x + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
    ly = sy - y0; lx = sx - x0;
    o0 = pk >= 0 ? (unsigned)(y0 * ws + x0) | ((unsigned)(y0 * ws + x1) << 16) : 0u;
    o1 = pk >= 0 ? (unsigned)(y1 * ws + x0) | ((unsigned)(y1 * ws + x1) << 16) : 0u;
  }
  __device__ __forceinline__ void load(const EncConvParams& p, int ic0, int cin) {
    const int ic = ic0 + (pk >> 16);
    a00 = a01 = a10 = a11 = 0.f;
    if (pk >= 0 && ic < cin) {
      if (ic < p.ca) {
        const float* s = p.src_a + ic * ((p.h_in >> 1) * (p.w_in >> 1));
        a00 = __ldg(s + (o0 & 0xffffu)); a01 = __ldg(s + (o0 >> 16)); a10 = __ldg(s + (o1 & 0xffffu)); a11 = __ldg(s + (o1 >> 16));
      } else {
        a00 = __ldg(p.src_b + (ic - p.ca) * (p.h_in * p.w_in) + (pk & 0xffff));
      }
    }
  }
  __device__ __forceinline__ float value(const EncConvParams& p, int ic0) const {
    if (ic0 + (pk >> 16) >= p.ca) return a00;            // src_b channel (or padding: 0)
    return (1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11);
  }
};

template <int STRIDE, int MODE>
__global__ void __launch_bounds__(kEncThreads, 1) encoder_conv3x3_kernel(const EncConvParams p) {
  constexpr bool UP = MODE == kSlotUp;
  constexpr bool LINEAR = MODE >= kSlotGrad;            // data gradient: no bias, no ReLU
  constexpr int PH = (kEncTile - 1) * STRIDE + 3;       // input patch side: 10 (stride 1) / 17 (stride 2)
  constexpr int NCOL = 3 * STRIDE + 3;                  // input columns one thread touches per row: 6 / 9
  constexpr int IN_FLOATS = kEncIcb * PH * PH, W_FLOATS = kEncIcb * 9 * kEncOcb;
  constexpr int RED_FLOATS = (kEncGroups - 1) * 16 * kEncGroupThreads;
  constexpr int SMEM_FLOATS = IN_FLOATS + W_FLOATS > RED_FLOATS ? IN_FLOATS + W_FLOATS : RED_FLOATS;
  __shared__ __align__(16) float smem[SMEM_FLOATS];
  float (*s_w)[9][kEncOcb] = reinterpret_cast<float (*)[9][kEncOcb]>(smem);                       // [kEncIcb][9][kEncOcb]
  float (*s_in)[PH][PH] = reinterpret_cast<float (*)[PH][PH]>(smem + W_FLOATS);                   // [kEncIcb][PH][PH]: flat index = staging index
  const int tiles_x = (p.w_out + kEncTile - 1) / kEncTile;
  const int ty0 = (blockIdx.x / tiles_x) * kEncTile, tx0 = (blockIdx.x % tiles_x) * kEncTile;
  const int oc_base = blockIdx.y * kEncOcb;
  const int tid = threadIdx.x;
  const int grp = tid / kEncGroupThreads, t = tid % kEncGroupThreads;
  const int og = t & 7;                    // output-channel group: channels oc_base + 4*og .. +3
  const int pg = t >> 3;                   // pixel group 0..15: row pg >> 1, columns 4*(pg & 1) .. +3
  // zero-inserted gradient (stride-2 data gradient): three quarters of the patch are structural zeros.  Rows are dealt so that a warp
  // owns two rows of the SAME parity (0,2 | 1,3 | 4,6 | 5,7): the tap rows that meet only zero rows are skipped warp-uniformly, and
  // within a row the (pixel, tap column) pairs that meet zero columns are skipped at compile time - a quarter of the FMAs remain
  constexpr bool ZI = MODE == kSlotGradZi;
  const int wq = t >> 5, q4 = pg & 3;
  const int py = ZI ? 4 * (wq >> 1) + (wq & 1) + 2 * (q4 >> 1) : pg >> 1;
  const int px0 = (pg & 1) * 4;
  const int cin = p.ca + p.cb;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[j][o] = 0.f;

  // Register-staged software pipeline: the global loads of chunk c+1 (fixed trip counts, fully unrolled, so all of a thread's
  // loads are in flight at once) are issued before the FFMA loop of chunk c and land in shared memory after it.
  constexpr int N_IN = (IN_FLOATS + kEncThreads - 1) / kEncThreads;
  constexpr int N_W = (W_FLOATS / 4 + kEncThreads - 1) / kEncThreads;
  EncSlot<MODE> vin[N_IN];
  float4 vw[N_W];
  int woff[N_W];                             // (channel within chunk << 24) | offset inside the chunk's [i][9][c_out] weight rows, or -1
#pragma unroll
  for (int q = 0; q < N_IN; ++q) {
    const int e = tid + q * kEncThreads;
    const int i = e / (PH * PH), r = (e / PH) % PH, c = e % PH;
    vin[q].setup(p, i, ty0 * STRIDE - 1 + r, tx0 * STRIDE - 1 + c, e < IN_FLOATS);
  }
#pragma unroll
  for (int q = 0; q < N_W; ++q) {
    const int e = tid + q * kEncThreads;                    // float4 index: [i][k][o4]
    const int i = e / (9 * kEncOcb / 4), k = (e / (kEncOcb / 4)) % 9, o = (e % (kEncOcb / 4)) * 4;
    const int ks = LINEAR ? 8 - k : k;                      // data gradient: the kernel is flipped in both axes
    woff[q] = (e < W_FLOATS / 4 && oc_base + o < p.c_out) ? ((i << 24) | ((i * 9 + ks) * p.c_out + oc_base + o)) : -1;   // c_out % 4 == 0
  }
  auto load_chunk = [&](int ic0) {
#pragma unroll
    for (int q = 0; q < N_IN; ++q) vin[q].load(p, ic0, cin);
    const float* wbase = p.weight_t + ic0 * 9 * p.c_out;
#pragma unroll
    for (int q = 0; q < N_W; ++q) {
      vw[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (woff[q] >= 0 && ic0 + (woff[q] >> 24) < cin) vw[q] = __ldg(reinterpret_cast<const float4*>(wbase + (woff[q] & 0xffffff)));
    }
  };
  const int rank = p.split > 1 ? (int)blockIdx.z : 0;        // cluster dims are (1, 1, split) and gridDim.z == split
  const int ic_step = kEncIcb * p.split;
  load_chunk(rank * kEncIcb);
  for (int ic0 = rank * kEncIcb; ic0 < cin; ic0 += ic_step) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < N_IN; ++q) {
      const int e = tid + q * kEncThreads;
      if (e < IN_FLOATS) {
        if constexpr (UP) smem[W_FLOATS + e] = vin[q].value(p, ic0);
        else smem[W_FLOATS + e] = vin[q].value();
      }
    }
#pragma unroll
    for (int q = 0; q < N_W; ++q) {
      const int e = tid + q * kEncThreads;
      if (e < W_FLOATS / 4) reinterpret_cast<float4*>(smem)[e] = vw[q];
    }
    __syncthreads();
    if (ic0 + ic_step < cin) load_chunk(ic0 + ic_step);
    if (ic0 + grp * kEncIcPerGroup < cin) {                   // warp-uniform: skip channel slices beyond the input
#pragma unroll 2
      for (int ii = 0; ii < kEncIcPerGroup; ++ii) {
        const int i = grp * kEncIcPerGroup + ii;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          if (ZI && ((py + ky + 1) & 1)) continue;            // image row ty0 + py + ky - 1 is odd: a zero row (ty0 is a multiple of 8)
          float a[NCOL];
          const float* row = &s_in[i][py * STRIDE + ky][px0 * STRIDE];
#pragma unroll
          for (int c = 0; c < NCOL; ++c) a[c] = row[c];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 w = *reinterpret_cast<const float4*>(&s_w[i][ky * 3 + kx][4 * og]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (ZI && !((j + kx) & 1)) continue;            // image column tx0 + px0 + j + kx - 1 is odd: a zero column
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
  }
  // fixed-order reduction of the groups' partial sums through shared memory (deterministic)
  __syncthreads();
  if (grp > 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int o = 0; o < 4; ++o) smem[((grp - 1) * 16 + j * 4 + o) * kEncGroupThreads + t] = acc[j][o];
  }
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int g = 0; g < kEncGroups - 1; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[j][o] += smem[(g * 16 + j * 4 + o) * kEncGroupThreads + t];
  }
  // ... and of the cluster's CTAs (small images: the input channels are split over a cluster so more SMs work on the
  // layer) through distributed shared memory, summed by rank 0 in rank order
  if (p.split > 1) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __syncthreads();
    if (grp == 0 && rank > 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 0; o < 4; ++o) smem[(j * 4 + o) * kEncGroupThreads + t] = acc[j][o];
    }
    cluster.sync();
    if (grp == 0 && rank == 0) {
      for (int r = 1; r < p.split; ++r) {
        const float* remote = cluster.map_shared_rank(smem, r);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int o = 0; o < 4; ++o) acc[j][o] += remote[(j * 4 + o) * kEncGroupThreads + t];
      }
    }
    cluster.sync();                                           // remote shared memory stays alive until rank 0 has read it
  }
  if (grp > 0 || rank > 0) return;
  const int y = ty0 + py;
  if (y >= p.h_out) return;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int oc = oc_base + 4 * og + o;
    if (oc >= p.c_out) continue;
    const float b = LINEAR ? 0.f : __ldg(p.bias + oc);
    float* dst = p.out + ((long long)oc * p.h_out + y) * p.w_out + tx0 + px0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (tx0 + px0 + j < p.w_out) dst[j] = LINEAR ? (p.accumulate ? dst[j] + acc[j][o] : acc[j][o]) : fmaxf(acc[j][o] + b, 0.f);
  }
}

static int launch_enc(EncConvParams& p, int c_in, int mode, int stride, cudaStream_t stream) {
  if (p.c_out % 4 != 0 || (long long)p.h_in * p.w_in > 65536 || (long long)c_in * 9 * p.c_out >= (1 << 24)) return UORF_ERR_UNSUPPORTED;
  const int base_ctas = ((p.h_out + kEncTile - 1) / kEncTile) * ((p.w_out + kEncTile - 1) / kEncTile) * ((p.c_out + kEncOcb - 1) / kEncOcb);
  const int n_chunks = (c_in + kEncIcb - 1) / kEncIcb;
  p.split = 1;
  while (p.split < 8 && base_ctas * p.split * 2 <= 148 && p.split * 2 <= n_chunks) p.split *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((p.h_out + kEncTile - 1) / kEncTile) * ((p.w_out + kEncTile - 1) / kEncTile), (p.c_out + kEncOcb - 1) / kEncOcb, p.split);
  cfg.blockDim = dim3(kEncThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = p.split;
  cfg.attrs = attr;
  cfg.numAttrs = p.split > 1 ? 1 : 0;
  cudaError_t err;
  if (mode == kSlotGrad) err = cudaLaunchKernelEx(&cfg, encoder_conv3x3_kernel<1, kSlotGrad>, p);
  else if (mode == kSlotGradZi) err = cudaLaunchKernelEx(&cfg, encoder_conv3x3_kernel<1, kSlotGradZi>, p);
  else if (stride == 1 && mode == kSlotUp) err = cudaLaunchKernelEx(&cfg, encoder_conv3x3_kernel<1, kSlotUp>, p);
  else if (stride == 1) err = cudaLaunchKernelEx(&cfg, encoder_conv3x3_kernel<1, kSlotPlain>, p);
  else if (stride == 2 && mode == kSlotPlain) err = cudaLaunchKernelEx(&cfg, encoder_conv3x3_kernel<2, kSlotPlain>, p);
  else return UORF_ERR_UNSUPPORTED;
  if (err != cudaSuccess) return UORF_ERR_DEVICE;
  return finish_launch(1);
}

int launch_encoder_conv3x3(const float* src_a, int ca, int up_a, const float* src_b, int cb, const float* weight_t, const float* bias,
                           int c_out, int h_in, int w_in, int stride, float* out, cudaStream_t stream) {
  EncConvParams p = {};
  p.src_a = src_a; p.ca = ca; p.up_a = up_a;
  p.src_b = src_b; p.cb = cb;
  p.weight_t = weight_t; p.bias = bias;
  p.c_out = c_out; p.h_in = h_in; p.w_in = w_in;
  p.h_out = (h_in - 1) / stride + 1;
  p.w_out = (w_in - 1) / stride + 1;
  p.out = out;
  return launch_enc(p, ca + cb, up_a ? kSlotUp : kSlotPlain, stride, stream);
}

// Data gradient of one layer: grad_in [c_in][h_in][w_in] = transposed convolution of (grad_out where kept_out > 0), both
// [c_out][h_out][w_out], with the layer's kernel; weight_g = the Conv2d weight as [c_out][3][3][c_in] (the kernel flips the taps).
int launch_encoder_dgrad3x3(const float* grad_out, const float* kept_out, int c_out, const float* weight_g, int c_in, int h_in, int w_in,
                            int stride, int accumulate, float* grad_in, cudaStream_t stream) {
  EncConvParams p = {};
  p.accumulate = accumulate;
  p.src_a = grad_out; p.gate_a = kept_out; p.ca = c_out;
  p.h_a = (h_in - 1) / stride + 1; p.w_a = (w_in - 1) / stride + 1;
  p.weight_t = weight_g;
  p.c_out = c_in; p.h_in = h_in; p.w_in = w_in; p.h_out = h_in; p.w_out = w_in;
  p.out = grad_in;
  return launch_enc(p, c_out, stride == 1 ? kSlotGrad : kSlotGradZi, 1, stream);
}

// Backward of nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (model.py:27-32) as a gather: a low-resolution
// element collects its (up to) 4 x 4 high-resolution gradients with the separable weights 1/4, 3/4, 3/4, 1/4 (the clamped border
// rows / columns fold their outer weight into the 3/4).  One thread per element, fixed order: bit-reproducible, unlike the
// atomic scatter of the library op.
__global__ void encoder_upsample2x_bwd_kernel(const float* __restrict__ g_hi, int c, int h, int w, float* __restrict__ g_lo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)c * h * w) return;
  const int x = (int)(idx % w), y = (int)((idx / w) % h);
  const float* src = g_hi + (idx / ((long long)h * w)) * (4LL * h * w);
  const float wy[4] = {y > 0 ? 0.25f : 0.f, y > 0 ? 0.75f : 1.f, y < h - 1 ? 0.75f : 1.f, y < h - 1 ? 0.25f : 0.f};
  const float wx[4] = {x > 0 ? 0.25f : 0.f, x > 0 ? 0.75f : 1.f, x < w - 1 ? 0.75f : 1.f, x < w - 1 ? 0.25f : 0.f};
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int yy = 2 * y - 1 + a;
    if (yy < 0 || yy >= 2 * h) continue;
    float r = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int xx = 2 * x - 1 + b;
      if (xx >= 0 && xx < 2 * w) r = fmaf(wx[b], __ldg(src + (long long)yy * (2 * w) + xx), r);
    }
    s = fmaf(wy[a], r, s);
  }
  g_lo[idx] = s;
}

int launch_encoder_upsample2x_bwd(const float* grad_hi, int c, int h_lo, int w_lo, float* grad_lo, cudaStream_t stream) {
  const long long n = (long long)c * h_lo * w_lo;
  if (n <= 0) return UORF_OK;
  encoder_upsample2x_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(grad_hi, c, h_lo, w_lo, grad_lo);
  return finish_launch(1);
}

// Both weight layouts the layer kernels read, for every layer of the U-Net in one launch (the weights move every training step):
// Conv2d weight [c_out][c_in][3][3] -> forward [c_in][3][3][c_out] and data-gradient [c_out][3][3][c_in] (either may be null).
struct EncPrepTable { uorf_encoder_prep_layer l[8]; };
__global__ void encoder_weight_prep_kernel(const EncPrepTable t) {
  const uorf_encoder_prep_layer L = t.l[blockIdx.y];
  const int n = L.c_out * L.c_in * 9;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int k = e % 9, ci = (e / 9) % L.c_in, co = e / (9 * L.c_in);
    const float v = __ldg(L.weight + e);
    if (L.fwd_t) L.fwd_t[(ci * 9 + k) * L.c_out + co] = v;
    if (L.dgrad_t) L.dgrad_t[(co * 9 + k) * L.c_in + ci] = v;
  }
}

int launch_encoder_weight_prep(const uorf_encoder_prep_layer* layers, int n_layers, cudaStream_t stream) {
  if (n_layers < 1) return UORF_OK;
  if (n_layers > 8) return UORF_ERR_UNSUPPORTED;
  EncPrepTable t = {};
  for (int i = 0; i < n_layers; ++i) t.l[i] = layers[i];
  encoder_weight_prep_kernel<<<dim3(96, n_layers), 256, 0, stream>>>(t);
  return finish_launch(1);
}

// ---- weight / bias gradient ---------------------------------------------------------------------------------------
// dW[o][c][ky][kx] = sum over output pixels of G[o][y][x] * X[c][y*s + ky - 1][x*s + kx - 1], G = grad_out where kept_out > 0, X the
// layer's input read exactly as the forward reads it (two sources, source A optionally through the upsample; zero padding).
// CTA = 16 input channels x 32 output channels (4608 weights) over its share of the 8x8 output tiles; the four warp groups take two
// tile rows each; thread = one input channel x four output channels x nine taps (36 accumulators).  Partial sums of the pixel
// shares go to a workspace and are added in share order by the reduce kernel (deterministic).
constexpr int kEwPhp = 12;   // padded patch row (floats): rows start 16-byte aligned for stride 1
template <int STRIDE, int MODE>
__global__ void __launch_bounds__(kEncThreads, 1) encoder_wgrad3x3_kernel(const EncConvParams p, const float* __restrict__ grad_out,
                                                                           const float* __restrict__ kept_out, float* __restrict__ partial) {
  constexpr bool UP = MODE == kSlotUp;
  constexpr int PH = (kEncTile - 1) * STRIDE + 3;       // 10 / 17
  constexpr int PHP = STRIDE == 1 ? kEwPhp : 20;        // padded row length
  constexpr int NCOL = 7 * STRIDE + 3;                  // input columns of one tile row: 10 / 17
  constexpr int IN_ELEMS = kEncIcb * PH * PH;
  constexpr int N_IN = (IN_ELEMS + kEncThreads - 1) / kEncThreads;
  constexpr int N_G = kEncOcb * 64 / kEncThreads;        // 4
  constexpr int RED_FLOATS = 36 * kEncGroupThreads;
  constexpr int STAGE_FLOATS = kEncIcb * PH * PHP + 64 * kEncOcb;
  constexpr int SMEM_FLOATS = STAGE_FLOATS > RED_FLOATS ? STAGE_FLOATS : RED_FLOATS;
  __shared__ __align__(16) float smem[SMEM_FLOATS];
  float* s_in = smem;                                   // [16][PH][PHP]
  float* s_g = smem + kEncIcb * PH * PHP;               // [64 pixels][32 channels]
  const int tid = threadIdx.x;
  const int grp = tid / kEncGroupThreads, t = tid % kEncGroupThreads;
  const int og = t & 7, ci = t >> 3;
  const int ic0 = blockIdx.x * kEncIcb, oc_base = blockIdx.y * kEncOcb;
  const int cin = p.ca + p.cb;
  const int tiles_x = (p.w_out + kEncTile - 1) / kEncTile;
  const int n_tiles = tiles_x * ((p.h_out + kEncTile - 1) / kEncTile);
  float acc[4][9];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[o][k] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};

  EncSlot<MODE> vin[N_IN];
  float vg[N_G];
  auto load_tile = [&](int tile) {
    const int ty0 = (tile / tiles_x) * kEncTile, tx0 = (tile % tiles_x) * kEncTile;
#pragma unroll
    for (int q = 0; q < N_IN; ++q) {
      const int e = tid + q * kEncThreads;
      const int i = e / (PH * PH), r = (e / PH) % PH, c = e % PH;
      vin[q].setup(p, i, ty0 * STRIDE - 1 + r, tx0 * STRIDE - 1 + c, e < IN_ELEMS);
      vin[q].load(p, ic0, cin);
    }
#pragma unroll
    for (int q = 0; q < N_G; ++q) {
      const int e = tid + q * kEncThreads;              // [channel][pixel]: a tile row of 8 pixels is contiguous in global memory
      const int o = e >> 6, y = ty0 + ((e >> 3) & 7), x = tx0 + (e & 7);
      vg[q] = 0.f;
      if (oc_base + o < p.c_out && y < p.h_out && x < p.w_out) {
        const long long off = ((long long)(oc_base + o) * p.h_out + y) * p.w_out + x;
        vg[q] = __ldg(kept_out + off) > 0.f ? __ldg(grad_out + off) : 0.f;
      }
    }
  };
  const int share = blockIdx.z, n_share = gridDim.z;
  if (share < n_tiles) load_tile(share);
  for (int tile = share; tile < n_tiles; tile += n_share) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < N_IN; ++q) {
      const int e = tid + q * kEncThreads;
      if (e < IN_ELEMS) {
        const int i = e / (PH * PH), r = (e / PH) % PH, c = e % PH;
        float v;
        if constexpr (UP) v = vin[q].value(p, ic0);
        else v = vin[q].value();
        s_in[(i * PH + r) * PHP + c] = v;
      }
    }
#pragma unroll
    for (int q = 0; q < N_G; ++q) {
      const int e = tid + q * kEncThreads;
      s_g[(e & 63) * kEncOcb + (e >> 6)] = vg[q];
    }
    __syncthreads();
    if (tile + n_share < n_tiles) load_tile(tile + n_share);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = grp * 2 + rr;
      float4 g[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) g[x] = *reinterpret_cast<const float4*>(&s_g[(row * 8 + x) * kEncOcb + 4 * og]);
      if (ic0 == 0 && ci == 0) {
#pragma unroll
        for (int x = 0; x < 8; ++x) { bsum[0] += g[x].x; bsum[1] += g[x].y; bsum[2] += g[x].z; bsum[3] += g[x].w; }
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float a[NCOL];
        const float* src = &s_in[(ci * PH + row * STRIDE + ky) * PHP];
#pragma unroll
        for (int c = 0; c < NCOL; ++c) a[c] = src[c];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const float v = a[x * STRIDE + kx];
            acc[0][ky * 3 + kx] = fmaf(g[x].x, v, acc[0][ky * 3 + kx]);
            acc[1][ky * 3 + kx] = fmaf(g[x].y, v, acc[1][ky * 3 + kx]);
            acc[2][ky * 3 + kx] = fmaf(g[x].z, v, acc[2][ky * 3 + kx]);
            acc[3][ky * 3 + kx] = fmaf(g[x].w, v, acc[3][ky * 3 + kx]);
          }
        }
      }
    }
  }
  // fixed-order sum of the four row groups (one group at a time through shared memory), then the CTA's partial:
  // [share][c_out][cin][9] and the bias sums behind them
  for (int g = 1; g < kEncGroups; ++g) {
    __syncthreads();
    if (grp == g) {
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k = 0; k < 9; ++k) smem[(o * 9 + k) * kEncGroupThreads + t] = acc[o][k];
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[o][k] += smem[(o * 9 + k) * kEncGroupThreads + t];
    }
  }
  __syncthreads();
  if (ic0 == 0 && ci == 0 && grp > 0) {
#pragma unroll
    for (int o = 0; o < 4; ++o) smem[((grp - 1) * 4 + o) * 8 + og] = bsum[o];
  }
  __syncthreads();
  if (grp > 0) return;
  const long long w_elems = (long long)p.c_out * cin * 9;
  float* dst = partial + (long long)share * (w_elems + p.c_out);
  if (ic0 + ci < cin) {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int oc = oc_base + 4 * og + o;
      if (oc >= p.c_out) continue;
#pragma unroll
      for (int k = 0; k < 9; ++k) dst[((long long)oc * cin + ic0 + ci) * 9 + k] = acc[o][k];
    }
  }
  if (ic0 == 0 && ci == 0) {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int oc = oc_base + 4 * og + o;
      if (oc >= p.c_out) continue;
      float b = bsum[o];
#pragma unroll
      for (int g = 0; g < kEncGroups - 1; ++g) b += smem[(g * 4 + o) * 8 + og];
      dst[w_elems + oc] = b;
    }
  }
}

// grad_weight [c_out][cin][3][3] followed by grad_bias [c_out] <- sum of the shares in order
__global__ void encoder_wgrad_reduce_kernel(const float* __restrict__ partial, int n_share, long long n, long long n_w, float* __restrict__ gw,
                                            float* __restrict__ gb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < n_share; ++z) s += partial[(long long)z * n + i];
  if (i < n_w) gw[i] = s;
  else gb[i - n_w] = s;
}

static int wgrad_shares(int c_in, int c_out, int h_out, int w_out) {
  const int n_tiles = ((h_out + kEncTile - 1) / kEncTile) * ((w_out + kEncTile - 1) / kEncTile);
  const int base = ((c_in + kEncIcb - 1) / kEncIcb) * ((c_out + kEncOcb - 1) / kEncOcb);
  int s = 1;
  while (s * 2 <= n_tiles && base * s * 2 <= 148) s *= 2;   // at most one wave of CTAs: more tiles per CTA amortise staging, reduction and partial traffic
  return s;
}

size_t encoder_wgrad_workspace_floats(int c_in, int c_out, int h_in, int w_in, int stride) {
  const int h_out = (h_in - 1) / stride + 1, w_out = (w_in - 1) / stride + 1;
  return (size_t)wgrad_shares(c_in, c_out, h_out, w_out) * ((size_t)c_out * c_in * 9 + c_out);
}

int launch_encoder_wgrad3x3(const float* src_a, int ca, int up_a, const float* src_b, int cb, const float* grad_out, const float* kept_out,
                            int c_out, int h_in, int w_in, int stride, float* workspace, float* grad_weight, float* grad_bias,
                            cudaStream_t stream) {
  EncConvParams p = {};
  p.src_a = src_a; p.ca = ca; p.up_a = up_a;
  p.src_b = src_b; p.cb = cb;
  p.c_out = c_out; p.h_in = h_in; p.w_in = w_in;
  p.h_out = (h_in - 1) / stride + 1;
  p.w_out = (w_in - 1) / stride + 1;
  const int cin = ca + cb;
  if ((long long)h_in * w_in > 65536 || (stride == 2 && up_a)) return UORF_ERR_UNSUPPORTED;
  const int shares = wgrad_shares(cin, c_out, p.h_out, p.w_out);
  dim3 grid((cin + kEncIcb - 1) / kEncIcb, (c_out + kEncOcb - 1) / kEncOcb, shares);
  if (stride == 1 && up_a) encoder_wgrad3x3_kernel<1, kSlotUp><<<grid, kEncThreads, 0, stream>>>(p, grad_out, kept_out, workspace);
  else if (stride == 1) encoder_wgrad3x3_kernel<1, kSlotPlain><<<grid, kEncThreads, 0, stream>>>(p, grad_out, kept_out, workspace);
  else encoder_wgrad3x3_kernel<2, kSlotPlain><<<grid, kEncThreads, 0, stream>>>(p, grad_out, kept_out, workspace);
  const long long n_w = (long long)c_out * cin * 9, n = n_w + c_out;
  encoder_wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(workspace, shares, n, n_w, grad_weight, grad_bias);
  return finish_launch(2);
}

}  // namespace uorf
