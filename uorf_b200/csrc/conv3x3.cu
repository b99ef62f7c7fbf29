// Kernel 6 — 3x3 convolution (stride 1, pad 1) as an implicit GEMM on tcgen05 / TMEM, fp32-parity by the 3-term 16-bit split.
//
// Replaces the cuDNN fp32 convolutions of the perceptual-loss network in the training step: VGG16 features up to relu4_3
// (reference models/model.py:316-324 `get_perceptual_net`, used at models/uorf_nogan_model.py:181-183) - forward on both images and
// the data gradient of the rendered branch.  The reference runs them in fp32 (util/util.py:210-211 switches TF32 off); cuDNN's
// fp32 engines reach ~12 % of the FFMA peak on these shapes and were 30 % of the training step.
//
//   y[p, co] = bias[co] + sum_{tap, ci} x[p + tap, ci] * w[co, ci, tap]        p = (n, h, w) NHWC pixel, tap in 3 x 3, zero padding
//
// GEMM view: M = pixels (128 per CTA, thread <-> pixel <-> TMEM lane), N = 64 output channels per CTA, K = 9 taps x C_in, walked in
// chunks of 64 input channels.  Operands are 16-bit pairs hi + lo prepared once per tensor (conv_split_kernel for activations,
// the host-side packer for weights):  x.w ~= x_hi.w_hi + x_lo.w_hi + x_hi.w_lo  with fp32 accumulation in TMEM.
//   forward:  fp16 pairs (~22 significand bits; activations and weights are O(1))
//   dgrad:    the same kernel on (grad_out * [y > 0]) with the flipped / transposed weight image, bf16 pairs (gradients of a
//             weighted loss are far below the fp16 range; bf16 keeps the fp32 exponent, ~16 significand bits)
// Pipeline: 4 producer warps (cp.async of the shifted input rows, zero fill outside the image; one bulk copy of the 16 KB weight
// tile pair) -> 4 shared-memory stages -> 1 issuer warp (12 tcgen05.mma per stage) -> accumulator in TMEM -> the producer warps
// turn into the epilogue (bias, ReLU, fp32 NHWC store).
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kCvTileM = 128;
constexpr int kCvTileN = 64;
constexpr int kCvStages = 4;
constexpr int kCvABytes = kCvTileM * 128;          // [128 pixels][64 x 16 bit], one plane
constexpr int kCvBBytes = kCvTileN * 128;          // [64 c_out][64 x 16 bit], one plane
constexpr int kCvStageBytes = 2 * kCvABytes + 2 * kCvBBytes;   // A hi | A lo | B hi | B lo = 48 KB
constexpr int kCvThreads = 5 * 32;
constexpr int kCvLag = 2;                          // cp.async groups in flight per producer thread
constexpr int kCvAcc = 8;                          // TMEM accumulators the K iterations rotate over (8 x 64 = 512 columns)

struct ConvParams {
  const unsigned short* x_hi;   // [pixels][c_in]
  const unsigned short* x_lo;
  const unsigned char* w_img;   // [c_out / 64][9][c_in / 64][hi 8 KB | lo 8 KB], tiles K-major, 128-byte swizzle
  const float* bias;            // [c_out] or null
  float* y;                     // [pixels][c_out]  (splits == 1), else partial sums [split][pixels][c_out] without bias / ReLU
  int n, h, w, c_in, c_out, relu;
  int splits;                   // K iterations divided over gridDim.z CTAs
  int* err_flag;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool F16>
__global__ void __launch_bounds__(kCvThreads, 1) conv3x3_x3_kernel(const ConvParams p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kCvStages], bar_empty[kCvStages], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t smem_a = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = p.c_in >> 6;
  const int iters_all = 9 * kchunks;
  // split-K: this CTA walks K iterations [it0, it0 + iters) (layers with few pixel tiles would otherwise leave most SMs idle)
  const int it0 = (int)(((long long)iters_all * blockIdx.z) / p.splits);
  const int iters = (int)(((long long)iters_all * (blockIdx.z + 1)) / p.splits) - it0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCvStages; ++s) {
      mbar_init(&bar_full[s], kCvTileM + 1);   // 128 producer threads (data landed + proxy fence) + the weight copy's expect_tx arrival
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(&tmem_base_s, kCvAcc * kCvTileN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 4) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_f32acc(128, kCvTileN, F16);
    for (int it = 0; it < iters; ++it) {
      const int s = it % kCvStages;
      mbar_wait_lean(smem_u32(&bar_full[s]), (it / kCvStages) & 1, p.err_flag, 600);
      tc_fence_after();
      const uint32_t st = smem_a + s * kCvStageBytes;
      const uint64_t a_hi = make_sdesc_sw128(st), a_lo = make_sdesc_sw128(st + kCvABytes);
      const uint64_t b_hi = make_sdesc_sw128(st + 2 * kCvABytes), b_lo = make_sdesc_sw128(st + 2 * kCvABytes + kCvBBytes);
      // K iterations rotate over kCvAcc accumulators: the tensor core rounds its fp32 accumulate toward zero, and one chain of
      // 9 * c_in / 16 steps (864 at c_in = 512) carried a bias of 3e-5 of the result; eight chains summed with round-to-nearest
      // adds in the epilogue bring it to the level of an fp32 convolution (measured against fp64)
      const uint32_t d = tmem_base + kCvTileN * (it % kCvAcc);
      mma_group_ss<4, 2, 2>(d, a_hi, b_hi, idesc, it >= kCvAcc ? 1u : 0u);
      mma_group_ss<4, 2, 2>(d, a_lo, b_hi, idesc, 1u);
      mma_group_ss<4, 2, 2>(d, a_hi, b_lo, idesc, 1u);
      tc_commit_elect(&bar_empty[s]);       // the stage may be refilled once these MMAs have read it
    }
    tc_commit_elect(&bar_acc);
    __syncwarp();
  } else {
    // =========================== producers, then epilogue ===========================
    // Copy mapping: 8 consecutive threads move the 8 16-byte chunks of ONE pixel row (128 contiguous bytes per plane), a warp
    // instruction covers 4 rows = 4 full lines; each thread serves rows (tid >> 3) + 16 i, i = 0..7.  (One thread per row put 32
    // different lines behind every instruction: 2600 cycles per K iteration, all of it address serialisation in the load unit.)
    const int t = threadIdx.x;                                   // epilogue: pixel row of the tile == TMEM lane
    const long long npix = (long long)p.n * p.h * p.w;
    const int chunk = t & 7;
    int r_h[8], r_w[8];
    long long r_pix[8];
    uint32_t r_dst[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = (t >> 3) + 16 * i;
      const long long px = (long long)blockIdx.x * kCvTileM + row;
      r_pix[i] = px < npix ? px : -1;
      r_w[i] = (int)(px % p.w);
      r_h[i] = (int)((px / p.w) % p.h);
      r_dst[i] = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((chunk ^ (row & 7)) & 7) << 4));
    }
    const unsigned char* w_tiles = p.w_img + ((size_t)blockIdx.y * iters_all + it0) * (2 * kCvBBytes);
    auto signal = [&](int it) {            // data of iteration `it` has landed: make it visible to the async proxy and arrive
      fence_proxy_async_smem();
      mbar_arrive(&bar_full[it % kCvStages]);
    };
    for (int it = 0; it < iters; ++it) {
      const int s = it % kCvStages;
      if (it >= kCvStages) mbar_wait_lean(smem_u32(&bar_empty[s]), ((it / kCvStages) - 1) & 1, p.err_flag, 610);
      const int tap = (it0 + it) / kchunks, kc = (it0 + it) - tap * kchunks;
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const long long shift = (long long)dy * p.w + dx;
      const uint32_t st = smem_a + s * kCvStageBytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = r_pix[i] >= 0 && (unsigned)(r_h[i] + dy) < (unsigned)p.h && (unsigned)(r_w[i] + dx) < (unsigned)p.w;
        const long long e = ok ? (r_pix[i] + shift) * p.c_in + kc * 64 + chunk * 8 : 0;
        const uint32_t nbytes = ok ? 16u : 0u;                   // zero fill outside the image
        cp_async16(st + r_dst[i], p.x_hi + e, nbytes);
        cp_async16(st + kCvABytes + r_dst[i], p.x_lo + e, nbytes);
      }
      cp_async_commit();
      if (t == 0) {                                              // weight tile pair of this iteration: one 16 KB bulk copy
        mbar_arrive_expect_tx(&bar_full[s], 2 * kCvBBytes);
        bulk_g2s(smem + s * kCvStageBytes + 2 * kCvABytes, w_tiles + (size_t)it * (2 * kCvBBytes), 2 * kCvBBytes, &bar_full[s]);
      }
      if (it >= kCvLag) {
        cp_async_wait<kCvLag>();
        signal(it - kCvLag);
      }
    }
    cp_async_wait<0>();
    for (int it = iters > kCvLag ? iters - kCvLag : 0; it < iters; ++it) signal(it);

    // ---- epilogue: sum of the accumulators -> bias -> ReLU -> fp32 NHWC ----
    mbar_wait_lean(smem_u32(&bar_acc), 0, p.err_flag, 620);
    tc_fence_after();
    const long long pix = (long long)blockIdx.x * kCvTileM + t;
    const bool row_ok = pix < npix;
    const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* dst = p.y + ((long long)blockIdx.z * npix + pix) * p.c_out + blockIdx.y * kCvTileN;
    const bool final_out = p.splits == 1;
    const float* bias = (p.bias && final_out) ? p.bias + blockIdx.y * kCvTileN : nullptr;
    const int nacc = iters < kCvAcc ? iters : kCvAcc;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];
      {
        uint32_t r[32];
        tmem_ld32(tb + 32 * half, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      }
      for (int a = 1; a < nacc; ++a) {      // round-to-nearest adds of the partial accumulators
        uint32_t r[32];
        tmem_ld32(tb + kCvTileN * a + 32 * half, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(r[i]);
      }
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + 32 * half + 4 * q));
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (p.relu && final_out) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          *reinterpret_cast<float4*>(dst + 32 * half + 4 * q) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, kCvAcc * kCvTileN);
}

// x (fp32) [* (gate > 0)] -> 16-bit hi + lo planes.  n multiple of 2.
template <bool F16>
__global__ void __launch_bounds__(256) conv_split_kernel(const float* __restrict__ x, const float* __restrict__ gate, unsigned int* __restrict__ hi,
                                                         unsigned int* __restrict__ lo, long long npairs) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += (long long)gridDim.x * blockDim.x) {
    float2 v = __ldg(reinterpret_cast<const float2*>(x) + i);
    if (gate != nullptr) {
      const float2 g = __ldg(reinterpret_cast<const float2*>(gate) + i);
      if (!(g.x > 0.f)) v.x = 0.f;
      if (!(g.y > 0.f)) v.y = 0.f;
    }
    uint32_t h, l;
    split2<F16>(v.x, v.y, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// nn.MaxPool2d(2, 2) folded into the operand split of the NEXT layer: x NHWC [n][h][w][c] -> hi / lo planes [n][h/2][w/2][c] of the
// pooled tensor, which itself never reaches memory (its only reader is that layer).  One thread per window and channel pair.
template <bool F16>
__global__ void __launch_bounds__(256) conv_split_pool_kernel(const float* __restrict__ x, unsigned int* __restrict__ hi, unsigned int* __restrict__ lo,
                                                              int n, int h, int w, int c) {
  const unsigned cp = c >> 1, wo = w >> 1, ho = h >> 1;
  const unsigned total = (unsigned)n * ho * wo * cp;                    // the launcher keeps n h w c below 2^31: 32-bit index arithmetic
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned win = i / cp, ch = i - win * cp;
    const unsigned row = win / wo, ox = win - row * wo;
    const unsigned img = row / ho, oy = row - img * ho;
    const float2* src = reinterpret_cast<const float2*>(x) + ((size_t)(img * h + 2 * oy) * w + 2 * ox) * cp + ch;
    const float2 a = __ldg(src), b = __ldg(src + cp), d = __ldg(src + (size_t)w * cp), e = __ldg(src + (size_t)w * cp + cp);
    uint32_t vh, vl;
    split2<F16>(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)), vh, vl);
    hi[i] = vh;
    lo[i] = vl;
  }
}

// ... and its backward folded into the split that feeds the data gradient of the layer BEFORE the pool: g [n][h/2][w/2][c] (gradient
// of the pooled tensor), x [n][h][w][c] (the pool's input = that layer's ReLU output) -> hi / lo planes [n][h][w][c] of
// (g routed to the first maximum of each window, where x > 0).  First maximum in row-major window order, as torch's max-pool picks.
template <bool F16>
__global__ void __launch_bounds__(256) conv_split_unpool_kernel(const float* __restrict__ g, const float* __restrict__ x, unsigned int* __restrict__ hi,
                                                                unsigned int* __restrict__ lo, int n, int h, int w, int c) {
  const unsigned cp = c >> 1, wo = w >> 1, ho = h >> 1;
  const unsigned total = (unsigned)n * ho * wo * cp;                    // the launcher keeps n h w c below 2^31: 32-bit index arithmetic
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned win = i / cp, ch = i - win * cp;
    const unsigned row = win / wo, ox = win - row * wo;
    const unsigned img = row / ho, oy = row - img * ho;
    const size_t base = ((size_t)(img * h + 2 * oy) * w + 2 * ox) * cp + ch;      // float2 / word index of the window's first pixel
    const float2* src = reinterpret_cast<const float2*>(x) + base;
    const float2 v[4] = {__ldg(src), __ldg(src + cp), __ldg(src + (size_t)w * cp), __ldg(src + (size_t)w * cp + cp)};
    const float2 gv = __ldg(reinterpret_cast<const float2*>(g) + i);
    int ax = 0, ay = 0;
    float mx = v[0].x, my = v[0].y;
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      if (v[q].x > mx) { mx = v[q].x; ax = q; }
      if (v[q].y > my) { my = v[q].y; ay = q; }
    }
    const float ox_ = mx > 0.f ? gv.x : 0.f, oy_ = my > 0.f ? gv.y : 0.f;      // ReLU gate of the layer that produced x
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t vh, vl;
      split2<F16>(q == ax ? ox_ : 0.f, q == ay ? oy_ : 0.f, vh, vl);
      const size_t dst = base + (q & 1) * cp + (q >> 1) * (size_t)w * cp;
      hi[dst] = vh;
      lo[dst] = vl;
    }
  }
}

// split-K: y = [relu](bias + sum of the partial sums), fixed order (deterministic)
__global__ void __launch_bounds__(256) conv_reduce_kernel(const float4* __restrict__ part, int splits, long long n4, const float* __restrict__ bias,
                                                          int c_out, int relu, float4* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = __ldg(part + i);
    for (int s = 1; s < splits; ++s) {
      const float4 b = __ldg(part + (long long)s * n4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (i % (c_out / 4)));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    y[i] = a;
  }
}

int conv3x3_splits(int n, int h, int w, int c_in, int c_out, int num_sms) {
  const long long npix = (long long)n * h * w;
  const long long ctas = ((npix + kCvTileM - 1) / kCvTileM) * (c_out / kCvTileN);
  const int iters = 9 * (c_in / 64);
  int s = 1;
  while (s < 8 && ctas * s * 2 <= num_sms && iters / (s * 2) >= 9) s *= 2;   // keep >= 9 K iterations per CTA
  return s;
}

int launch_conv_split(const float* x, const float* gate, void* hi, void* lo, long long n, int f16, int num_sms, cudaStream_t stream) {
  if (n <= 0) return UORF_OK;
  const long long npairs = n / 2;
  long long blocks = (npairs + 255) / 256;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  if (f16) conv_split_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(x, gate, static_cast<unsigned int*>(hi), static_cast<unsigned int*>(lo), npairs);
  else conv_split_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(x, gate, static_cast<unsigned int*>(hi), static_cast<unsigned int*>(lo), npairs);
  return finish_launch(1);
}

int launch_conv_split_pool(const float* x, const float* g, void* hi, void* lo, int n, int h, int w, int c, int f16, int num_sms, cudaStream_t stream) {
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 2);
  if (total <= 0) return UORF_OK;
  if ((long long)n * h * w * c >= (1LL << 31)) return UORF_ERR_UNSUPPORTED;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  unsigned int* ph = static_cast<unsigned int*>(hi);
  unsigned int* pl = static_cast<unsigned int*>(lo);
  if (g == nullptr) {
    if (f16) conv_split_pool_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(x, ph, pl, n, h, w, c);
    else conv_split_pool_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(x, ph, pl, n, h, w, c);
  } else {
    if (f16) conv_split_unpool_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(g, x, ph, pl, n, h, w, c);
    else conv_split_unpool_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(g, x, ph, pl, n, h, w, c);
  }
  return finish_launch(1);
}

int launch_conv3x3_x3(const void* x_hi, const void* x_lo, const void* w_img, const float* bias, float* y, float* partials, int splits, int n,
                      int h, int w, int c_in, int c_out, int relu, int f16, int num_sms, int* err_flag, cudaStream_t stream) {
  ConvParams p;
  p.x_hi = static_cast<const unsigned short*>(x_hi);
  p.x_lo = static_cast<const unsigned short*>(x_lo);
  p.w_img = static_cast<const unsigned char*>(w_img);
  p.bias = bias;
  if (splits < 1 || (splits > 1 && partials == nullptr)) splits = 1;
  p.y = splits > 1 ? partials : y;
  p.splits = splits;
  p.n = n; p.h = h; p.w = w; p.c_in = c_in; p.c_out = c_out; p.relu = relu;
  p.err_flag = err_flag;
  const long long npix = (long long)n * h * w;
  if (npix <= 0) return UORF_OK;
  const size_t smem = 1024 + (size_t)kCvStages * kCvStageBytes;
  dim3 grid((unsigned)((npix + kCvTileM - 1) / kCvTileM), (unsigned)(c_out / kCvTileN), (unsigned)splits);
  auto launch = [&](auto kern) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UORF_ERR_DEVICE;
    kern<<<grid, kCvThreads, smem, stream>>>(p);
    return UORF_OK;
  };
  const int rc = f16 ? launch(conv3x3_x3_kernel<true>) : launch(conv3x3_x3_kernel<false>);
  if (rc != UORF_OK) return rc;
  if (splits > 1) {
    const long long n4 = npix * c_out / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
    conv_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(partials), splits, n4, bias, c_out, relu,
                                                            reinterpret_cast<float4*>(y));
    return finish_launch(2);
  }
  return finish_launch(1);
}

// ---- 3-channel stem of the perceptual network (conv1_1 of VGG16: 3 -> c_out channels, 3x3, pad 1) in fp32 FFMA --------------
// Forward: x NCHW [n][3][h][w] -> y NHWC [n][h][w][c_out] = relu(conv + bias) (the layout the tensor-core layers read).
// 27 inputs per pixel: a thread owns one pixel and four output channels; the 16 threads of a pixel share its loads through L1.
// weight [c_out][3][3][3] (the Conv2d tensor as it is), c_out a multiple of 4 and <= 64.
__global__ void __launch_bounds__(256) vgg_stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias,
                                                            int n, int h, int w, int c_out, float* __restrict__ y) {
  __shared__ float s_w[27][64];
  for (int e = threadIdx.x; e < 27 * c_out; e += blockDim.x) s_w[e % 27][e / 27] = weight[e];     // [o][c][ky][kx] -> [c ky kx][o]
  __syncthreads();
  const int groups = c_out >> 2;
  const unsigned npix = (unsigned)n * h * w, hw = (unsigned)h * w;      // the launcher keeps n h w c_out below 2^31: 32-bit index arithmetic
  // grid-stride: the weight tile is staged once per block, not once per 16 pixels
  for (unsigned gid = blockIdx.x * blockDim.x + threadIdx.x; gid < npix * groups; gid += gridDim.x * blockDim.x) {
    const unsigned pix = gid / groups;
    const int o4 = (int)(gid % groups) * 4;
    const unsigned img = pix / hw, rem = pix - img * hw;
    const int py = (int)(rem / w), px = (int)(rem - (rem / w) * w);
    const float* src = x + (size_t)img * 3 * hw;
    float a0 = bias[o4], a1 = bias[o4 + 1], a2 = bias[o4 + 2], a3 = bias[o4 + 3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py + ky - 1, xx = px + kx - 1;
          const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(src + (c * h + yy) * w + xx) : 0.f;
          const float4 wv = *reinterpret_cast<const float4*>(&s_w[(c * 3 + ky) * 3 + kx][o4]);
          a0 = fmaf(v, wv.x, a0); a1 = fmaf(v, wv.y, a1); a2 = fmaf(v, wv.z, a2); a3 = fmaf(v, wv.w, a3);
        }
    *reinterpret_cast<float4*>(y + (size_t)pix * c_out + o4) = make_float4(fmaxf(a0, 0.f), fmaxf(a1, 0.f), fmaxf(a2, 0.f), fmaxf(a3, 0.f));
  }
}

// Data gradient: g NHWC [n][h][w][c_out] taken where the kept output y > 0 -> gx NCHW [n][3][h][w].  One warp per pixel: a lane owns
// output channels lane, lane + 32 and three partial sums (one per input channel), folded over the taps in registers and over the
// lanes by a fixed shuffle tree (bit-reproducible).
__global__ void __launch_bounds__(256) vgg_stem_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ weight,
                                                              int n, int h, int w, int c_out, float* __restrict__ gx) {
  __shared__ float s_w[27][64];
  for (int e = threadIdx.x; e < 27 * 64; e += blockDim.x) s_w[e % 27][e / 27] = (e / 27) < c_out ? weight[e] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned npix = (unsigned)n * h * w, hw = (unsigned)h * w;      // below 2^31 / c_out (launcher): 32-bit index arithmetic
  const unsigned wpb = blockDim.x >> 5;
  const bool second = lane + 32 < c_out;
  for (unsigned pix = blockIdx.x * wpb + (threadIdx.x >> 5); pix < npix; pix += gridDim.x * wpb) {   // warp-uniform
    const unsigned img = pix / hw, rem = pix - img * hw;
    const int py = (int)(rem / w), px = (int)(rem - (rem / w) * w);
    // all 36 loads of the pixel are issued before the first use (one round trip instead of eighteen)
    float yv[9][2], gv[9][2];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + 1 - t / 3, xx = px + 1 - t % 3;          // the output pixel whose tap t read this input pixel
      const bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
      const size_t off = ((size_t)img * hw + (unsigned)(yy * w + xx)) * c_out + lane;
      yv[t][0] = in && lane < c_out ? __ldg(y + off) : 0.f;
      gv[t][0] = in && lane < c_out ? __ldg(g + off) : 0.f;
      yv[t][1] = in && second ? __ldg(y + off + 32) : 0.f;
      gv[t][1] = in && second ? __ldg(g + off + 32) : 0.f;
    }
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float v = yv[t][half] > 0.f ? gv[t][half] : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = fmaf(v, s_w[c * 9 + t][lane + 32 * half], acc[c]);
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = acc[c];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane == 0) gx[((size_t)img * 3 + c) * hw + rem] = v;
    }
  }
}

int launch_vgg_stem_fwd(const float* x, const float* weight, const float* bias, int n, int h, int w, int c_out, float* y, cudaStream_t stream) {
  if (c_out % 4 != 0 || c_out > 64 || c_out < 4) return UORF_ERR_UNSUPPORTED;
  const long long threads = (long long)n * h * w * (c_out / 4);
  if (threads <= 0) return UORF_OK;
  if ((long long)n * h * w * c_out >= (1LL << 31)) return UORF_ERR_UNSUPPORTED;
  long long blocks = (threads + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  vgg_stem_fwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, weight, bias, n, h, w, c_out, y);
  return finish_launch(1);
}

int launch_vgg_stem_dgrad(const float* g, const float* y, const float* weight, int n, int h, int w, int c_out, float* gx, cudaStream_t stream) {
  if (c_out > 64 || c_out < 1) return UORF_ERR_UNSUPPORTED;
  const long long npix = (long long)n * h * w;
  if (npix <= 0) return UORF_OK;
  if (npix * c_out >= (1LL << 31)) return UORF_ERR_UNSUPPORTED;
  long long blocks = (npix + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  vgg_stem_dgrad_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g, y, weight, n, h, w, c_out, gx);
  return finish_launch(1);
}

}  // namespace uorf
