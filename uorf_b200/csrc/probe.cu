// Test hook: one 128 x N x (16*ksteps) GEMM through exactly the tcgen05 data path the decoder uses
// (A rows packed to bf16 [hi, lo] and written to TMEM by their owning threads, B as a 128B-swizzled
// K-major shared-memory tile, fp32 accumulator in TMEM, 1 or 3 bf16 passes).  tests/ compares the result
// with a CPU product, which pins the descriptor / layout conventions independently of the decoder.
#include "common.cuh"
#include "decoder_layout.h"

namespace uorf {

__global__ void __launch_bounds__(160, 1) probe_gemm_kernel(const float* __restrict__ a, const float* __restrict__ b, int N,
                                                            int ksteps, int passes, int f16, float* __restrict__ d, int* err_flag) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_d;
  __shared__ uint32_t tmem_base_s;
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // [2 planes][64 rows][128 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_a, 128);
    mbar_init(&bar_d, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  // B tile: fp32 [N][64] -> bf16 hi/lo, swizzled
  for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
    const int n = e >> 6, k = e & 63;
    const float v = n < N ? b[n * 64 + k] : 0.f;
    const int off = (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2;
    const unsigned short hi = f16 ? to16<true>(v) : to16<false>(v);
    *reinterpret_cast<unsigned short*>(smem + off) = hi;
    const float res = v - (f16 ? from16<true>(hi) : from16<false>(hi));
    *reinterpret_cast<unsigned short*>(smem + kTileBytes + off) = f16 ? to16<true>(res) : to16<false>(res);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  constexpr int colD = 0, colHi = 64, colLo = 96;
  if (warp < 4) {
    const int t = threadIdx.x;
    const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (f16) split2<true>(a[t * 64 + 2 * i], a[t * 64 + 2 * i + 1], hi[i], lo[i]);
      else split2<false>(a[t * 64 + 2 * i], a[t * 64 + 2 * i + 1], hi[i], lo[i]);
    }
    tmem_st16(tb + colHi, hi);
    tmem_st16(tb + colHi + 16, hi + 16);
    tmem_st16(tb + colLo, lo);
    tmem_st16(tb + colLo + 16, lo + 16);
    tc_wait_st();
    tc_fence_before();
    mbar_arrive(&bar_a);
    mbar_wait(&bar_d, 0, err_flag, 2);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tb + colD + c0, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) d[t * N + c0 + i] = __uint_as_float(r[i]);
    }
  } else {
    mbar_wait(&bar_a, 0, err_flag, 1);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f32acc(128, N, f16 != 0);
      const uint32_t ws = smem_u32(smem);
      const uint64_t dhi = make_sdesc_sw128(ws), dlo = make_sdesc_sw128(ws + kTileBytes);
      bool first = true;
      for (int j = 0; j < ksteps; ++j) { mma_ts(tmem_base + colD, tmem_base + colHi + 8 * j, dhi + 2 * j, idesc, first ? 0u : 1u); first = false; }
      if (passes == 3) {
        for (int j = 0; j < ksteps; ++j) mma_ts(tmem_base + colD, tmem_base + colLo + 8 * j, dhi + 2 * j, idesc, 1u);
        for (int j = 0; j < ksteps; ++j) mma_ts(tmem_base + colD, tmem_base + colHi + 8 * j, dlo + 2 * j, idesc, 1u);
      }
      tc_commit(&bar_d);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, 256);
}

int launch_probe_gemm(const float* a, const float* b, int N, int ksteps, int passes, int f16, float* d, int* err_flag, cudaStream_t stream) {
  const int smem = 1024 + 2 * kTileBytes;
  probe_gemm_kernel<<<1, 160, smem, stream>>>(a, b, N, ksteps, passes, f16, d, err_flag);
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

}  // namespace uorf
