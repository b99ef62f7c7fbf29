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

// ---------------------------------------------------------------------------------------------------
// Latency probe (test/diagnostic hook): cycle counts of the building blocks of the decoder's handshake,
// measured with clock64() on one SM.  out[0..] = see the labels in scripts/probe_timing.py.
// ---------------------------------------------------------------------------------------------------
namespace uorf {

__global__ void __launch_bounds__(160, 1) probe_timing_kernel(long long* out, int reps) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_d;
  __shared__ uint32_t tmem_base_s;
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_a, 128);
    mbar_init(&bar_d, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  for (int e = threadIdx.x; e < 8 * kTileBytes / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_f32acc(128, 64, true);
  const uint64_t dhi = make_sdesc_sw128(smem_u32(smem));
  uint32_t na = 0, nd = 0;
  // ---- experiment A: 48 back-to-back MMAs (elect-issued, uniform operands) + commit; cycles until the commit lands ----
  // variant v: 0: N=64 TS same D | 1: N=64 TS 2 rotating D | 2: N=64 TS 4 rotating D | 3: N=128 TS same D | 4: N=256 TS same D
  //            5: N=64 SS same D | 6: N=64 SS 4 rotating D | 7: N=32 TS same D
  if (warp == 4) {
    const uint32_t tbu = __shfl_sync(0xffffffffu, tmem_base, 0);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int N = v == 3 ? 128 : v == 4 ? 256 : v == 7 ? 32 : 64;
      const int ND = v == 1 ? 2 : (v == 2 || v == 6) ? 4 : 1;
      const bool ss = v == 5 || v == 6;
      const uint32_t id = make_idesc_f32acc(128, N, true);
      long long best = 1LL << 60, best_issue = 1LL << 60;
      for (int r = 0; r < reps; ++r) {
        const long long t0 = clock64();
#pragma unroll
        for (int j = 0; j < 48; ++j) {
          const uint32_t d = tbu + (j % ND) * 64;
          const uint32_t acc = j >= ND ? 1u : 0u;
          if (ss) mma_ss_elect(d, dhi + 2 * (j & 3), dhi + 512 + 2 * (j & 3), id, acc);
          else mma_ts_elect(d, tbu + 256 + 8 * (j & 3), dhi + 2 * (j & 3), id, acc);
        }
        tc_commit_elect(&bar_d);
        const long long t1 = clock64();
        mbar_wait(&bar_d, nd & 1);
        ++nd;
        const long long t2 = clock64();
        best = min(best, t2 - t0);
        best_issue = min(best_issue, t1 - t0);
      }
      if (lane == 0) { out[32 + v * 2] = best; out[32 + v * 2 + 1] = best_issue; }
    }
  }
  __syncthreads();
  // ---- experiment B: epilogue pieces on 4 warps ----
  if (warp < 4) {
    const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
    long long b_ld = 1LL << 60, b_cmp = 1LL << 60, b_st = 1LL << 60, b_fence = 1LL << 60, b_all = 1LL << 60;
    float bias = 0.01f * lane;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      uint32_t v[32], hi[16], lo[16];
      tmem_ld32(tb, v);
      tc_wait_ld();
      const long long t1 = clock64();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float a = fmaxf(__uint_as_float(v[2 * q]) + bias, 0.f), b = fmaxf(__uint_as_float(v[2 * q + 1]) + bias, 0.f);
        split2<true>(a, b, hi[q], lo[q]);
      }
      const long long t2 = clock64();
      tmem_st16(tb + 64, hi);
      tmem_st16(tb + 96, lo);
      tc_wait_st();
      const long long t3 = clock64();
      tc_fence_before();
      mbar_arrive(&bar_a);
      const long long t4 = clock64();
      b_ld = min(b_ld, t1 - t0); b_cmp = min(b_cmp, t2 - t1); b_st = min(b_st, t3 - t2); b_fence = min(b_fence, t4 - t3);
      b_all = min(b_all, t4 - t0);
      mbar_wait(&bar_a, na & 1);
      ++na;
    }
    if (threadIdx.x == 0) { out[16] = b_ld; out[17] = b_cmp; out[18] = b_st; out[19] = b_fence; out[20] = b_all; }
  }
  __syncthreads();
  // ---- experiment C: full round trip  epilogue arrive -> issuer wake -> 4 MMAs -> commit -> epilogue wake ----
  {
    long long best = 1LL << 60;
    for (int r = 0; r < reps; ++r) {
      if (warp < 4) {
        const long long t0 = clock64();
        tc_fence_before();
        mbar_arrive(&bar_a);
        mbar_wait(&bar_d, nd & 1);
        tc_fence_after();
        const long long t1 = clock64();
        best = min(best, t1 - t0);
      } else {
        mbar_wait(&bar_a, na & 1);
        tc_fence_after();
        if (lane == 0) {
          for (int j = 0; j < 4; ++j) mma_ts(tmem_base, tmem_base + 64 + 8 * j, dhi + 2 * j, idesc, j ? 1u : 0u);
          tc_commit(&bar_d);
        }
        __syncwarp();
      }
      ++na;
      ++nd;
      __syncthreads();
    }
    if (threadIdx.x == 0) out[24] = best;
  }
  // ---- experiments D..G: do TMEM loads / stores / plain ALU work of the epilogue warps and a running MMA stream slow each other?
  //   mode 0: LDTM loop (2 x ld32 + wait = one 128 x 64 fp32 accumulator read per iteration)   mode 1: STTM loop (4 x st16 + wait)
  //   mode 2: ALU loop (16 hi/lo splits on registers, no TMEM access)
  //   each mode is timed alone (out[64 + 4m]) and while warp 4 streams 192 N=64 TS MMAs (out[65 + 4m]); the MMA stream is timed
  //   alone (out[80]) and against each mode (out[66 + 4m]).  64 iterations per loop; cycles for the whole loop.
  {
    const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tbu = __shfl_sync(0xffffffffu, tmem_base, 0);
    for (int pass = 0; pass < 7; ++pass) {          // 0: MMA alone; 1..3: mode alone; 4..6: mode + MMA
      const int mode = (pass - 1) % 3;
      const bool with_mma = pass == 0 || pass >= 4, with_epi = pass >= 1;
      long long best_e = 1LL << 60, best_m = 1LL << 60;
      for (int r = 0; r < 5; ++r) {
        __syncthreads();
        if (warp == 4 && with_mma) {
          const long long t0 = clock64();
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int j = 0; j < 48; ++j) mma_ts_elect(tbu + 384, tbu + 256 + 8 * (j & 3), dhi + 2 * (j & 3), idesc, (g | j) ? 1u : 0u);
          }
          tc_commit_elect(&bar_d);
          mbar_wait(&bar_d, nd & 1);
          best_m = min(best_m, clock64() - t0);
        }
        if (with_mma) ++nd;
        if (warp < 4 && with_epi) {
          uint32_t v[32], hi[16], lo[16];
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(0.001f * (q + lane));
          const long long t0 = clock64();
#pragma unroll 1
          for (int i = 0; i < 64; ++i) {
            if (mode == 0) {
              tmem_ld32(tb, v);
              tc_wait_ld();
              uint32_t w[32];
              tmem_ld32(tb + 32, w);
              tc_wait_ld();
              v[0] ^= w[0];
            } else if (mode == 1) {
              tmem_st16(tb + 64, v); tmem_st16(tb + 80, v + 16); tmem_st16(tb + 96, v); tmem_st16(tb + 112, v + 16);
              tc_wait_st();
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) split2_relu<true>(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]), hi[q], lo[q]);
#pragma unroll
              for (int q = 0; q < 16; ++q) { v[2 * q] ^= hi[q] & 0x00010001u; v[2 * q + 1] ^= lo[q] & 0x00010001u; }
            }
          }
          const long long t1 = clock64();
          if (v[0] == 0xdeadbeefu) out[200] = 1;
          best_e = min(best_e, t1 - t0);
        }
      }
      if (threadIdx.x == 0 && with_epi) out[64 + 4 * mode + (with_mma ? 1 : 0)] = best_e;
      if (warp == 4 && lane == 0 && with_mma) out[pass == 0 ? 80 : 66 + 4 * mode] = best_m;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

int launch_probe_timing(long long* out, int reps, cudaStream_t stream) {
  const int smem = 1024 + 8 * kTileBytes;
  cudaFuncSetAttribute(probe_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_timing_kernel<<<1, 160, smem, stream>>>(out, reps);
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

}  // namespace uorf

// ---------------------------------------------------------------------------------------------------
// Test hook for the weight-gradient data path: D[m][n] = sum_p A[p][m] * B[p][n]  (p < 128 points, m, n < 64)
// with BOTH operands read from [point][feature] shared-memory tiles through MN-major descriptors, M = 64, and two
// accumulators interleaved in the same TMEM columns (lane offset 16).  d1 = A^T B1, d2 = A^T B2.
// ---------------------------------------------------------------------------------------------------
namespace uorf {

// smem matrix descriptor, MN-major operand, 128B swizzle: rows (K index = point) are 128 B = 64 features;
// 8-row groups 1024 B apart (SBO); LBO = distance between 64-feature blocks (single block here).
__device__ __forceinline__ uint64_t make_sdesc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4)) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}

__global__ void __launch_bounds__(160, 1) probe_gemm_tn_kernel(const float* __restrict__ a, const float* __restrict__ b1,
                                                               const float* __restrict__ b2, float* __restrict__ d1,
                                                               float* __restrict__ d2, int* err_flag) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_d;
  __shared__ uint32_t tmem_base_s;
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // 3 tiles [128 rows][128 B]
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_d, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(&tmem_base_s, 64);
    tmem_relinquish();
  }
  for (int e = threadIdx.x; e < 3 * 128 * 64; e += blockDim.x) {
    const int which = e / (128 * 64), r = (e / 64) % 128, c = e % 64;
    const float v = (which == 0 ? a : which == 1 ? b1 : b2)[r * 64 + c];
    const int off = which * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 3) ^ (r & 7)) & 7) << 4) + (c & 7) * 2;
    *reinterpret_cast<unsigned short*>(smem + off) = to16<false>(v);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 4) {
    const uint32_t tbu = __shfl_sync(0xffffffffu, tmem_base, 0);
    // M = 64, N = 64, bf16, both operands MN-major (bits 15, 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
    const uint32_t s0 = smem_u32(smem);
    const uint64_t da = make_sdesc_mn_sw128(s0, 16384), db1 = make_sdesc_mn_sw128(s0 + 16384, 16384),
                   db2 = make_sdesc_mn_sw128(s0 + 32768, 16384);
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // K-step = 16 points = 2048 B
      mma_ss_elect(tbu, da + 128 * j, db1 + 128 * j, idesc, j ? 1u : 0u);
      mma_ss_elect(tbu + (16u << 16), da + 128 * j, db2 + 128 * j, idesc, j ? 1u : 0u);
    }
    tc_commit_elect(&bar_d);
    __syncwarp();
  } else {
    mbar_wait(&bar_d, 0, err_flag, 3);
    tc_fence_after();
    const int lane = threadIdx.x & 31;
    const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* dst = (lane < 16 ? d1 : d2) + (warp * 16 + (lane & 15)) * 64;
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tb + c0, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) dst[c0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, 64);
}

int launch_probe_gemm_tn(const float* a, const float* b1, const float* b2, float* d1, float* d2, int* err_flag, cudaStream_t stream) {
  const int smem = 1024 + 3 * 16384;
  cudaFuncSetAttribute(probe_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_gemm_tn_kernel<<<1, 160, smem, stream>>>(a, b1, b2, d1, d2, err_flag);
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

}  // namespace uorf

// ---------------------------------------------------------------------------------------------------
// extern "C" surface of libuorf_b200_probe.so (include/uorf_b200_probe.h).  This file is built into its own TEST library:
// the product library libuorf_b200.so contains none of these kernels.
// ---------------------------------------------------------------------------------------------------
#include "../../include/uorf_b200_probe.h"

namespace uorf {
int finish_launch(int) { return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE; }   // the probe library's own copy
}

extern "C" {

int uorf_probe_gemm(const float* a, const float* b, int32_t N, int32_t ksteps, int32_t precision, float* d, void* stream) {
  if (!a || !b || !d) return UORF_ERR_ARG;
  if (N < 16 || N > 64 || (N % 16) || ksteps < 1 || ksteps > 4 || precision < 1 || precision > 4) return UORF_ERR_ARG;
  return uorf::launch_probe_gemm(a, b, N, ksteps, precision >= 3 ? 3 : 1, (precision % 2) == 0 ? 1 : 0, d, nullptr, static_cast<cudaStream_t>(stream));
}

int uorf_probe_gemm_tn(const float* a, const float* b1, const float* b2, float* d1, float* d2, void* stream) {
  if (!a || !b1 || !b2 || !d1 || !d2) return UORF_ERR_ARG;
  return uorf::launch_probe_gemm_tn(a, b1, b2, d1, d2, nullptr, static_cast<cudaStream_t>(stream));
}

int uorf_probe_timing(int64_t* out, int32_t reps, void* stream) {
  if (!out || reps < 1) return UORF_ERR_ARG;
  return uorf::launch_probe_timing(reinterpret_cast<long long*>(out), reps, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
