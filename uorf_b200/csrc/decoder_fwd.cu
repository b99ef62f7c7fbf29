// Kernel 2 — fused K-slot decoder forward on tcgen05 / TMEM.
//
// Replaces Decoder.forward (reference models/model.py:106-161) and sin_emb (:261-277):
//   per-slot object-frame transform + locality box test, sin/cos positional encoding, the background
//   MLP (7 linears) and the shared foreground MLP (9 linears, f_after_latent folded into f_color.0),
//   ReLU density / tanh colour, density-weighted composition over the K slots.
// No activation ever leaves the SM: the only HBM traffic is coords in, API-mandated outputs out.
//
// Structure (one persistent CTA per SM, 288 threads):
//   warps 0-3  "channel 0"  : 128 threads, thread t owns point t of the channel's current 128-point tile
//   warps 4-7  "channel 1"  : same, on another tile  (the two channels ping-pong on the tensor pipe)
//   warp  8    MMA issuer   : lane 0 issues tcgen05.mma; also owns TMEM alloc and the weight bulk copies
// Per layer:  A (activations, bf16 hi [+lo]) lives in TMEM, written by the channel's threads with
//   tcgen05.st; B (weights) is a 128B-swizzled K-major tile in shared memory; D (fp32) in TMEM.
//   channel --a_full mbarrier--> issuer --tcgen05.commit / d_full mbarrier--> channel (bias+ReLU+split).
// Precision: passes = 1: one 16-bit MMA pass; passes = 3: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (fp32 accumulate);
// operands bf16 (range-safe, ~16 bits after the split) or fp16 (saturating, ~22 bits after the split).
// The kernel makes two sweeps over its tiles: background MLP first (raw bg output parked in a 16 B/point
// scratch), then the foreground MLP for slots 1..K-1 followed by the cross-slot composition.  The split
// lets the bf16x3 weight image of ONE MLP (124 KB) stay resident in shared memory.
#include "common.cuh"
#include "decoder_layout.h"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kTileM = 128;
constexpr int kThreads = 288;
constexpr int kTmemCols = 512;
// TMEM column map of one channel (channel c at column c*256)
constexpr int kColD = 0;      // 64 fp32 accumulator columns
constexpr int kColHHi = 64;   // 32 columns: hidden activations, bf16 pairs (K = 64)
constexpr int kColHLo = 96;   // 32 columns: residuals
constexpr int kColPHi = 128;  // 24 columns: positional encoding (K = 48)
constexpr int kColPLo = 160;  // 24 columns
constexpr int kChanCols = 256;

struct DecFwdParams {
  const float* coor_bg;
  const float* coor_fg;
  long long fg_slot_stride;
  const float* noise;
  const unsigned char* image;
  float4* scratch_bg;
  float4* raws;
  float4* masked;
  float4* unmasked;
  float* masks;
  unsigned char* outside;
  const float* fg_transform;  // device: 3x3, or 4x4 when fixed_locality
  float locality_ratio;
  float dens_noise;
  long long P;
  int K;
  int fixed_locality;
  int locality;
  int num_tiles;
  int* err_flag;
};

// ---- channel-side helpers -------------------------------------------------------------------------

// positional encoding of one point -> TMEM (48 K-columns: 33 values + zero padding), reference order
// [x y z | sin(2^i x..z) cos(2^i x..z)]_{i<5}  (model.py:261-277).  Frequencies 2,4,8,16 by angle doubling.
template <int PASSES, bool F16>
__device__ __forceinline__ void posenc_to_tmem(float x, float y, float z, uint32_t t_phi, uint32_t t_plo) {
  float v[kPosencK];
  v[0] = x; v[1] = y; v[2] = z;
  float s[3], c[3];
  sincosf(x, &s[0], &c[0]);
  sincosf(y, &s[1], &c[1]);
  sincosf(z, &s[2], &c[2]);
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      v[3 + 6 * i + j] = s[j];
      v[3 + 6 * i + 3 + j] = c[j];
    }
    if (i < 4) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float s2 = 2.f * s[j] * c[j];
        const float c2 = 1.f - 2.f * s[j] * s[j];
        s[j] = s2;
        c[j] = c2;
      }
    }
  }
#pragma unroll
  for (int i = kPosenc; i < kPosencK; ++i) v[i] = 0.f;
  uint32_t hi[24], lo[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    if (PASSES == 3) split2<F16>(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    else hi[i] = pack2<F16>(v[2 * i], v[2 * i + 1]);
  }
  tmem_st16(t_phi, hi);
  tmem_st8(t_phi + 16, hi + 16);
  if (PASSES == 3) {
    tmem_st16(t_plo, lo);
    tmem_st8(t_plo + 16, lo + 16);
  }
  tc_wait_st();
}

// hidden-layer epilogue: D (64 fp32 cols) + bias -> ReLU -> bf16 hi[/lo] -> H operand columns
template <int PASSES, bool F16>
__device__ __forceinline__ void hidden_epilogue(uint32_t t_d, uint32_t t_hhi, uint32_t t_hlo, const float* bias) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    tmem_ld32(t_d + half * 32, r);
    tc_wait_ld();
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = *reinterpret_cast<const float4*>(bias + half * 32 + 4 * q);
      const float v0 = fmaxf(__uint_as_float(r[4 * q + 0]) + b.x, 0.f);
      const float v1 = fmaxf(__uint_as_float(r[4 * q + 1]) + b.y, 0.f);
      const float v2 = fmaxf(__uint_as_float(r[4 * q + 2]) + b.z, 0.f);
      const float v3 = fmaxf(__uint_as_float(r[4 * q + 3]) + b.w, 0.f);
      if (PASSES == 3) {
        split2<F16>(v0, v1, hi[2 * q], lo[2 * q]);
        split2<F16>(v2, v3, hi[2 * q + 1], lo[2 * q + 1]);
      } else {
        hi[2 * q] = pack2<F16>(v0, v1);
        hi[2 * q + 1] = pack2<F16>(v2, v3);
      }
    }
    tmem_st16(t_hhi + half * 16, hi);
    if (PASSES == 3) tmem_st16(t_hlo + half * 16, lo);
  }
  tc_wait_st();
}

// ---- issuer-side helper ---------------------------------------------------------------------------
// D (+)= A[tmem] . W[tile]^T over `ksteps` K-steps of 16, in 1 or 3 bf16 passes
template <int PASSES>
__device__ __forceinline__ void issue_gemm(uint32_t t_d, uint32_t t_ahi, uint32_t t_alo, uint32_t w_smem,
                                           int ksteps, uint32_t idesc, bool& first) {
  const uint64_t dhi = make_sdesc_sw128(w_smem);
  const uint64_t dlo = make_sdesc_sw128(w_smem + kPlaneBytes);
  for (int j = 0; j < ksteps; ++j) {
    mma_ts(t_d, t_ahi + 8 * j, dhi + 2 * j, idesc, first ? 0u : 1u);
    first = false;
  }
  if (PASSES == 3) {
    for (int j = 0; j < ksteps; ++j) mma_ts(t_d, t_alo + 8 * j, dhi + 2 * j, idesc, 1u);
    for (int j = 0; j < ksteps; ++j) mma_ts(t_d, t_ahi + 8 * j, dlo + 2 * j, idesc, 1u);
  }
}

template <int PASSES, bool F16>
__global__ void __launch_bounds__(kThreads, 1) decoder_fwd_kernel(const DecFwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ __align__(8) uint64_t bar_a[2];
  __shared__ __align__(8) uint64_t bar_d[2];
  __shared__ uint32_t tmem_base_s;

  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int K = p.K;
  constexpr int NPL = PASSES == 3 ? 2 : 1;
  const int w_region = NPL * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  const float* s_params = reinterpret_cast<const float*>(smem + NPL * kPlaneBytes);
  const float* s_slotb = s_params + kParamFloats;
  float4* s_stash = reinterpret_cast<float4*>(smem + ((w_region + 127) & ~127));  // [2][K][128]

  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_a[0], kTileM);
    mbar_init(&bar_a[1], kTileM);
    mbar_init(&bar_d[0], 1);
    mbar_init(&bar_d[1], 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // rows of the fg transform: [r][0..2] rotation, [r][3] translation (0 unless fixed_locality)
  float T[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      T[r * 4 + c] = p.fixed_locality ? __ldg(p.fg_transform + r * 4 + c) : (c < 3 ? __ldg(p.fg_transform + r * 3 + c) : 0.f);
  }

  // running completion counts of the channel barriers (parity = count & 1); every role tracks its own copy
  uint32_t n_a[2] = {0, 0}, n_d[2] = {0, 0};

  for (int phase = 0; phase < 2; ++phase) {  // 0: background MLP, 1: foreground MLP + composition
    const int nslots = phase == 0 ? 1 : K - 1;
    // ---- load this phase's weight image ----------------------------------------------------------
    if (warp == 8 && lane == 0) {
      fence_proxy_async_smem();
      const unsigned char* src = p.image + (phase ? blob_fg_offset(PASSES) : 0);
      const uint32_t bytes = (uint32_t)blob_bytes_one(PASSES, nslots);
      mbar_arrive_expect_tx(&bar_w, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
        bulk_g2s(smem + off, src + off, n, &bar_w);
      }
    }
    mbar_wait(&bar_w, phase & 1, p.err_flag, 100 + phase);

    const uint32_t idesc64 = make_idesc_f32acc(128, 64, F16);
    const uint32_t idesc_head = make_idesc_f32acc(128, phase == 0 ? 16 : 32, F16);

    if (warp == 8) {
      // =========================== MMA issuer ===========================
      const uint32_t w_smem = smem_u32(smem);
      for (int it = 0;; ++it) {
        const long long tile0 = (long long)blockIdx.x + (long long)gridDim.x * (2 * it);
        const long long tile1 = tile0 + gridDim.x;
        if (tile0 >= p.num_tiles) break;
        const bool has1 = tile1 < p.num_tiles;
        for (int s = 0; s < nslots; ++s) {
          for (int st = 0; st < 7; ++st) {
            for (int ch = 0; ch < 2; ++ch) {
              if (ch == 1 && !has1) continue;
              mbar_wait(&bar_a[ch], n_a[ch] & 1, p.err_flag, 200 + st);
              ++n_a[ch];
              tc_fence_after();
              if (lane == 0) {
                const uint32_t tb = tmem_base + ch * kChanCols;
                bool first = true;
                switch (st) {
                  case 0: issue_gemm<PASSES>(tb + kColD, tb + kColPHi, tb + kColPLo, w_smem + 0 * kTileBytes, 3, idesc64, first); break;
                  case 1: issue_gemm<PASSES>(tb + kColD, tb + kColHHi, tb + kColHLo, w_smem + 1 * kTileBytes, 4, idesc64, first); break;
                  case 2: issue_gemm<PASSES>(tb + kColD, tb + kColHHi, tb + kColHLo, w_smem + 2 * kTileBytes, 4, idesc64, first); break;
                  case 3:
                    issue_gemm<PASSES>(tb + kColD, tb + kColPHi, tb + kColPLo, w_smem + 3 * kTileBytes, 3, idesc64, first);
                    issue_gemm<PASSES>(tb + kColD, tb + kColHHi, tb + kColHLo, w_smem + 4 * kTileBytes, 4, idesc64, first);
                    break;
                  case 4: issue_gemm<PASSES>(tb + kColD, tb + kColHHi, tb + kColHLo, w_smem + 5 * kTileBytes, 4, idesc64, first); break;
                  case 5: issue_gemm<PASSES>(tb + kColD, tb + kColHHi, tb + kColHLo, w_smem + 6 * kTileBytes, 4, idesc64, first); break;
                  default: issue_gemm<PASSES>(tb + kColD, tb + kColHHi, tb + kColHLo, w_smem + 7 * kTileBytes, 4, idesc_head, first); break;
                }
                tc_commit(&bar_d[ch]);
              }
              __syncwarp();
            }
          }
        }
      }
    } else {
      // =========================== channel (one point per thread) ===========================
      const int ch = warp >> 2;
      const int t = threadIdx.x & 127;
      const uint32_t tb = tmem_base + ch * kChanCols + ((uint32_t)((warp & 3) * 32) << 16);
      float4* stash = s_stash + (size_t)ch * K * kTileM;
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * (2 * it + ch);
        if (tile >= p.num_tiles) break;
        const long long pt = tile * kTileM + t;
        const bool valid = pt < p.P;
        for (int s = 0; s < nslots; ++s) {
          // ---- stage 0: coordinates -> object frame -> locality flag -> positional encoding ----
          float x = 0.f, y = 0.f, z = 0.f;
          bool outside = false;
          if (valid) {
            const float* src = phase == 0 ? p.coor_bg + pt * 3 : p.coor_fg + (long long)s * p.fg_slot_stride + pt * 3;
            x = __ldg(src); y = __ldg(src + 1); z = __ldg(src + 2);
          }
          if (phase == 1) {
            if (p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
            // unfused multiply/add in the order ((r0*x + r1*y) + r2*z) [+ t]: bit-equal to the reference's
            // CPU matmul
            const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[3]);
            const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z)), T[7]);
            const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z)), T[11]);
            x = tx; y = ty; z = tz;
            if (!p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
            if (p.outside && valid) p.outside[(long long)s * p.P + pt] = outside ? 1 : 0;
          }
          posenc_to_tmem<PASSES, F16>(x, y, z, tb + kColPHi, tb + kColPLo);
          tc_fence_before();
          mbar_arrive(&bar_a[ch]);

          // ---- stages 1..6: hidden layers ----
          const float* sb = s_slotb + s * kSlotBiasFloats;
#pragma unroll 1
          for (int st = 1; st <= 6; ++st) {
            const float* bias = st == 1 ? sb : st == 2 ? s_params + kPB1 : st == 3 ? s_params + kPB2
                              : st == 4 ? sb + 64 : st == 5 ? s_params + kPA1 : s_params + kPA2;
            mbar_wait(&bar_d[ch], n_d[ch] & 1, p.err_flag, 300 + st);
            ++n_d[ch];
            tc_fence_after();
            hidden_epilogue<PASSES, F16>(tb + kColD, tb + kColHHi, tb + kColHLo, bias);
            tc_fence_before();
            mbar_arrive(&bar_a[ch]);
          }
          // ---- stage 7: heads ----
          mbar_wait(&bar_d[ch], n_d[ch] & 1, p.err_flag, 307);
          ++n_d[ch];
          tc_fence_after();
          if (phase == 0) {
            uint32_t r[16];
            tmem_ld16(tb + kColD, r);
            tc_wait_ld();
            const float* hb = s_params + kPHeadB;
            float4 o;
            o.x = __uint_as_float(r[0]) + hb[0];
            o.y = __uint_as_float(r[1]) + hb[1];
            o.z = __uint_as_float(r[2]) + hb[2];
            o.w = __uint_as_float(r[3]) + hb[3];
            if (valid) p.scratch_bg[pt] = o;
          } else {
            uint32_t r[32];
            tmem_ld32(tb + kColD, r);
            tc_wait_ld();
            const float* hb = s_params + kPHeadB;
            const float* wc2 = s_params + kPWc2;
            float c0 = s_params[kPBc2 + 0], c1 = s_params[kPBc2 + 1], c2 = s_params[kPBc2 + 2];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float h = fmaxf(__uint_as_float(r[i]) + hb[i], 0.f);
              c0 = fmaf(wc2[i], h, c0);
              c1 = fmaf(wc2[16 + i], h, c1);
              c2 = fmaf(wc2[32 + i], h, c2);
            }
            float sig = __uint_as_float(r[kHeadShapeRow]) + hb[kHeadShapeRow];
            if (p.locality && outside) sig = 0.f;  // model.py:147-148
            stash[(s + 1) * kTileM + t] = make_float4(c0, c1, c2, sig);
          }
        }
        if (phase == 1) {
          // ---- cross-slot composition (model.py:151-159) ----
          if (valid) stash[t] = p.scratch_bg[pt];
          float S = 0.f;
          for (int k = 0; k < K; ++k) S += fmaxf(stash[k * kTileM + t].w, 0.f);
          const float denom = S + 1e-5f;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int k = 0; k < K; ++k) {
            const float4 rw = stash[k * kTileM + t];
            const float dens = fmaxf(rw.w, 0.f);
            const float m = dens / denom;
            float4 u;
            u.x = (tanhf(rw.x) + 1.f) * 0.5f;
            u.y = (tanhf(rw.y) + 1.f) * 0.5f;
            u.z = (tanhf(rw.z) + 1.f) * 0.5f;
            u.w = dens;
            if (p.noise != nullptr && valid) u.w = dens + p.dens_noise * __ldg(p.noise + (long long)k * p.P + pt);
            const float4 mk = make_float4(u.x * m, u.y * m, u.z * m, u.w * m);
            acc.x += mk.x; acc.y += mk.y; acc.z += mk.z; acc.w += mk.w;
            if (valid) {
              if (p.unmasked) p.unmasked[(long long)k * p.P + pt] = u;
              if (p.masked) p.masked[(long long)k * p.P + pt] = mk;
              if (p.masks) p.masks[(long long)k * p.P + pt] = m;
            }
          }
          if (valid && p.raws) p.raws[pt] = acc;
        }
      }
    }
    // all MMAs of this phase have been consumed; shared memory may be re-used for the next image
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 8) tmem_dealloc(tmem_base, kTmemCols);
}

static size_t decoder_smem_bytes(int passes, int K) {
  const int npl = passes == 3 ? 2 : 1;
  const int w_region = npl * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  return 1024 + ((w_region + 127) & ~127) + (size_t)2 * K * kTileM * sizeof(float4);
}

int launch_decoder_fwd(const uorf_decoder_args* a, int num_sms, int* err_flag, cudaStream_t stream) {
  DecFwdParams p;
  p.coor_bg = a->coor_bg;
  p.coor_fg = a->coor_fg;
  p.fg_slot_stride = a->fg_slot_stride;
  p.noise = a->dens_noise != 0.f ? a->noise : nullptr;
  p.image = static_cast<const unsigned char*>(a->image);
  p.scratch_bg = reinterpret_cast<float4*>(a->workspace);
  p.raws = reinterpret_cast<float4*>(a->raws);
  p.masked = reinterpret_cast<float4*>(a->masked_raws);
  p.unmasked = reinterpret_cast<float4*>(a->unmasked_raws);
  p.masks = a->masks;
  p.outside = a->outside;
  p.fg_transform = a->fg_transform;
  p.locality_ratio = a->locality_ratio;
  p.dens_noise = a->dens_noise;
  p.P = a->P;
  p.K = a->K;
  p.fixed_locality = a->fixed_locality;
  p.locality = a->locality;
  p.num_tiles = (int)((a->P + kTileM - 1) / kTileM);
  p.err_flag = err_flag;
  const size_t smem = decoder_smem_bytes(a->precision >= 3 ? 3 : 1, a->K);
  if (smem > 227 * 1024) return UORF_ERR_UNSUPPORTED;
  int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (grid < 1) return UORF_OK;
  const int passes = a->precision >= 3 ? 3 : 1;
  const bool f16 = (a->precision % 2) == 0;
  auto launch = [&](auto kern) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UORF_ERR_DEVICE;
    kern<<<grid, kThreads, smem, stream>>>(p);
    return UORF_OK;
  };
  int rc;
  if (passes == 3) rc = f16 ? launch(decoder_fwd_kernel<3, true>) : launch(decoder_fwd_kernel<3, false>);
  else rc = f16 ? launch(decoder_fwd_kernel<1, true>) : launch(decoder_fwd_kernel<1, false>);
  if (rc != UORF_OK) return rc;
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

}  // namespace uorf
