// Kernel 2, fp32-parity (3-pass) form — fused K-slot decoder forward on tcgen05 / TMEM with HALF-TILE pipelining.
//
// Replaces Decoder.forward (reference models/model.py:106-161) and sin_emb (:261-277); same data path, weight image, precision
// and outputs as decoder_fwd.cu (see the design notes there).  What differs is how a tile moves between the tensor pipe and the
// CUDA cores.  TMEM (512 columns) caps the tiles in flight at four (D 64 | H hi 32 | H lo 32 columns each), and with whole-layer
// hand-offs a tile sits idle for the two wake-up latencies (issuer ~185, MMA drain + epilogue wake ~315 cycles) and the other
// party's whole service time of every layer: the tensor pipe was 46 % busy, the issue slots 44 %.  Here every layer is cut in
// two along BOTH GEMM dimensions:
//   * the epilogue of a layer runs as two halves E(0), E(1) of 32 accumulator columns; each produces one K-half (two K-steps)
//     of the next layer's A operand and has its own "operand ready" barrier a[0], a[1];
//   * the MMAs of a layer run as four groups G(k, n) = K-half k x N-half n (N = 32: 6 MMAs = 2 K-steps x 3 passes), and the two
//     N-halves have their own "accumulator ready" barriers d[0], d[1].
//   issuer:    wait a[0] -> [P(0)] G(0,0) | wait a[1] -> G(1,0), commit d[0] | [P(1)] G(0,1) G(1,1), commit d[1]
//   epilogue:  wait d[0] -> E(0): load D[0:32], bias preload, convert | wait d[1] | store H[0:16], arrive a[0]
//              -> E(1): load D[32:64], ..., store H[16:32], arrive a[1]
//   so G(0,0) of the next layer runs under E(1), and E(0) of the next layer runs under G(.,1) and its drain: what stays on a
//   tile's dependent chain per layer is  wake + G(1,0) + wake + E(0) + E(1)  instead of  wake + 12 MMAs + wake + E.
//   (E(0) may only overwrite H[0:16] once G(0,1) has read it: that is the wait on d[1] before its stores.)
//   N = 32 MMAs cost 19.8 instead of 34.6 / 2 cycles (+14 % tensor time); the tensor pipe had the headroom.
// The issuer is ONE code path for all four channels, seven stages and both phases (channel and stage are warp-uniform run-time
// values: 54 tcgen05.mma in the kernel instead of 672) - the per-(channel, stage) instantiation of decoder_fwd.cu made the kernel
// 156 KB of code, most of it issuer paths that run once per layer and evict each other from the instruction cache.
#include "decoder_fwd_common.cuh"

#ifndef UORF_X3_WARP_ARRIVE
#define UORF_X3_WARP_ARRIVE 0
#endif

namespace uorf {

constexpr int kX3Ch = 4;                       // channels = tiles in flight
constexpr int kX3EpiWarps = 4 * kX3Ch;         // one thread per point
constexpr int kX3Threads = (kX3EpiWarps + kX3Ch) * 32;
constexpr int kX3StripBytes = kTileM * 128;    // positional-encoding operand of one channel: 128 rows x [hi 32 | lo 32] 16-bit

struct X3Bars {
  uint64_t a[2];   // K-half k of the next layer's A operand written (every thread of the channel)
  uint64_t d[2];   // N-half n of the current layer's accumulator complete (tcgen05.commit)
};

__device__ __forceinline__ void x3_arrive(uint64_t* bar) {
#if UORF_X3_WARP_ARRIVE
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
#else
  mbar_arrive(bar);
#endif
}

// One half-layer of the epilogue: 32 accumulator columns -> [+ wx * xe] -> next layer's bias into the same columns -> ReLU +
// hi/lo split -> 16 + 16 operand columns.  `gate` (E(0) only): the barrier the stores of H have to wait for.
template <bool F16>
__device__ __forceinline__ void x3_half(uint32_t t_d, uint32_t t_hhi, uint32_t t_hlo, const float* wx, float xe, const float* next_bias,
                                        uint64_t* gate, uint32_t gate_parity, int* err_flag) {
  uint32_t r[32];
  tmem_ld32(t_d, r);
  tc_wait_ld();
  if (wx != nullptr) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(wx + 4 * q);
      r[4 * q + 0] = __float_as_uint(fmaf(w.x, xe, __uint_as_float(r[4 * q + 0])));
      r[4 * q + 1] = __float_as_uint(fmaf(w.y, xe, __uint_as_float(r[4 * q + 1])));
      r[4 * q + 2] = __float_as_uint(fmaf(w.z, xe, __uint_as_float(r[4 * q + 2])));
      r[4 * q + 3] = __float_as_uint(fmaf(w.w, xe, __uint_as_float(r[4 * q + 3])));
    }
  }
  if (next_bias != nullptr) preload_bias32(t_d, next_bias);
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (F16) {
      split2_relu<F16>(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]), hi[q], lo[q]);
    } else {   // bf16: keep round-to-nearest on the hi part (truncation would cost one of its 16 bits)
      split2<F16>(fmaxf(__uint_as_float(r[2 * q]), 0.f), fmaxf(__uint_as_float(r[2 * q + 1]), 0.f), hi[q], lo[q]);
    }
  }
  if (gate != nullptr) {
    mbar_wait(gate, gate_parity, err_flag, 310);
    tc_fence_after();
  }
  tmem_st16(t_hhi, hi);
  tmem_st16(t_hlo, lo);
}

template <bool F16>
__global__ void __launch_bounds__(kX3Threads, 1) decoder_fwd_x3_kernel(const DecFwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ __align__(8) X3Bars bars[kX3Ch];
  __shared__ uint32_t tmem_base_s;

  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int K = p.K;
  const int w_region = 2 * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  const float* s_params = reinterpret_cast<const float*>(smem + 2 * kPlaneBytes);
  unsigned char* s_strips = smem + ((w_region + 1023) & ~1023);   // [channel][128 rows x 128 B]

  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int c = 0; c < kX3Ch; ++c) {
      mbar_init(&bars[c].a[0], UORF_X3_WARP_ARRIVE ? 4 : kTileM);
      mbar_init(&bars[c].a[1], UORF_X3_WARP_ARRIVE ? 4 : kTileM);
      mbar_init(&bars[c].d[0], 1);
      mbar_init(&bars[c].d[1], 1);
    }
    fence_mbar_init();
  }
  if (warp == kX3EpiWarps) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  uint32_t n_stage = 0;   // stages completed by this thread's channel (parity of all four barriers = n_stage & 1)

  for (int phase = 0; phase < 2; ++phase) {  // 0: background MLP, 1: foreground MLP + composition
    const int nslots = phase == 0 ? 1 : K - 1;
    if (warp == kX3EpiWarps && lane == 0) {
      fence_proxy_async_smem();
      const unsigned char* src = p.image + (phase ? blob_fg_offset(3) : 0);
      const uint32_t bytes = (uint32_t)blob_bytes_one(3, nslots);
      mbar_arrive_expect_tx(&bar_w, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
        bulk_g2s(smem + off, src + off, n, &bar_w);
      }
    }
    mbar_wait(&bar_w, phase & 1, p.err_flag, 100 + phase);

    if (warp >= kX3EpiWarps) {
      // =========================== MMA issuer of channel c ===========================
      const int c = __shfl_sync(0xffffffffu, warp - kX3EpiWarps, 0);   // warp-uniform: operands stay on the uniform datapath
      X3Bars& cb = bars[c];
      const uint32_t w_smem = smem_u32(smem);
      const uint32_t td = tmem_base + 128 * c;           // accumulator D (64 columns)
      const uint32_t th = td + 64, tl = td + 96;         // operand H: hi / lo (32 columns each)
      const uint64_t pdesc = make_sdesc_sw128(smem_u32(s_strips) + c * kX3StripBytes);
      const uint32_t id64 = make_idesc_f32acc(128, 64, F16);
      const uint32_t id32 = make_idesc_f32acc(128, 32, F16);
      const uint32_t idh = make_idesc_f32acc(128, phase == 0 ? 16 : 32, F16);
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * ((long long)it * kX3Ch + c);
        if (tile >= p.num_tiles) break;
        for (int s = 0; s < nslots; ++s) {
#pragma unroll 1
          for (int st = 0; st < 7; ++st) {
            const uint32_t par = n_stage & 1;
            ++n_stage;
            // hidden-layer weight tile of MMA stage st: t1 t2 | t4 t5 t6 | head t7; encoding tiles t0 (st 0), t3 (st 3)
            const uint32_t wt = w_smem + (uint32_t)(st < 3 ? st : st + 1) * kTileBytes;
            const uint64_t b_hi = make_sdesc_sw128(wt), b_lo = make_sdesc_sw128(wt + kPlaneBytes);
            mbar_wait(&cb.a[0], par, p.err_flag, 200 + st);
            tc_fence_after();
            if (st == 0) {
              // layer before.0: encoding operand only (the slot-latent part is the preloaded bias); one N = 64 group
              mma_group_p_ss(td, pdesc, b_hi, b_lo, id64, 1u);
              mbar_wait(&cb.a[1], par, p.err_flag, 210);
              tc_commit_elect(&cb.d[0]);
              tc_commit_elect(&cb.d[1]);
            } else if (st == 6) {
              // heads: N = 32 (fg) / 16 (bg), no preloaded bias; K-halves as they arrive
              mma_group_ts<2, 3>(td, th, tl, b_hi, b_lo, idh, 0u);
              mbar_wait(&cb.a[1], par, p.err_flag, 216);
              tc_fence_after();
              mma_group_ts<2, 3>(td, th + 16, tl + 16, b_hi + 4, b_lo + 4, idh, 1u);
              tc_commit_elect(&cb.d[0]);
              tc_commit_elect(&cb.d[1]);
            } else {
              const uint64_t p_hi = make_sdesc_sw128(w_smem + 3 * kTileBytes), p_lo = make_sdesc_sw128(w_smem + 3 * kTileBytes + kPlaneBytes);
              if (st == 3) mma_group_p_ss(td, pdesc, p_hi, p_lo, id32, 1u);                      // P(0)
              mma_group_ts<2, 3>(td, th, tl, b_hi, b_lo, id32, 1u);                              // G(0,0)
              mbar_wait(&cb.a[1], par, p.err_flag, 210 + st);
              tc_fence_after();
              mma_group_ts<2, 3>(td, th + 16, tl + 16, b_hi + 4, b_lo + 4, id32, 1u);            // G(1,0)
              tc_commit_elect(&cb.d[0]);
              if (st == 3) mma_group_p_ss(td + 32, pdesc, p_hi + 256, p_lo + 256, id32, 1u);     // P(1): weight rows 32..63
              mma_group_ts<4, 3>(td + 32, th, tl, b_hi + 256, b_lo + 256, id32, 1u);             // G(0,1) G(1,1)
              tc_commit_elect(&cb.d[1]);
            }
            __syncwarp();
          }
        }
      }
    } else {
      // =========================== epilogue warps ===========================
      const int ch = warp >> 2;
      const int t = (warp & 3) * 32 + lane;    // point within the tile == TMEM lane
      const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t td = tmem_base + 128 * ch + lane_bits;
      const uint32_t th = td + 64, tl = td + 96;
      X3Bars& cb = bars[ch];
      // per-slot head results of the tile in flight: this channel's slice of a global buffer - each thread only ever reads back
      // what it wrote itself, so no fence is needed, and the 10-40 KB stay L2-resident
      float4* stash = p.stash_g + ((size_t)blockIdx.x * kX3Ch + ch) * K * kTileM;
      unsigned char* strip = s_strips + ch * kX3StripBytes;
      // rows of the fg transform: [r][0..2] rotation, [r][3] translation (0 unless fixed_locality)
      float T[12];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          T[r * 4 + cc] = p.fixed_locality ? __ldg(p.fg_transform + r * 4 + cc) : (cc < 3 ? __ldg(p.fg_transform + r * 3 + cc) : 0.f);
      }
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * ((long long)it * kX3Ch + ch);
        if (tile >= p.num_tiles) break;
        const long long pt = tile * kTileM + t;
        const bool valid = pt < p.P;
        // coordinates are fetched one slot ahead (and once per tile when all fg slots share them, fg_slot_stride == 0)
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (valid) {
          const float* src = (phase == 0 ? p.coor_bg : p.coor_fg) + pt * 3;
          nx = __ldg(src); ny = __ldg(src + 1); nz = __ldg(src + 2);
        }
        // When every fg slot sees the same coordinates (the reference's expand() view, no moved object) the encoding operand
        // written for the first slot is still in the strip, and the locality flag is the same: later slots skip the encoding.
        const bool shared_enc = phase == 1 && p.fg_slot_stride == 0 && p.fg_slot_offset == nullptr;
        float x_last = 0.f;
        bool outside = false;
        for (int s = 0; s < nslots; ++s) {
          const float* sb = s_params + kParamFloats + s * kSlotBiasFloats;
          // ---- stage 0: coordinates -> object frame -> locality flag -> positional encoding ----
          if (s == 0 || !shared_enc) {
            float x = nx, y = ny, z = nz;
            outside = false;
            if (valid && phase == 1 && p.fg_slot_stride != 0 && s + 1 < nslots) {
              const float* src = p.coor_fg + (long long)(s + 1) * p.fg_slot_stride + pt * 3;
              nx = __ldg(src); ny = __ldg(src + 1); nz = __ldg(src + 2);
            }
            if (phase == 1) {
              if (p.fg_slot_offset != nullptr) {   // moved object: same fp32 subtraction the reference applies to its coordinate copy
                x = __fsub_rn(x, __ldg(p.fg_slot_offset + 3 * s));
                y = __fsub_rn(y, __ldg(p.fg_slot_offset + 3 * s + 1));
                z = __fsub_rn(z, __ldg(p.fg_slot_offset + 3 * s + 2));
              }
              if (p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
              // unfused multiply/add in the order ((r0*x + r1*y) + r2*z) [+ t]: bit-equal to the reference's CPU matmul
              const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[3]);
              const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z)), T[7]);
              const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z)), T[11]);
              x = tx; y = ty; z = tz;
              if (!p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
            }
            float v[kPosencK];
            posenc48<false>(x, y, z, v);
            posenc_store_smem<F16>(v, strip, t);
            fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
            x_last = v[kPosencMma3];           // cos(16 z): added by FMA in the epilogues of layers before.0 / after.0
          }
          if (phase == 1 && p.outside && valid) p.outside[(long long)s * p.P + pt] = outside ? 1 : 0;
          preload_bias32(td, sb);
          preload_bias32(td + 32, sb + 32);
          tc_wait_st();
          tc_fence_before();
          x3_arrive(&cb.a[0]);
          x3_arrive(&cb.a[1]);

          // ---- stages 1..6: hidden layers, two halves each ----
#pragma unroll 1
          for (int st = 1; st <= 6; ++st) {
            const uint32_t par = n_stage & 1;
            ++n_stage;
            // bias of the layer MMA stage `st` computes (preloaded into the accumulator; the head, st == 6, keeps its bias in registers)
            const float* nb = st == 1 ? s_params + kPB1 : st == 2 ? s_params + kPB2 : st == 3 ? sb + 64
                            : st == 4 ? s_params + kPA1 : st == 5 ? s_params + kPA2 : nullptr;
            const float* wx = st == 1 ? s_params + kPW0x : st == 4 ? s_params + kPW3x : nullptr;
            mbar_wait(&cb.d[0], par, p.err_flag, 300 + st);
            tc_fence_after();
#pragma unroll 1
            for (int n = 0; n < 2; ++n) {
              x3_half<F16>(td + 32 * n, th + 16 * n, tl + 16 * n, wx ? wx + 32 * n : nullptr, x_last, nb ? nb + 32 * n : nullptr,
                           n == 0 ? &cb.d[1] : nullptr, par, p.err_flag);
              tc_wait_st();
              tc_fence_before();
              x3_arrive(&cb.a[n]);
            }
          }
          // ---- stage 7: heads ----
          {
            const uint32_t par = n_stage & 1;
            ++n_stage;
            mbar_wait(&cb.d[0], par, p.err_flag, 307);
            mbar_wait(&cb.d[1], par, p.err_flag, 308);
            tc_fence_after();
          }
          if (phase == 0) {
            uint32_t r[16];
            tmem_ld16(td, r);
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
            tmem_ld32(td, r);
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
#pragma unroll 4
          for (int k = 0; k < K; ++k) S += fmaxf(stash[k * kTileM + t].w, 0.f);
          const float denom = S + 1e-5f;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
          for (int k = 0; k < K; ++k) {
            const float4 rw = stash[k * kTileM + t];
            const float dens = fmaxf(rw.w, 0.f);
            const float m = dens / denom;
            float4 u;
            u.x = rgb_act<false>(rw.x);
            u.y = rgb_act<false>(rw.y);
            u.z = rgb_act<false>(rw.z);
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
          if (valid && p.raws != nullptr) p.raws[pt] = acc;
        }
      }
    }
    // all MMAs of this phase have been consumed; shared memory may be re-used for the next image
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == kX3EpiWarps) tmem_dealloc(tmem_base, kTmemCols);
}

size_t decoder_fwd_x3_smem_bytes(int K) {
  const int w_region = 2 * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  return 1024 + ((w_region + 1023) & ~1023) + (size_t)kX3Ch * kX3StripBytes;
}

int launch_decoder_fwd_x3(const DecFwdParams& p, bool f16, int num_sms, cudaStream_t stream) {
  const size_t smem = decoder_fwd_x3_smem_bytes(p.K);
  if (smem > 227 * 1024) return UORF_ERR_UNSUPPORTED;
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (grid < 1) return UORF_OK;
  auto launch = [&](auto kern) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UORF_ERR_DEVICE;
    kern<<<grid, kX3Threads, smem, stream>>>(p);
    return UORF_OK;
  };
  const int rc = f16 ? launch(decoder_fwd_x3_kernel<true>) : launch(decoder_fwd_x3_kernel<false>);
  if (rc != UORF_OK) return rc;
  return finish_launch(1);
}

}  // namespace uorf
