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
// UORF_X3_RELAY=1 (experiment, off): "accumulator ready" reaches the epilogue warps through hardware named barriers - the issuer
// warp is the only poller of the d[] mbarriers (tcgen05.commit can only signal an mbarrier) and relays each completion with
// bar.arrive, the 16 epilogue warps block in bar.sync and spend no issue slots on probing.  Measured: 2.93 ms against 2.77 ms with
// every waiting warp probing its mbarrier - the relay's extra hop on the dependent chain costs more than the probe loops.
// K-split of the operand hand-off (a[0] after E(0), a[1] after E(1)); 0: one hand-off per layer after E(1) (N-split only): 3.10 ms
#ifndef UORF_X3_KSPLIT
#define UORF_X3_KSPLIT 1
#endif
// next layer's bias fetched before the accumulator / gate waits (1) or under the TMEM load (0, default: measured 2 % faster -
// the 32 extra live registers cost more than the shared-memory latency they hide)
#ifndef UORF_X3_PREFETCH
#define UORF_X3_PREFETCH 0
#endif
// issuer warps hand their registers to the epilogue warps (setmaxnreg): 0 -> 3.10 ms (the epilogue re-derives addresses per stage)
#ifndef UORF_X3_SETMAXNREG
#define UORF_X3_SETMAXNREG 1
#endif
// cross-slot composition (experiment, off: no measurable difference): colour activation (tanh + 1) / 2 as 1 / (1 + 2^(-2 x log2 e))
// with ex2.approx and a correctly rounded reciprocal (abs error < 3e-7), masks as density * rcp(sum) instead of K divisions
#ifndef UORF_X3_FAST_ACT
#define UORF_X3_FAST_ACT 0
#endif
// E(1) stores its operand words in two batches of eight as they are converted (1) or all sixteen at the end (0)
#ifndef UORF_X3_EARLY_ST
#define UORF_X3_EARLY_ST 0
#endif
// E(0) also reads the second accumulator half (after its gate) and preloads it before the first hand-off: the issuer runs G(0,1)
// and P(1) under E(1), only G(1,0) and G(1,1) remain behind the second hand-off
#ifndef UORF_X3_EARLY_D1
#define UORF_X3_EARLY_D1 0
#endif
#ifndef UORF_X3_DISCARD_BG
#define UORF_X3_DISCARD_BG 0
#endif
#ifndef UORF_X3_RELAY
#define UORF_X3_RELAY 0
#endif

// -DUORF_TIMELINE_X3 (tuning variant only, scripts/timeline_x3.py): CTA 0 / channel 0 records clock64 stamps of the hand-off
// points of every stage of the fg sweep into a global buffer read back through uorf_debug_timeline_x3()
#ifdef UORF_TIMELINE_X3
__device__ long long g_tlx_epi[16384];
__device__ long long g_tlx_iss[16384];
#define TLX_EPI(id) do { if (tl_on && tl_n < 16384) g_tlx_epi[tl_n++] = (clock64() << 8) | (id); } while (0)
#define TLX_ISS(id) do { if (tl_on && tl_n < 16384) g_tlx_iss[tl_n++] = (clock64() << 8) | (id); } while (0)
#else
#define TLX_EPI(id)
#define TLX_ISS(id)
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

__device__ __forceinline__ void x3_wait(uint32_t bar_addr, uint32_t parity, int* err_flag, int site) {
  mbar_wait_lean(bar_addr, parity, err_flag, site);
}
__device__ __forceinline__ void x3_arrive_addr(uint32_t bar_addr) { mbar_arrive_addr(bar_addr); }
template <typename T>
__device__ __forceinline__ T* opaque_ptr(T* q) {
  asm volatile("" : "+l"(q));
  return q;
}
// L2 residency of the background sweep's 16 B / point scratch: it is written in phase 0 and read back in phase 1, with ~700 MB of
// output stores in between.  The scratch is stored with an evict_last policy and the outputs as streaming (evict-first) stores, so
// the read-back hits L2 instead of making a 134 MB DRAM round trip.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_global_hint(float4* ptr, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
// After its read-back a scratch line is dead: dropping it from L2 (no write-back) also saves the 67 MB the dirty lines would cost
// when they are finally evicted.  128-byte aligned line wholly inside the scratch.
__device__ __forceinline__ void l2_discard_line(const void* ptr) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(ptr) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// epilogue side: wait until N-half n of the channel's accumulator is complete
// bars_addr: shared address of the channel's X3Bars; bar_id: first named barrier of the channel
__device__ __forceinline__ void x3_wait_d(uint32_t bars_addr, int bar_id, int n, uint32_t parity, int* err_flag, int site) {
#if UORF_X3_RELAY
  named_bar_sync(bar_id + n, kTileM + 32);
#else
  x3_wait(bars_addr + 16 + 8 * n, parity, err_flag, site);
#endif
  tc_fence_after();
}
// issuer side: relay the completions of d[0], d[1] (in this order) to the channel's named barriers
__device__ __forceinline__ void x3_relay_d(uint32_t bars_addr, int bar_id, uint32_t parity, int* err_flag) {
#if UORF_X3_RELAY
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    x3_wait(bars_addr + 16 + 8 * n, parity, err_flag, 220 + n);
    tc_fence_after();
    tc_fence_before();
    named_bar_arrive(bar_id + n, kTileM + 32);
  }
#endif
}

__device__ __forceinline__ void x3_arrive(uint32_t bar_addr) {
#if UORF_X3_WARP_ARRIVE
  __syncwarp();
  if ((threadIdx.x & 31) == 0) x3_arrive_addr(bar_addr);
#else
  x3_arrive_addr(bar_addr);
#endif
}

// MMA groups of the issuer for PASSES = 3 (hi.hi + lo.hi + hi.lo) and PASSES = 1 (hi.hi only: the one-pass fast modes)
// NK K-steps of the TMEM operand H (hi at a_hi, lo at a_lo) against the weight tile pair (b_hi, b_lo)
template <int PASSES, int NK>
__device__ __forceinline__ void x3_group_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, uint32_t acc) {
  if constexpr (PASSES == 3) {
    mma_group_ts<NK, 3>(d, a_hi, a_lo, b_hi, b_lo, idesc, acc);
  } else {
    static_assert(NK == 2 || NK == 4, "unsupported group");
    if constexpr (NK == 4) {
      mma_group_ts<4, 1>(d, a_hi, a_lo, b_hi, b_lo, idesc, acc);
    } else {
      asm volatile(
          "{\n\t.reg .pred e, p, t;\n\t.reg .b32 a1;\n\t.reg .b64 h1;\n\t"
          "elect.sync _|e, 0xffffffff;\n\t"
          "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, %4, %4;\n\t"
          "add.u32 a1, %1, 8;\n\tadd.u64 h1, %2, 2;\n\t"
          "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
          "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], h1, %3, t;\n\t"
          "}" ::"r"(d), "r"(a_hi), "l"(b_hi), "r"(idesc), "r"(acc) : "memory");
    }
  }
}
// the two encoding K-steps from the shared-memory strip (SS)
template <int PASSES>
__device__ __forceinline__ void x3_group_p(uint32_t d, uint64_t pdesc, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, uint32_t acc) {
  if constexpr (PASSES == 3) mma_group_p_ss(d, pdesc, b_hi, b_lo, idesc, acc);
  else mma_group_ss<2, 2, 2>(d, pdesc, b_hi, idesc, acc);
}

// One half-layer of the epilogue, in two steps.
// x3_convert: 32 accumulator columns -> registers; the next layer's bias (fetched by the caller BEFORE it waited for the accumulator,
//   so no shared-memory latency sits behind the wait) is stored over the columns just read; [+ wx * xe]; ReLU + hi/lo split.
// The caller then stores the 16 + 16 operand words (E(0): after the gate on d[1]) and hands them to the issuer.
__device__ __forceinline__ void x3_load_bias(uint32_t addr, float4 (&b)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) b[q] = lds128(addr + 16 * q);
}
// [+ wx * xe]; ReLU + hi/lo split of 32 accumulator values already in registers
template <int PASSES, bool F16>
__device__ __forceinline__ void x3_split32(uint32_t (&r)[32], bool has_wx, uint32_t wx, float xe, uint32_t (&hi)[16], uint32_t (&lo)[16]) {
  if (has_wx) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 w = lds128(wx + 16 * q);
      r[4 * q + 0] = __float_as_uint(fmaf(w.x, xe, __uint_as_float(r[4 * q + 0])));
      r[4 * q + 1] = __float_as_uint(fmaf(w.y, xe, __uint_as_float(r[4 * q + 1])));
      r[4 * q + 2] = __float_as_uint(fmaf(w.z, xe, __uint_as_float(r[4 * q + 2])));
      r[4 * q + 3] = __float_as_uint(fmaf(w.w, xe, __uint_as_float(r[4 * q + 3])));
    }
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (PASSES == 1) {
      hi[q] = pack2_relu<F16>(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
      lo[q] = hi[q];
    } else if (F16) {
      split2_relu<F16>(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]), hi[q], lo[q]);
    } else {
      split2<F16>(fmaxf(__uint_as_float(r[2 * q]), 0.f), fmaxf(__uint_as_float(r[2 * q + 1]), 0.f), hi[q], lo[q]);
    }
  }
}
// 32 accumulator columns <- bias (32 floats at shared address `addr`)
__device__ __forceinline__ void x3_preload32(uint32_t t_d, uint32_t addr) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t w[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = lds128(addr + 64 * h + 16 * q);
      w[4 * q + 0] = __float_as_uint(b.x); w[4 * q + 1] = __float_as_uint(b.y);
      w[4 * q + 2] = __float_as_uint(b.z); w[4 * q + 3] = __float_as_uint(b.w);
    }
    tmem_st16(t_d + 16 * h, w);
  }
}
template <int PASSES, bool F16, bool STORE = false>
__device__ __forceinline__ void x3_convert(uint32_t t_d, bool has_wx, uint32_t wx, float xe, bool has_bias, const float4 (&b)[8],
                                           uint32_t (&hi)[16], uint32_t (&lo)[16], uint32_t t_hhi = 0, uint32_t t_hlo = 0) {
  uint32_t r[32];
  tmem_ld32(t_d, r);
  tc_wait_ld();
  if (has_bias) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t w[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        w[4 * q + 0] = __float_as_uint(b[4 * h + q].x); w[4 * q + 1] = __float_as_uint(b[4 * h + q].y);
        w[4 * q + 2] = __float_as_uint(b[4 * h + q].z); w[4 * q + 3] = __float_as_uint(b[4 * h + q].w);
      }
      tmem_st16(t_d + 16 * h, w);
    }
  }
  if (has_wx) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 w = lds128(wx + 16 * q);
      r[4 * q + 0] = __float_as_uint(fmaf(w.x, xe, __uint_as_float(r[4 * q + 0])));
      r[4 * q + 1] = __float_as_uint(fmaf(w.y, xe, __uint_as_float(r[4 * q + 1])));
      r[4 * q + 2] = __float_as_uint(fmaf(w.z, xe, __uint_as_float(r[4 * q + 2])));
      r[4 * q + 3] = __float_as_uint(fmaf(w.w, xe, __uint_as_float(r[4 * q + 3])));
    }
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (PASSES == 1) {
      hi[q] = pack2_relu<F16>(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
      lo[q] = hi[q];      // (not stored; keeps the ordering anchor of the caller meaningful)
    } else if (F16) {
      split2_relu<F16>(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]), hi[q], lo[q]);
    } else {   // bf16: keep round-to-nearest on the hi part (truncation would cost one of its 16 bits)
      split2<F16>(fmaxf(__uint_as_float(r[2 * q]), 0.f), fmaxf(__uint_as_float(r[2 * q + 1]), 0.f), hi[q], lo[q]);
    }
    if (STORE && (q & 7) == 7) {   // E(1): operand words stored as soon as eight are ready (nothing gates them)
      tmem_st8(t_hhi + (q - 7), &hi[q - 7]);
      if (PASSES == 3) tmem_st8(t_hlo + (q - 7), &lo[q - 7]);
    }
  }
}

template <int PASSES, bool F16>
__global__ void __launch_bounds__(kX3Threads, 1) decoder_fwd_x3_kernel(const DecFwdParams p) {
  constexpr int NPL = PASSES == 3 ? 2 : 1;     // weight planes (hi [+ lo])
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ __align__(8) X3Bars bars[kX3Ch];
  __shared__ uint32_t tmem_base_s;

  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int K = p.K;
  const int w_region = NPL * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  const float* s_params = reinterpret_cast<const float*>(smem + NPL * kPlaneBytes);
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

  // The two roles never share code after this point (each has its own phase loop), so the issuer warps - one warpgroup - can hand
  // most of their registers to the epilogue warps: 128 x (96 - 40) = 512 x 14 -> 104 registers per epilogue thread.
  if (warp >= kX3EpiWarps) {
#if UORF_X3_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
#endif
    for (int phase = 0; phase < 2; ++phase) {  // 0: background MLP, 1: foreground MLP + composition
      const int nslots = phase == 0 ? 1 : K - 1;
      if (warp == kX3EpiWarps && lane == 0) {
        fence_proxy_async_smem();
        const unsigned char* src = p.image + (phase ? blob_fg_offset(PASSES) : 0);
        const uint32_t bytes = (uint32_t)blob_bytes_one(PASSES, nslots);
        mbar_arrive_expect_tx(&bar_w, bytes);
        for (uint32_t off = 0; off < bytes; off += 32768u) {
          const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
          bulk_g2s(smem + off, src + off, n, &bar_w);
        }
      }
      x3_wait(smem_u32(&bar_w), phase & 1, p.err_flag, 100 + phase);
      // =========================== MMA issuer of channel c ===========================
      const int c = __shfl_sync(0xffffffffu, warp - kX3EpiWarps, 0);   // warp-uniform: operands stay on the uniform datapath
      X3Bars& cb = bars[c];
      const uint32_t cba = smem_u32(&bars[c]);
      const uint32_t w_smem = smem_u32(smem);
      const uint32_t td = tmem_base + 128 * c;           // accumulator D (64 columns)
      const uint32_t th = td + 64, tl = td + 96;         // operand H: hi / lo (32 columns each)
      const uint64_t pdesc = make_sdesc_sw128(smem_u32(s_strips) + c * kX3StripBytes);
      const uint32_t id64 = make_idesc_f32acc(128, 64, F16);
      const uint32_t id32 = make_idesc_f32acc(128, 32, F16);
      const uint32_t idh = make_idesc_f32acc(128, phase == 0 ? 16 : 32, F16);
#ifdef UORF_TIMELINE_X3
      const bool tl_on = blockIdx.x == 0 && c == 0 && lane == 0 && phase == 1;
      int tl_n = 0;
#endif
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
            const uint64_t b_hi = make_sdesc_sw128(wt), b_lo = make_sdesc_sw128(wt + (NPL - 1) * kPlaneBytes);
            x3_wait(cba, par, p.err_flag, 200 + st);
            tc_fence_after();
            TLX_ISS(0x60 + st);
            if (st == 0) {
              // layer before.0: encoding operand only (the slot-latent part is the preloaded bias); one N = 64 group
              x3_group_p<PASSES>(td, pdesc, b_hi, b_lo, id64, 1u);
              if (UORF_X3_KSPLIT) x3_wait(cba + 8, par, p.err_flag, 210);
              tc_commit_elect(&cb.d[0]);
              tc_commit_elect(&cb.d[1]);
            } else if (st == 6) {
              // heads: N = 32 (fg) / 16 (bg), no preloaded bias; K-halves as they arrive
              x3_group_ts<PASSES, 2>(td, th, tl, b_hi, b_lo, idh, 0u);
              if (UORF_X3_KSPLIT) {
                x3_wait(cba + 8, par, p.err_flag, 216);
                tc_fence_after();
              }
              x3_group_ts<PASSES, 2>(td, th + 16, tl + 16, b_hi + 4, b_lo + 4, idh, 1u);
              tc_commit_elect(&cb.d[0]);
              tc_commit_elect(&cb.d[1]);
            } else {
              const uint64_t p_hi = make_sdesc_sw128(w_smem + 3 * kTileBytes), p_lo = make_sdesc_sw128(w_smem + 3 * kTileBytes + (NPL - 1) * kPlaneBytes);
              if (st == 3) x3_group_p<PASSES>(td, pdesc, p_hi, p_lo, id32, 1u);                      // P(0)
              x3_group_ts<PASSES, 2>(td, th, tl, b_hi, b_lo, id32, 1u);                              // G(0,0)
#if UORF_X3_EARLY_D1
              if (st == 3) x3_group_p<PASSES>(td + 32, pdesc, p_hi + 256, p_lo + 256, id32, 1u);     // P(1): weight rows 32..63
              x3_group_ts<PASSES, 2>(td + 32, th, tl, b_hi + 256, b_lo + 256, id32, 1u);             // G(0,1)
#endif
              if (UORF_X3_KSPLIT) {
                x3_wait(cba + 8, par, p.err_flag, 210 + st);
                tc_fence_after();
              }
              TLX_ISS(0x70 + st);
              x3_group_ts<PASSES, 2>(td, th + 16, tl + 16, b_hi + 4, b_lo + 4, id32, 1u);            // G(1,0)
              tc_commit_elect(&cb.d[0]);
              TLX_ISS(0x80 + st);
#if UORF_X3_EARLY_D1
              x3_group_ts<PASSES, 2>(td + 32, th + 16, tl + 16, b_hi + 260, b_lo + 260, id32, 1u);   // G(1,1)
#else
              if (st == 3) x3_group_p<PASSES>(td + 32, pdesc, p_hi + 256, p_lo + 256, id32, 1u);     // P(1): weight rows 32..63
              x3_group_ts<PASSES, 4>(td + 32, th, tl, b_hi + 256, b_lo + 256, id32, 1u);             // G(0,1) G(1,1)
#endif
              tc_commit_elect(&cb.d[1]);
            }
            TLX_ISS(0x90 + st);
            __syncwarp();
            x3_relay_d(cba, 1 + 2 * c, par, p.err_flag);
          }
        }
      }
      // all MMAs of this phase have been consumed; shared memory may be re-used for the next image
      tc_fence_before();
      named_bar_sync(0, kX3Threads);
      tc_fence_after();
    }
    if (warp == kX3EpiWarps) tmem_dealloc(tmem_base, kTmemCols);
  } else {
#if UORF_X3_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
#endif
    for (int phase = 0; phase < 2; ++phase) {
      const int nslots = phase == 0 ? 1 : K - 1;
      x3_wait(smem_u32(&bar_w), phase & 1, p.err_flag, 102 + phase);
      // =========================== epilogue warps ===========================
      const int ch = warp >> 2;
      const int t = (warp & 3) * 32 + lane;    // point within the tile == TMEM lane
      const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t td = opaque(tmem_base + 128 * ch + lane_bits);
      const uint32_t th = td + 64, tl = td + 96;
      const uint32_t cba = opaque(smem_u32(&bars[ch]));   // a[0] +0, a[1] +8, d[0] +16, d[1] +24
      const int bar_id = (int)opaque((uint32_t)(1 + 2 * ch));
      const float* s_prm = opaque_ptr(s_params);          // (kept in registers: the compiler otherwise re-derives the shared window per stage)
      const uint32_t prma = opaque(smem_u32(s_params));
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
#ifdef UORF_TIMELINE_X3
      const bool tl_on = blockIdx.x == 0 && warp == 0 && lane == 0 && phase == 1;
#else
      const bool tl_on = false;
#endif
      int tl_n = 0;
      // the NEXT tile's coordinates (and, in the fg sweep, this tile's bg result) are fetched in the middle of the current tile:
      // a DRAM round trip at every tile boundary sat on the channel's chain otherwise
      float pfx = 0.f, pfy = 0.f, pfz = 0.f;
      {
        const long long pt0 = ((long long)blockIdx.x + (long long)gridDim.x * ch) * kTileM + t;
        if (pt0 < p.P) {
          const float* src = (phase == 0 ? p.coor_bg : p.coor_fg) + pt0 * 3;
          pfx = __ldg(src); pfy = __ldg(src + 1); pfz = __ldg(src + 2);
        }
      }
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * ((long long)it * kX3Ch + ch);
        if (tile >= p.num_tiles) break;
        const long long pt = tile * kTileM + t;
        const bool valid = pt < p.P;
        // coordinates are fetched one slot ahead (and once per tile when all fg slots share them, fg_slot_stride == 0)
        float nx = pfx, ny = pfy, nz = pfz;
        float4 bgv = make_float4(0.f, 0.f, 0.f, 0.f);
        // When every fg slot sees the same coordinates (the reference's expand() view, no moved object) the encoding operand
        // written for the first slot is still in the strip, and the locality flag is the same: later slots skip the encoding.
        const bool shared_enc = phase == 1 && p.fg_slot_stride == 0 && p.fg_slot_offset == nullptr;
        float x_last = 0.f;
        bool outside = false;
        for (int s = 0; s < nslots; ++s) {
          const float* sb = s_prm + kParamFloats + s * kSlotBiasFloats;
          const uint32_t sba = prma + 4 * (kParamFloats + s * kSlotBiasFloats);
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
            posenc48<PASSES == 1>(x, y, z, v);
            posenc_store_smem<F16>(v, strip, t);
            fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
            x_last = v[kPosencMma3];           // cos(16 z): added by FMA in the epilogues of layers before.0 / after.0
          }
          if (phase == 1 && p.outside && valid) p.outside[(long long)s * p.P + pt] = outside ? 1 : 0;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            uint32_t w[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b = lds128(sba + 64 * h + 16 * q);
              w[4 * q + 0] = __float_as_uint(b.x); w[4 * q + 1] = __float_as_uint(b.y);
              w[4 * q + 2] = __float_as_uint(b.z); w[4 * q + 3] = __float_as_uint(b.w);
            }
            tmem_st16(td + 16 * h, w);
          }
          tc_wait_st();
          tc_fence_before();
          x3_arrive(cba);
          if (UORF_X3_KSPLIT) x3_arrive(cba + 8);
          TLX_EPI(0x01);
          if (s == nslots - 1) {   // prefetch for the tile boundary
            const long long npt = pt + (long long)gridDim.x * kX3Ch * kTileM;
            if (npt < p.P) {
              const float* src = (phase == 0 ? p.coor_bg : p.coor_fg) + npt * 3;
              pfx = __ldg(src); pfy = __ldg(src + 1); pfz = __ldg(src + 2);
            }
            if (phase == 1 && valid) bgv = __ldcs(p.scratch_bg + pt);   // last use: streaming load
          }

          // ---- stages 1..6: hidden layers, two halves each ----
#pragma unroll 1
          for (int st = 1; st <= 6; ++st) {
            const uint32_t par = n_stage & 1;
            ++n_stage;
            // bias of the layer MMA stage `st` computes (preloaded into the accumulator; the head, st == 6, keeps its bias in registers)
            // (plain arithmetic on the stage index: a chain of selects became a jump table, ~200 cycles per stage on the tile's chain)
            static_assert(kPB1 == 0 && kPB2 == 64 && kPA1 == 128 && kPA2 == 192 && kPW3x == kPW0x + 64, "parameter block layout");
            const uint32_t nb = st == 3 ? sba + 256 : prma + 256 * (st - 1 - (st > 3 ? 1 : 0));   // shared addresses
            const uint32_t wx = prma + 4 * kPW0x + (st == 4 ? 256 : 0);
            const bool has_wx = (opaque(0x12u) >> st) & 1u;   // st == 1 || st == 4 (kept as a bit test: the comparison pair became a jump table)
            const bool has_bias = st != 6;
            float4 b0[8], b1[8];
            if (UORF_X3_PREFETCH) x3_load_bias(nb, b0);   // fetched before the wait: no shared-memory latency behind it (st == 6: not used)
            x3_wait_d(cba, bar_id, 0, par, p.err_flag, 300 + st);
            TLX_EPI(0x10 + st);
            if (!UORF_X3_PREFETCH) x3_load_bias(nb, b0);
            {
              uint32_t hi[16], lo[16];
              x3_convert<PASSES, F16>(td, has_wx, wx, x_last, has_bias, b0, hi, lo);
              if (UORF_X3_PREFETCH) x3_load_bias(nb + 128, b1);   // second half's bias: fetched under the gate wait
              TLX_EPI(0x20 + st);
              // The conversions have to be complete BEFORE the gate (the point of E(0) is to run under G(.,1)); ptxas otherwise sinks
              // the pack / FMA instructions below the wait.  So the barrier address is made to depend on every residual word, times
              // a kernel parameter that is zero at run time (which the assembler cannot know).
              uint32_t any = 0;
#pragma unroll
              for (int q = 0; q < 16; ++q) any |= lo[q];
              any *= (uint32_t)p.zero;
              x3_wait_d(cba + any, bar_id + (int)any, 1, par, p.err_flag, 310);
              TLX_EPI(0x30 + st);
#if UORF_X3_EARLY_D1
              // d[1] has been waited for: the second accumulator half is read (and the next layer's bias stored over it) BEFORE the
              // first hand-off, so a[0] also means "D[32:64] is free" and the issuer can run G(0,1) [and P(1)] under E(1) as well
              uint32_t r1[32];
              tmem_ld32(td + 32, r1);
              tmem_st16(th, hi);
              if (PASSES == 3) tmem_st16(tl, lo);
              tc_wait_ld();
              if (has_bias) x3_preload32(td + 32, nb + 128);
              tc_wait_st();
              tc_fence_before();
              x3_arrive(cba);
              TLX_EPI(0x40 + st);
              x3_split32<PASSES, F16>(r1, has_wx, wx + 128, x_last, hi, lo);
              TLX_EPI(0x28 + st);
              TLX_EPI(0x38 + st);
              tmem_st16(th + 16, hi);
              if (PASSES == 3) tmem_st16(tl + 16, lo);
              tc_wait_st();
              tc_fence_before();
              x3_arrive(cba + 8);
              TLX_EPI(0x48 + st);
            }
#else
              tmem_st16(th, hi);
              if (PASSES == 3) tmem_st16(tl, lo);
              if (UORF_X3_KSPLIT) {
                tc_wait_st();
                tc_fence_before();
                x3_arrive(cba);
              }
              TLX_EPI(0x40 + st);
            }
            {
              uint32_t hi[16], lo[16];
              if (!UORF_X3_PREFETCH) x3_load_bias(nb + 128, b1);
#if UORF_X3_EARLY_ST
              x3_convert<PASSES, F16, true>(td + 32, has_wx, wx + 128, x_last, has_bias, b1, hi, lo, th + 16, tl + 16);
#else
              x3_convert<PASSES, F16>(td + 32, has_wx, wx + 128, x_last, has_bias, b1, hi, lo);
              TLX_EPI(0x28 + st);
              TLX_EPI(0x38 + st);
              tmem_st16(th + 16, hi);
              if (PASSES == 3) tmem_st16(tl + 16, lo);
#endif
              tc_wait_st();
              tc_fence_before();
              x3_arrive(cba + (UORF_X3_KSPLIT ? 8 : 0));
              TLX_EPI(0x48 + st);
            }
#endif
          }
          // ---- stage 7: heads ----
          {
            const uint32_t par = n_stage & 1;
            ++n_stage;
            x3_wait_d(cba, bar_id, 0, par, p.err_flag, 307);
            x3_wait_d(cba, bar_id, 1, par, p.err_flag, 308);
            TLX_EPI(0x17);
          }
          if (phase == 0) {
            uint32_t r[16];
            tmem_ld16(td, r);
            tc_wait_ld();
            const float* hb = s_prm + kPHeadB;
            float4 o;
            o.x = __uint_as_float(r[0]) + hb[0];
            o.y = __uint_as_float(r[1]) + hb[1];
            o.z = __uint_as_float(r[2]) + hb[2];
            o.w = __uint_as_float(r[3]) + hb[3];
            if (valid) st_global_hint(p.scratch_bg + pt, o, l2_policy_evict_last());
          } else {
            uint32_t r[32];
            tmem_ld32(td, r);
            tc_wait_ld();
            const float* hb = s_prm + kPHeadB;
            const float* wc2 = s_prm + kPWc2;
            float c0 = s_prm[kPBc2 + 0], c1 = s_prm[kPBc2 + 1], c2 = s_prm[kPBc2 + 2];
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
          if (valid) stash[t] = bgv;
#if UORF_X3_DISCARD_BG
          // every lane's read-back value has been consumed: lanes 0..3 drop the warp's four scratch lines (32 points x 16 B)
          __syncwarp();
          if (p.discard_bg && lane < 4) {
            const long long line_pt = pt - lane + 8 * lane;
            if (line_pt + 8 <= p.P) l2_discard_line(p.scratch_bg + line_pt);
          }
#endif
          float S = 0.f;
#pragma unroll 4
          for (int k = 0; k < K; ++k) S += fmaxf(stash[k * kTileM + t].w, 0.f);
          const float rdenom = UORF_X3_FAST_ACT ? __frcp_rn(S + 1e-5f) : 1.f / (S + 1e-5f);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
          for (int k = 0; k < K; ++k) {
            const float4 rw = stash[k * kTileM + t];
            const float dens = fmaxf(rw.w, 0.f);
            const float m = dens * rdenom;
            float4 u;
            u.x = rgb_act<(PASSES == 1) || (UORF_X3_FAST_ACT != 0)>(rw.x);
            u.y = rgb_act<(PASSES == 1) || (UORF_X3_FAST_ACT != 0)>(rw.y);
            u.z = rgb_act<(PASSES == 1) || (UORF_X3_FAST_ACT != 0)>(rw.z);
            u.w = dens;
            if (p.noise != nullptr && valid) u.w = dens + p.dens_noise * __ldg(p.noise + (long long)k * p.P + pt);
            const float4 mk = make_float4(u.x * m, u.y * m, u.z * m, u.w * m);
            acc.x += mk.x; acc.y += mk.y; acc.z += mk.z; acc.w += mk.w;
            if (valid) {
              if (p.unmasked) __stcs(p.unmasked + (long long)k * p.P + pt, u);
              if (p.masked) __stcs(p.masked + (long long)k * p.P + pt, mk);
              if (p.masks) __stcs(p.masks + (long long)k * p.P + pt, m);
            }
          }
          if (valid && p.raws != nullptr) __stcs(p.raws + pt, acc);
          TLX_EPI(0x02);
        }
      }
      tc_fence_before();
      named_bar_sync(0, kX3Threads);
      tc_fence_after();
    }
  }
}

size_t decoder_fwd_x3_smem_bytes(int K, int passes) {
  const int w_region = (passes == 3 ? 2 : 1) * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  return 1024 + ((w_region + 1023) & ~1023) + (size_t)kX3Ch * kX3StripBytes;
}

int launch_decoder_fwd_x3(const DecFwdParams& p, int passes, bool f16, int num_sms, cudaStream_t stream) {
  const size_t smem = decoder_fwd_x3_smem_bytes(p.K, passes);
  if (smem > 227 * 1024) return UORF_ERR_UNSUPPORTED;
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (grid < 1) return UORF_OK;
  auto launch = [&](auto kern) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UORF_ERR_DEVICE;
    kern<<<grid, kX3Threads, smem, stream>>>(p);
    return UORF_OK;
  };
  const int rc = passes == 3 ? (f16 ? launch(decoder_fwd_x3_kernel<3, true>) : launch(decoder_fwd_x3_kernel<3, false>))
                             : (f16 ? launch(decoder_fwd_x3_kernel<1, true>) : launch(decoder_fwd_x3_kernel<1, false>));
  if (rc != UORF_OK) return rc;
  return finish_launch(1);
}

}  // namespace uorf

#ifdef UORF_TIMELINE_X3
extern "C" int uorf_debug_timeline_x3(long long* epi, long long* iss, int n) {
  if (n > 16384) n = 16384;
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(epi, g_tlx_epi, n * sizeof(long long)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(iss, g_tlx_iss, n * sizeof(long long)) != cudaSuccess) return -1;
  return 0;
}
#endif
