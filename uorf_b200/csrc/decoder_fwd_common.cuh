// Shared pieces of the fused decoder forward kernels (decoder_fwd.cu: 1-pass modes and the legacy 3-pass form;
// decoder_fwd_x3.cu: the 3-pass fp32-parity kernel): parameter block, positional encoding, epilogue building blocks,
// grouped tcgen05.mma issue.  See decoder_fwd.cu for the design notes.
#pragma once
#include "common.cuh"
#include "decoder_layout.h"
#include "../../include/uorf_b200.h"

// Diagnostic: only the first UORF_ACTIVE_CHANNELS channels take tiles (the others idle; output is then incomplete) - used with
// the timeline variant to see a channel's uncontended stage times.
#ifndef UORF_ACTIVE_CHANNELS
#define UORF_ACTIVE_CHANNELS 99
#endif
// Tuning experiments kept as compile-time switches (uorf_b200/build.py --tag builds a variant; scripts/gpu_variants.sh A/Bs it):
// One mbarrier arrival per epilogue WARP (after __syncwarp) instead of one per thread: measured 2 % slower, off.
#ifndef UORF_WARP_ARRIVE
#define UORF_WARP_ARRIVE 0
#endif
// 3-pass modes: the epilogue writes the NEXT layer's bias into the accumulator columns it has just read (tcgen05.st) and the
// layer's first MMA accumulates on top of it: 2 instructions per 32 columns instead of 32 FADDs in the following epilogue.
// Neutral in the round-1 form (3 channels x 8 warps: 3.47 -> 3.58 ms); with one thread per point (4 channels x 4 warps), where
// the epilogue of a tile and layer is one warp's serial instruction stream, it is worth 4-6 % (3.50 -> 3.34 ms): on.
#ifndef UORF_X3_BIAS_PRELOAD
#define UORF_X3_BIAS_PRELOAD 1
#endif
// Bias through the selector MMA in the 3-pass modes too (tuning experiment, off: measured 7 % (bf16x3) to 20 % (fp16x3) SLOWER -
// the two extra MMAs per layer sit on the dependent chain, which is what bounds these modes)
#ifndef UORF_X3_BIAS_MMA
#define UORF_X3_BIAS_MMA 0
#endif

namespace uorf {

constexpr int kTileM = 128;
constexpr int kTmemCols = 512;

template <int PASSES, int NCH_, bool PSM_ = false>
struct Cfg {
  static constexpr int NCH = NCH_;                      // channels (tiles in flight): 4 (1-pass; 3-pass PSM), legacy 3-pass: 3 or 2
  static constexpr bool PSM = PSM_;                     // positional-encoding operand in shared memory (A of SS MMAs), stash in global
  static constexpr int CS = (PASSES == 3 && !PSM) ? 2 : 1;   // threads per point (column split)
  static constexpr int WPC = 4 * CS;                    // warps per channel
  static constexpr int EPI_WARPS = NCH * WPC;           // 16 (1-pass, PSM), 24 / 16 (legacy 3-pass)
  static constexpr int THREADS = (EPI_WARPS + NCH) * 32;
  static constexpr int NKP = PASSES == 3 ? 2 : 3;       // K-steps of the positional-encoding operand (3-pass: elements 0..31
                                                        // on the tensor core, element 32 by an fp32 FMA in the epilogue)
  // TMEM map.  1-pass: channel c owns columns 128c .. 128c+127 = D 64 | H 32 | P 24 | selector 8.
  // 3-pass PSM: channel c owns 128c .. 128c+127 = D 64 | H hi 32 | H lo 32.
  // 3-pass legacy: accumulators first (D of channel c at 64c), then one 104-column operand block per channel at 192 + 104c =
  //         H hi 32 | H lo 32 | P hi 16 | P lo 16 | selector 8   (3 x 64 + 3 x 104 = 504 <= 512 columns)
  static constexpr int d_col(int c) { return (PASSES == 3 && !PSM) ? 64 * c : 128 * c; }
  static constexpr int a_col(int c) { return (PASSES == 3 && !PSM) ? 192 + 104 * c : 128 * c + 64; }
  static constexpr int COL_HHI = 0;                     // offsets inside the operand block
  static constexpr int COL_HLO = 32;                    // 3-pass only
  static constexpr int COL_PHI = PASSES == 3 ? 64 : 32; // (not with PSM)
  static constexpr int COL_PLO = 80;                    // legacy 3-pass only
  static constexpr int COL_SEL = PASSES == 3 ? 96 : 56; // 8 columns: one-hot K-step that selects a bias column (not with PSM)
  static constexpr int NPL = PASSES == 3 ? 2 : 1;       // weight planes
  // bias through the tensor core (selector MMA) pays off where the epilogue is the bottleneck (1-pass modes); in the 3-pass
  // modes the extra MMAs lengthen the dependent chain more than two FADDs per pair cost, so the bias stays in the epilogue
  static constexpr bool BIAS_MMA = !PSM && (PASSES == 1 || UORF_X3_BIAS_MMA);
  static constexpr int kStripBytes = kTileM * 128;      // PSM: one P strip (128 rows x 128 B)
};

struct DecFwdParams {
  const float* coor_bg;
  const float* coor_fg;
  long long fg_slot_stride;
  const float* noise;
  const unsigned char* image;
  float4* scratch_bg;
  float4* stash_g;      // PSM variant: [grid][NCH][K][128] per-slot head results (L2-resident, rewritten every tile)
  float4* raws;
  float4* masked;
  float4* unmasked;
  float* masks;
  unsigned char* outside;
  const float* fg_transform;  // device: 3x3, or 4x4 when fixed_locality
  const float* fg_slot_offset;  // device [K-1][3] or null
  float locality_ratio;
  float dens_noise;
  long long P;
  int K;
  int fixed_locality;
  int locality;
  int num_tiles;
  int discard_bg;       // scratch_bg is 128-byte aligned: read-back lines may be dropped from L2 (decoder_fwd_x3.cu)
  int zero;             // always 0: run-time zero for scheduling anchors (decoder_fwd_x3.cu)
  int* err_flag;
};

struct ChanBars {
  uint64_t a_full;      // A operand of the next layer written (all threads of the channel)
  uint64_t d_full;      // MMAs of the current layer complete (tcgen05.commit)
};

// ---- epilogue-side helpers ------------------------------------------------------------------------

// positional encoding of one point, reference order [x y z | sin(2^i x..z) cos(2^i x..z)]_{i<5}
// (model.py:261-277), zero padded to 48.  Frequencies 2,4,8,16 by angle doubling.
// FAST: MUFU sin/cos (abs error ~5e-7 on the |x| <= 2 NSS range; one-pass modes only), else the accurate sincosf.
template <bool FAST>
__device__ __forceinline__ void posenc48(float x, float y, float z, float (&v)[kPosencK]) {
  v[0] = x; v[1] = y; v[2] = z;
  float s[3], c[3];
  if (FAST) {
    __sincosf(x, &s[0], &c[0]);
    __sincosf(y, &s[1], &c[1]);
    __sincosf(z, &s[2], &c[2]);
  } else {
    sincosf(x, &s[0], &c[0]);
    sincosf(y, &s[1], &c[1]);
    sincosf(z, &s[2], &c[2]);
  }
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
}

// write values v[2*W0 .. 2*(W0+NW)) as NW packed words at TMEM columns W0.. of the P operand
template <int PASSES, bool F16, int W0, int NW>
__device__ __forceinline__ void posenc_store(const float (&v)[kPosencK], uint32_t t_phi, uint32_t t_plo) {
  uint32_t hi[NW], lo[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    if (PASSES == 3) split2<F16>(v[2 * (W0 + i)], v[2 * (W0 + i) + 1], hi[i], lo[i]);
    else hi[i] = pack2<F16>(v[2 * (W0 + i)], v[2 * (W0 + i) + 1]);
  }
  static_assert(NW == 24 || NW == 8, "posenc split");
  if constexpr (NW == 8) {
    tmem_st8(t_phi + W0, hi);
    if (PASSES == 3) tmem_st8(t_plo + W0, lo);
  } else if constexpr (NW == 24) {
    tmem_st16(t_phi + W0, hi);
    tmem_st8(t_phi + W0 + 16, hi + 16);
    if (PASSES == 3) { tmem_st16(t_plo + W0, lo); tmem_st8(t_plo + W0 + 16, lo + 16); }
  }
}

// PSM: elements 0..31 of the encoding as one 128-byte K-major row [hi 32 | lo 32] of the channel's shared-memory strip
// (row t at (t / 8) * 1024 + (t % 8) * 128, 16-byte chunk c stored at chunk c ^ (t & 7): the layout make_sdesc_sw128 describes).
// Lanes t .. t+7 of a quarter warp hit 8 different chunk positions: the 16-byte stores are bank-conflict free.
template <bool F16>
__device__ __forceinline__ void posenc_store_smem(const float (&v)[kPosencK], unsigned char* strip, int t) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) split2<F16>(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  unsigned char* row = strip + (t >> 3) * 1024 + (t & 7) * 128;
  const int sw = t & 7;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
    *reinterpret_cast<uint4*>(row + (((4 + c) ^ sw) << 4)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
  }
}

// 32 accumulator columns (bias already added by the selector MMA): ReLU -> 16-bit hi[/lo] -> 16 H operand columns
// wx / xe (3-pass modes, layers fed by the positional encoding): adds wx[col] * xe, the contribution of encoding element 32
template <int PASSES, bool F16>
__device__ __forceinline__ void epilogue32(uint32_t t_d, const float* bias, const float* wx, float xe, uint32_t t_hhi, uint32_t t_hlo) {
  uint32_t r[32];
  tmem_ld32(t_d, r);
  tc_wait_ld();
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    // ReLU is folded into the 16-bit conversion; the bias is either already in the accumulator (selector MMA) or added here
    float v0 = __uint_as_float(r[4 * q + 0]);
    float v1 = __uint_as_float(r[4 * q + 1]);
    float v2 = __uint_as_float(r[4 * q + 2]);
    float v3 = __uint_as_float(r[4 * q + 3]);
    if (PASSES == 3 && !UORF_X3_BIAS_MMA && !UORF_X3_BIAS_PRELOAD) {
      float4 b = *reinterpret_cast<const float4*>(bias + 4 * q);
      if (wx != nullptr) {
        const float4 w = *reinterpret_cast<const float4*>(wx + 4 * q);
        b.x = fmaf(w.x, xe, b.x); b.y = fmaf(w.y, xe, b.y); b.z = fmaf(w.z, xe, b.z); b.w = fmaf(w.w, xe, b.w);
      }
      v0 += b.x; v1 += b.y; v2 += b.z; v3 += b.w;
    } else if (PASSES == 3 && wx != nullptr) {
      const float4 w = *reinterpret_cast<const float4*>(wx + 4 * q);
      v0 = fmaf(w.x, xe, v0); v1 = fmaf(w.y, xe, v1); v2 = fmaf(w.z, xe, v2); v3 = fmaf(w.w, xe, v3);
    }
    if (PASSES == 3 && F16) {
      split2_relu<F16>(v0, v1, hi[2 * q], lo[2 * q]);
      split2_relu<F16>(v2, v3, hi[2 * q + 1], lo[2 * q + 1]);
    } else if (PASSES == 3) {   // bf16: keep round-to-nearest on the hi part (truncation would cost one of its 16 bits)
      split2<F16>(fmaxf(v0, 0.f), fmaxf(v1, 0.f), hi[2 * q], lo[2 * q]);
      split2<F16>(fmaxf(v2, 0.f), fmaxf(v3, 0.f), hi[2 * q + 1], lo[2 * q + 1]);
    } else {
      hi[2 * q] = pack2_relu<F16>(v0, v1);
      hi[2 * q + 1] = pack2_relu<F16>(v2, v3);
    }
  }
  tmem_st16(t_hhi, hi);
  if (PASSES == 3) tmem_st16(t_hlo, lo);
}

// 32 accumulator columns <- bias[0..31] (the same vector on every row / TMEM lane)
__device__ __forceinline__ void preload_bias32(uint32_t t_d, const float* bias) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = *reinterpret_cast<const float4*>(bias + 16 * h + 4 * q);
      r[4 * q + 0] = __float_as_uint(b.x); r[4 * q + 1] = __float_as_uint(b.y);
      r[4 * q + 2] = __float_as_uint(b.z); r[4 * q + 3] = __float_as_uint(b.w);
    }
    tmem_st16(t_d + 16 * h, r);
  }
}

// one-hot selector K-step (16 values -> 8 words) with 1.0 at position (bias_id & 15)
template <bool F16>
__device__ __forceinline__ void write_selector(uint32_t t_sel, int bias_id) {
  const uint32_t one = F16 ? 0x3C00u : 0x3F80u;
  const int pos = bias_id & 15;
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = (pos >> 1) == i ? (one << ((pos & 1) * 16)) : 0u;
  tmem_st8(t_sel, w);
}

// colour activation (tanh(x) + 1) / 2 = 1 / (1 + exp(-2x)); FAST uses ex2.approx + fast division (abs error ~1e-6)
template <bool FAST>
__device__ __forceinline__ float rgb_act(float x) {
  if (FAST) return __frcp_rn(1.f + __expf(-2.f * x));
  return (tanhf(x) + 1.f) * 0.5f;
}

// hand-off epilogue -> issuer: every thread has completed and fenced its tcgen05.st; the warp converges and one lane arrives
__device__ __forceinline__ void chan_arrive(uint64_t* bar) {
#if UORF_WARP_ARRIVE
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
#else
  mbar_arrive(bar);
#endif
}

// ---- issuer-side helper ---------------------------------------------------------------------------
// D (+)= A[tmem, K-steps 0..NK) . W[tile]^T, in 1 or 3 passes, issued under ONE elect.sync: the whole group is a single
// asm block, so per MMA the issuer spends the tcgen05.mma itself plus the operand moves - not an elect / vote / descriptor
// rebuild each (the issuer warp shares its scheduler with six epilogue warps; every instruction it saves shortens the
// dependent chain of its channel).  K-step j: A columns + 8j, B descriptor + 2j (32 bytes >> 4).
#define UORF_MMA(A, B, ACC) "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [" A "], " B ", %5, " ACC ";\n\t"
template <int NK, int PASSES>
__device__ __forceinline__ void mma_group_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                             uint32_t accumulate) {
  static_assert((PASSES == 1 && (NK == 3 || NK == 4)) || (PASSES == 3 && (NK == 2 || NK == 4)), "unsupported group");
  if constexpr (PASSES == 1 && NK == 3) {
    asm volatile(
        "{\n\t.reg .pred e, p, t;\n\t.reg .b32 a1, a2;\n\t.reg .b64 h1, h2;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, %6, %6;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\t"
        "add.u64 h1, %3, 2;\n\tadd.u64 h2, %3, 4;\n\t"
        UORF_MMA("%1", "%3", "p") UORF_MMA("a1", "h1", "t") UORF_MMA("a2", "h2", "t")
        "}" ::"r"(d), "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate) : "memory");
  } else if constexpr (PASSES == 1 && NK == 4) {
    asm volatile(
        "{\n\t.reg .pred e, p, t;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 h1, h2, h3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, %6, %6;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "add.u64 h1, %3, 2;\n\tadd.u64 h2, %3, 4;\n\tadd.u64 h3, %3, 6;\n\t"
        UORF_MMA("%1", "%3", "p") UORF_MMA("a1", "h1", "t") UORF_MMA("a2", "h2", "t") UORF_MMA("a3", "h3", "t")
        "}" ::"r"(d), "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate) : "memory");
  } else if constexpr (PASSES == 3 && NK == 2) {
    asm volatile(
        "{\n\t.reg .pred e, p, t;\n\t.reg .b32 a1, l1;\n\t.reg .b64 h1, w1;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, %6, %6;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 l1, %2, 8;\n\t"
        "add.u64 h1, %3, 2;\n\tadd.u64 w1, %4, 2;\n\t"
        UORF_MMA("%1", "%3", "p") UORF_MMA("a1", "h1", "t")      // a_hi . w_hi
        UORF_MMA("%2", "%3", "t") UORF_MMA("l1", "h1", "t")      // a_lo . w_hi
        UORF_MMA("%1", "%4", "t") UORF_MMA("a1", "w1", "t")      // a_hi . w_lo
        "}" ::"r"(d), "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred e, p, t;\n\t.reg .b32 a1, a2, a3, l1, l2, l3;\n\t.reg .b64 h1, h2, h3, w1, w2, w3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, %6, %6;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "add.u32 l1, %2, 8;\n\tadd.u32 l2, %2, 16;\n\tadd.u32 l3, %2, 24;\n\t"
        "add.u64 h1, %3, 2;\n\tadd.u64 h2, %3, 4;\n\tadd.u64 h3, %3, 6;\n\t"
        "add.u64 w1, %4, 2;\n\tadd.u64 w2, %4, 4;\n\tadd.u64 w3, %4, 6;\n\t"
        UORF_MMA("%1", "%3", "p") UORF_MMA("a1", "h1", "t") UORF_MMA("a2", "h2", "t") UORF_MMA("a3", "h3", "t")
        UORF_MMA("%2", "%3", "t") UORF_MMA("l1", "h1", "t") UORF_MMA("l2", "h2", "t") UORF_MMA("l3", "h3", "t")
        UORF_MMA("%1", "%4", "t") UORF_MMA("a1", "w1", "t") UORF_MMA("a2", "w2", "t") UORF_MMA("a3", "w3", "t")
        "}" ::"r"(d), "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate) : "memory");
  }
}
#undef UORF_MMA
// PSM: D (+)= P[smem strip] . W[tile]^T for the two encoding K-steps, three passes, both operands from shared memory (SS).
// Strip row = [hi K-steps 0,1 | lo K-steps 2,3]; descriptors advance by 2 (32 bytes >> 4) per K-step.
#define UORF_MMA_SS(A, B, ACC) "@e tcgen05.mma.cta_group::1.kind::f16 [%0], " A ", " B ", %4, " ACC ";\n\t"
__device__ __forceinline__ void mma_group_p_ss(uint32_t d, uint64_t a, uint64_t b_hi, uint64_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred e, p, t;\n\t.reg .b64 a1, a2, a3, h1, w1;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\tsetp.eq.b32 t, %5, %5;\n\t"
      "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
      "add.u64 h1, %2, 2;\n\tadd.u64 w1, %3, 2;\n\t"
      UORF_MMA_SS("%1", "%2", "p") UORF_MMA_SS("a1", "h1", "t")      // p_hi . w_hi
      UORF_MMA_SS("a2", "%2", "t") UORF_MMA_SS("a3", "h1", "t")      // p_lo . w_hi
      UORF_MMA_SS("%1", "%3", "t") UORF_MMA_SS("a1", "w1", "t")      // p_hi . w_lo
      "}" ::"r"(d), "l"(a), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate) : "memory");
}
#undef UORF_MMA_SS
template <int NK, int PASSES>
__device__ __forceinline__ void issue_group(uint32_t t_d, uint32_t t_ahi, uint32_t t_alo, uint32_t w_smem, uint32_t idesc, bool& first) {
  mma_group_ts<NK, PASSES>(t_d, t_ahi, t_alo, make_sdesc_sw128(w_smem), make_sdesc_sw128(w_smem + kPlaneBytes), idesc, first ? 0u : 1u);
  first = false;
}

}  // namespace uorf
