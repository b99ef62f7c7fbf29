// Kernel 2 — fused K-slot decoder forward on tcgen05 / TMEM.
//
// Replaces Decoder.forward (reference models/model.py:106-161) and sin_emb (:261-277):
//   per-slot object-frame transform + locality box test, sin/cos positional encoding, the background
//   MLP (7 linears) and the shared foreground MLP (9 linears, f_after_latent folded into f_color.0),
//   ReLU density / tanh colour, density-weighted composition over the K slots.
// No activation ever leaves the SM: the only HBM traffic is coords in, API-mandated outputs out.
//
// Structure: one persistent CTA per SM, epilogue warps grouped into NCH "channels" + one MMA-issuer warp per channel.
//   A channel owns one 128-point tile at a time and walks it through bg MLP / the K-1 fg slots.  Thread <-> TMEM lane <-> point.
//     3-pass modes : NCH = 4 channels x 4 warps, one thread per point (PSM variant, the default).  TMEM per channel: D 64 | H hi 32 |
//                    H lo 32 = 128 columns, so FOUR tiles are in flight in the 512 columns.  The positional-encoding operand P
//                    (2 of a slot's 26 K-steps) lives in SHARED memory instead (one 16 KB strip per channel: 128 rows x [hi 32 | lo 32]
//                    16-bit elements, K-major, 128-byte swizzle, written by the epilogue threads with st.shared + fence.proxy.async) and
//                    is the A operand of SS MMAs (52 instead of 34.6 cycles, for those two K-steps only); the per-slot head results
//                    are stashed in a small L2-resident global buffer instead of shared memory.
//                    Only 32 of the 33 encoding elements go through the tensor core; cos(16 z) is an fp32 FMA in the epilogue.
//                    (Previous form, kept as -DUORF_FWD_X3_LEGACY: 3 channels x 8 warps, P in TMEM, 160 columns per channel.)
//     1-pass modes : NCH = 4 channels x 4 warps (TMEM per channel: D 64 | H 32 | P 24 | selector 8 columns)
//   warp EPI + c    MMA issuer of channel c (elected lane issues tcgen05.mma); the first issuer warp also owns TMEM alloc and the
//                   weight bulk copies.  One issuer per channel keeps the channels out of lock-step: while one channel is in
//                   its epilogue the tensor pipe works for another.
// Per layer:  A (activations, 16-bit hi [+lo]) lives in TMEM, written by the epilogue threads with tcgen05.st;
//   B (weights) is a 128B-swizzled K-major tile in shared memory; D (fp32) in TMEM.
//   epilogue --a_full mbarrier--> issuer --tcgen05.commit / d_full mbarrier--> epilogue (ReLU + 16-bit split).
//   1-pass modes: biases never touch the CUDA cores: the weight image carries a bias tile whose COLUMN j is bias vector j, the epilogue
//   writes a one-hot "selector" K-step (8 TMEM columns) naming the next layer's bias, and the issuer appends one MMA
//   D += selector . bias_tile^T to the layer - the accumulator the epilogue reads already contains the bias.
// Precision: 1 pass of 16-bit operands, or 3 passes a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (fp32 accumulate);
//   operands bf16 (range-safe, ~16 bits after the split) or fp16 (saturating, ~22 bits after the split).
// The kernel makes two sweeps over its tiles: background MLP first (raw bg output parked in a 16 B/point
// scratch), then the foreground MLP for slots 1..K-1 followed by the cross-slot composition.  The split
// lets the 3-pass weight image of ONE MLP (136 KB) stay resident in shared memory.
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
// Measured neutral (the epilogue is bound by its dependent chain, not by issue slots), off.
#ifndef UORF_X3_BIAS_PRELOAD
#define UORF_X3_BIAS_PRELOAD 0
#endif
// Bias through the selector MMA in the 3-pass modes too (tuning experiment, off: measured 7 % (bf16x3) to 20 % (fp16x3) SLOWER -
// the two extra MMAs per layer sit on the dependent chain, which is what bounds these modes)
#ifndef UORF_X3_BIAS_MMA
#define UORF_X3_BIAS_MMA 0
#endif

// -DUORF_TIMELINE (tuning variant only, scripts/timeline.py): CTA 0 / channel 0 records clock64 stamps of the hand-off
// points of every stage into a global buffer read back through uorf_debug_timeline()
#ifdef UORF_TIMELINE
__device__ long long g_tl_epi[8192];
__device__ long long g_tl_iss[8192];
#define TL_EPI(id) do { if (tl_on && tl_ne < 8192) g_tl_epi[tl_ne++] = (clock64() << 8) | (id); } while (0)
#define TL_ISS(id) do { if (tl_on && tl_ni < 8192) g_tl_iss[tl_ni++] = (clock64() << 8) | (id); } while (0)
#else
#define TL_EPI(id)
#define TL_ISS(id)
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
  if (FAST) return __fdividef(1.f, 1.f + __expf(-2.f * x));
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

template <int PASSES, bool F16, int NCH, bool PSM = false>
__global__ void __launch_bounds__(Cfg<PASSES, NCH, PSM>::THREADS, 1) decoder_fwd_kernel(const DecFwdParams p) {
  using C = Cfg<PASSES, NCH, PSM>;
  constexpr int kEpiWarps = C::EPI_WARPS;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ __align__(8) ChanBars bars[C::NCH];
  __shared__ uint32_t tmem_base_s;

  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int K = p.K;
  const int w_region = C::NPL * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  const float* s_params = reinterpret_cast<const float*>(smem + C::NPL * kPlaneBytes);
  float4* s_stash = reinterpret_cast<float4*>(smem + ((w_region + 127) & ~127));  // [NCH][K][128]   (not PSM)
  unsigned char* s_strips = smem + ((w_region + 1023) & ~1023);                   // [NCH][128 rows x 128 B] (PSM)

  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int c = 0; c < C::NCH; ++c) {
      mbar_init(&bars[c].a_full, UORF_WARP_ARRIVE ? C::WPC : kTileM * C::CS);
      mbar_init(&bars[c].d_full, 1);
    }
    fence_mbar_init();
  }
  if (warp == kEpiWarps) {
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

  // completion counts of the barriers (parity = count & 1); the epilogue threads and the issuer keep their own copies
  uint32_t n_d = 0;   // epilogue: d_full completions consumed by this thread
  uint32_t n_a = 0;   // issuer: a_full completions consumed

  for (int phase = 0; phase < 2; ++phase) {  // 0: background MLP, 1: foreground MLP + composition
    const int nslots = phase == 0 ? 1 : K - 1;
    // ---- load this phase's weight image ----------------------------------------------------------
    if (warp == kEpiWarps && lane == 0) {
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

    if (warp >= kEpiWarps) {
      // =========================== MMA issuer of channel (warp - EPI_WARPS) ===========================
      // The loop is instantiated once per channel with the channel index as a compile-time constant: every MMA operand is
      // then derived from kernel parameters, constants and one shared-memory word, so the compiler keeps descriptors and
      // TMEM addresses on the uniform datapath (a runtime channel index costs ~10 vector instructions + R2URs per MMA, and
      // the issuer shares its scheduler with six epilogue warps: the hand-off timeline showed 80 cycles per MMA issued).
      auto issuer = [&](auto CH) {
      constexpr int c = decltype(CH)::value;
      const uint32_t idesc64 = make_idesc_f32acc(128, 64, F16);
      const uint32_t idesc_head = make_idesc_f32acc(128, phase == 0 ? 16 : 32, F16);
      const uint32_t w_smem = smem_u32(smem);
      const uint32_t tbd = tmem_base + C::d_col(c);
      const uint32_t tb = tmem_base + C::a_col(c);
#ifdef UORF_TIMELINE
      const bool tl_on = blockIdx.x == 0 && c == 0 && lane == 0 && phase == 1;
      int tl_ni = 0;
#endif
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * ((long long)it * C::NCH + c);
        if (tile >= p.num_tiles || (UORF_ACTIVE_CHANNELS < C::NCH && c >= UORF_ACTIVE_CHANNELS)) break;
        // the seven stages are instantiated separately (compile-time stage index): tile addresses, K-step counts and descriptors
        // are constants on the uniform datapath and the issuer's path between wake-up and commit is short straight-line code
        // (the loop + switch form spent ~500 cycles there per stage on branches and descriptor rebuilds - hand-off timeline)
        auto stage = [&](auto ST, int s) {
          constexpr int st = decltype(ST)::value;
          mbar_wait(&bars[c].a_full, n_a & 1, p.err_flag, 200 + st);
          ++n_a;
          tc_fence_after();
          TL_ISS(0x10 + st);
          bool first = !(PASSES == 3 && UORF_X3_BIAS_PRELOAD && st < 6);   // preloaded bias: accumulate from the first K-step
          const uint32_t t_d = tbd, t_p = tb + C::COL_PHI, t_pl = tb + C::COL_PLO, t_h = tb + C::COL_HHI, t_hl = tb + C::COL_HLO;
          if (C::BIAS_MMA && st < 6) {   // bias of this layer: D = selector . bias_tile^T  (K-step = bias_id / 16), hi [+ lo] plane
            const int bid = st == 0 ? bias_id_slot(s, 0) : st == 3 ? bias_id_slot(s, 1) : bias_id_fixed(st < 3 ? st - 1 : st - 2);
            const uint64_t db = make_sdesc_sw128(w_smem + kBiasTileOff) + 2 * (bid >> 4);
            mma_ts_elect(t_d, tb + C::COL_SEL, db, idesc64, 0u);
            if (PASSES == 3) mma_ts_elect(t_d, tb + C::COL_SEL, make_sdesc_sw128(w_smem + kPlaneBytes + kBiasTileOff) + 2 * (bid >> 4), idesc64, 1u);
            first = false;
          }
          if constexpr (st == 0 && PSM) {
            mma_group_p_ss(t_d, make_sdesc_sw128(smem_u32(s_strips) + c * C::kStripBytes), make_sdesc_sw128(w_smem + 0 * kTileBytes),
                           make_sdesc_sw128(w_smem + kPlaneBytes + 0 * kTileBytes), idesc64, first ? 0u : 1u);
            first = false;
          } else if constexpr (st == 0) {
            issue_group<C::NKP, PASSES>(t_d, t_p, t_pl, w_smem + 0 * kTileBytes, idesc64, first);
          } else if constexpr (st == 3 && PSM) {
            mma_group_p_ss(t_d, make_sdesc_sw128(smem_u32(s_strips) + c * C::kStripBytes), make_sdesc_sw128(w_smem + 3 * kTileBytes),
                           make_sdesc_sw128(w_smem + kPlaneBytes + 3 * kTileBytes), idesc64, first ? 0u : 1u);
            first = false;
            issue_group<4, PASSES>(t_d, t_h, t_hl, w_smem + 4 * kTileBytes, idesc64, first);
          } else if constexpr (st == 3) {
            issue_group<C::NKP, PASSES>(t_d, t_p, t_pl, w_smem + 3 * kTileBytes, idesc64, first);
            issue_group<4, PASSES>(t_d, t_h, t_hl, w_smem + 4 * kTileBytes, idesc64, first);
          } else if constexpr (st == 6) {
            issue_group<4, PASSES>(t_d, t_h, t_hl, w_smem + 7 * kTileBytes, idesc_head, first);
          } else {
            issue_group<4, PASSES>(t_d, t_h, t_hl, w_smem + (st < 3 ? st : st + 1) * kTileBytes, idesc64, first);
          }
          tc_commit_elect(&bars[c].d_full);
          TL_ISS(0x20 + st);
          __syncwarp();
        };
        for (int s = 0; s < nslots; ++s) {
          stage(std::integral_constant<int, 0>{}, s);
          stage(std::integral_constant<int, 1>{}, s);
          stage(std::integral_constant<int, 2>{}, s);
          stage(std::integral_constant<int, 3>{}, s);
          stage(std::integral_constant<int, 4>{}, s);
          stage(std::integral_constant<int, 5>{}, s);
          stage(std::integral_constant<int, 6>{}, s);
        }
      }
      };
      switch (warp - kEpiWarps) {
        case 0: issuer(std::integral_constant<int, 0>{}); break;
        case 1: issuer(std::integral_constant<int, 1>{}); break;
        case 2: if constexpr (C::NCH > 2) issuer(std::integral_constant<int, 2>{}); break;
        default: if constexpr (C::NCH > 3) issuer(std::integral_constant<int, 3>{}); break;
      }
    } else {
      // =========================== epilogue warps ===========================
      const int ch = warp / C::WPC;
      const int wic = warp % C::WPC;          // warp within channel
      const int half = wic >> 2;              // which half of the feature columns (3-pass modes); 0 otherwise
      const int t = (wic & 3) * 32 + lane;    // point within the tile == TMEM lane
      const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t tbd = tmem_base + C::d_col(ch) + lane_bits;             // accumulator
      const uint32_t tb = tmem_base + C::a_col(ch) + lane_bits;              // operand block
      ChanBars& cb = bars[ch];
      // per-slot head results of the tile in flight: shared memory, or (PSM) this channel's slice of a global buffer - each
      // thread only ever reads back what it wrote itself, so no fence is needed, and the 10-40 KB stay L2-resident
      float4* stash = PSM ? p.stash_g + ((size_t)blockIdx.x * C::NCH + ch) * K * kTileM : s_stash + (size_t)ch * K * kTileM;
      unsigned char* strip = s_strips + ch * C::kStripBytes;
      constexpr int NCOL = 64 / C::CS;        // accumulator columns this thread owns per hidden layer
      const int col0 = half * NCOL;
      // wait for the MMAs of the current stage
#ifdef UORF_TIMELINE
      const bool tl_on = blockIdx.x == 0 && ch == 0 && wic == 0 && lane == 0 && phase == 1;
      int tl_ne = 0;
#endif
      auto wait_d = [&](int site) {
        if (C::CS == 2) {
          // 3-pass modes (8 warps per channel): one polling warp, the others sleep in a hardware barrier and spend no issue
          // slots on probing (measured 3 % faster; no difference with 4 warps per channel)
          if (wic == 0) mbar_wait(&cb.d_full, n_d & 1, p.err_flag, site);
          named_bar_sync(13 + ch, C::WPC * 32);
        } else {
          mbar_wait(&cb.d_full, n_d & 1, p.err_flag, site);
        }
        ++n_d;
        tc_fence_after();
        TL_EPI(0x30 + (site & 15));
      };
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * ((long long)it * C::NCH + ch);
        if (tile >= p.num_tiles || (UORF_ACTIVE_CHANNELS < C::NCH && ch >= UORF_ACTIVE_CHANNELS)) break;
        const long long pt = tile * kTileM + t;
        const bool valid = pt < p.P;
        // coordinates are fetched one slot ahead (and once per tile when all fg slots share them, fg_slot_stride == 0), so
        // the global-load latency is off the per-slot critical path
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (valid) {
          const float* src = (phase == 0 ? p.coor_bg : p.coor_fg) + pt * 3;
          nx = __ldg(src); ny = __ldg(src + 1); nz = __ldg(src + 2);
        }
        // When every fg slot sees the same coordinates (the reference's expand() view, no moved object) the encoding operand
        // written for the first slot is still in TMEM, and the locality flag is the same: later slots skip straight to the MMA.
        const bool shared_enc = phase == 1 && p.fg_slot_stride == 0 && p.fg_slot_offset == nullptr;
        float x_last = 0.f;
        bool outside = false;
        for (int s = 0; s < nslots; ++s) {
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
              // unfused multiply/add in the order ((r0*x + r1*y) + r2*z) [+ t]: bit-equal to the reference's
              // CPU matmul
              const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[3]);
              const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z)), T[7]);
              const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z)), T[11]);
              x = tx; y = ty; z = tz;
              if (!p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
            }
            float v[kPosencK];
            posenc48<PASSES == 1>(x, y, z, v);
            if constexpr (PSM) {
              posenc_store_smem<F16>(v, strip, t);
              fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
            } else if (C::CS == 1) {
              posenc_store<PASSES, F16, 0, 24>(v, tb + C::COL_PHI, tb + C::COL_PLO);
            } else if (half == 0) {
              posenc_store<PASSES, F16, 0, 8>(v, tb + C::COL_PHI, tb + C::COL_PLO);
            } else {
              posenc_store<PASSES, F16, 8, 8>(v, tb + C::COL_PHI, tb + C::COL_PLO);
            }
            if (PASSES == 3) x_last = v[kPosencMma3];   // cos(16 z): added by FMA in the epilogues of layers before.0 / after.0
          }
          if (phase == 1 && p.outside && valid && half == 0) p.outside[(long long)s * p.P + pt] = outside ? 1 : 0;
          if (PASSES == 3 && UORF_X3_BIAS_PRELOAD) preload_bias32(tbd + col0, s_params + kParamFloats + s * kSlotBiasFloats + col0);
          if (C::BIAS_MMA && half == 0) write_selector<F16>(tb + C::COL_SEL, bias_id_slot(s, 0));
          tc_wait_st();
          tc_fence_before();
          chan_arrive(&cb.a_full);
          TL_EPI(0x40);

          // ---- stages 1..6: hidden layers ----
          const float* sb = s_params + kParamFloats + s * kSlotBiasFloats;
#pragma unroll 1
          for (int st = 1; st <= 6; ++st) {   // kept rolled: unrolling the six epilogue stages was 10 % slower (instruction cache)
            const float* bias = (st == 1 ? sb : st == 2 ? s_params + kPB1 : st == 3 ? s_params + kPB2
                               : st == 4 ? sb + 64 : st == 5 ? s_params + kPA1 : s_params + kPA2) + col0;
            wait_d(300 + st);
#pragma unroll
            for (int h32 = 0; h32 < NCOL / 32; ++h32) {
              const int c32 = col0 + 32 * h32;   // first accumulator column of this 32-column block
              const float* wx = PASSES == 3 && (st == 1 || st == 4) ? s_params + (st == 1 ? kPW0x : kPW3x) + c32 : nullptr;
              epilogue32<PASSES, F16>(tbd + c32, bias + 32 * h32, wx, x_last, tb + C::COL_HHI + c32 / 2, tb + C::COL_HLO + c32 / 2);
            }
            if (PASSES == 3 && UORF_X3_BIAS_PRELOAD && st < 6)   // bias of the layer MMA stage `st` computes (read back at st + 1)
              preload_bias32(tbd + col0, (st == 1 ? s_params + kPB1 : st == 2 ? s_params + kPB2 : st == 3 ? sb + 64
                                          : st == 4 ? s_params + kPA1 : s_params + kPA2) + col0);
            // selector for the bias of the layer the next MMA stage (index st) computes; the head (st == 6) keeps its bias here
            if (C::BIAS_MMA && half == 0 && st < 6)
              write_selector<F16>(tb + C::COL_SEL, st == 3 ? bias_id_slot(s, 1) : bias_id_fixed(st < 3 ? st - 1 : st - 2));
            TL_EPI(0x50 + st);
            tc_wait_st();
            TL_EPI(0x60 + st);
            tc_fence_before();
            chan_arrive(&cb.a_full);
            TL_EPI(0x40 + st);
          }
          // ---- stage 7: heads (computed by the first column half; every thread consumes the barrier phase) ----
          wait_d(307);
          if (half == 0) {
            if (phase == 0) {
              uint32_t r[16];
              tmem_ld16(tbd, r);
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
              tmem_ld32(tbd, r);
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
        }
        if (phase == 1) {
          // ---- cross-slot composition (model.py:151-159); the two column halves take alternate slots ----
          if (half == 0 && valid) stash[t] = p.scratch_bg[pt];
          if (C::CS == 2) named_bar_sync(1 + ch * 4 + (wic & 3), 64);  // head results of the partner warp are visible
          float S = 0.f;
#pragma unroll 4
          for (int k = 0; k < K; ++k) S += fmaxf(stash[k * kTileM + t].w, 0.f);
          const float denom = S + 1e-5f;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool want_sum = half == 0 && p.raws != nullptr;
#pragma unroll 2
          for (int k = 0; k < K; ++k) {
            const bool mine = C::CS == 1 || (k & 1) == half;
            if (!mine && !want_sum) continue;
            const float4 rw = stash[k * kTileM + t];
            const float dens = fmaxf(rw.w, 0.f);
            const float m = dens / denom;
            float4 u;
            u.x = rgb_act<PASSES == 1>(rw.x);
            u.y = rgb_act<PASSES == 1>(rw.y);
            u.z = rgb_act<PASSES == 1>(rw.z);
            u.w = dens;
            if (p.noise != nullptr && valid) u.w = dens + p.dens_noise * __ldg(p.noise + (long long)k * p.P + pt);
            const float4 mk = make_float4(u.x * m, u.y * m, u.z * m, u.w * m);
            acc.x += mk.x; acc.y += mk.y; acc.z += mk.z; acc.w += mk.w;
            if (valid && mine) {
              if (p.unmasked) p.unmasked[(long long)k * p.P + pt] = u;
              if (p.masked) p.masked[(long long)k * p.P + pt] = mk;
              if (p.masks) p.masks[(long long)k * p.P + pt] = m;
            }
          }
          if (valid && want_sum) p.raws[pt] = acc;
          if (C::CS == 2) named_bar_sync(1 + ch * 4 + (wic & 3), 64);  // stash may be overwritten by the next tile
        }
      }
    }
    // all MMAs of this phase have been consumed; shared memory may be re-used for the next image
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == kEpiWarps) tmem_dealloc(tmem_base, kTmemCols);
}

static size_t decoder_smem_bytes(int passes, int K, int nch, bool psm) {
  const int npl = passes == 3 ? 2 : 1;
  const int w_region = npl * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  if (psm) return 1024 + ((w_region + 1023) & ~1023) + (size_t)nch * kTileM * 128;
  return 1024 + ((w_region + 127) & ~127) + (size_t)nch * K * kTileM * sizeof(float4);
}

// 3-pass modes: the four-channel PSM variant unless the library is built with -DUORF_FWD_X3_LEGACY (tuning / A-B builds)
#ifdef UORF_FWD_X3_LEGACY
constexpr bool kX3Psm = false;
#else
constexpr bool kX3Psm = true;
#endif
constexpr int kPsmChannels = 4;

// floats of `workspace`: the background sweep's raw outputs [P][4] + (PSM) the per-channel head stash [SMs][4][K][128][4]
size_t decoder_fwd_workspace_floats(long long P, int K, int num_sms) {
  return (size_t)P * 4 + (size_t)num_sms * kPsmChannels * K * kTileM * 4;
}

int launch_decoder_fwd(const uorf_decoder_args* a, int num_sms, int* err_flag, cudaStream_t stream) {
  DecFwdParams p;
  p.coor_bg = a->coor_bg;
  p.coor_fg = a->coor_fg;
  p.fg_slot_stride = a->fg_slot_stride;
  p.noise = a->dens_noise != 0.f ? a->noise : nullptr;
  p.image = static_cast<const unsigned char*>(a->image);
  p.scratch_bg = reinterpret_cast<float4*>(a->workspace);
  p.stash_g = reinterpret_cast<float4*>(a->workspace) + a->P;
  p.raws = reinterpret_cast<float4*>(a->raws);
  p.masked = reinterpret_cast<float4*>(a->masked_raws);
  p.unmasked = reinterpret_cast<float4*>(a->unmasked_raws);
  p.masks = a->masks;
  p.outside = a->outside;
  p.fg_transform = a->fg_transform;
  p.fg_slot_offset = a->fg_slot_offset;
  p.locality_ratio = a->locality_ratio;
  p.dens_noise = a->dens_noise;
  p.P = a->P;
  p.K = a->K;
  p.fixed_locality = a->fixed_locality;
  p.locality = a->locality;
  p.num_tiles = (int)((a->P + kTileM - 1) / kTileM);
  p.err_flag = err_flag;
  const int passes = a->precision >= 3 ? 3 : 1;
  const bool f16 = (a->precision % 2) == 0;
  const bool psm = passes == 3 && kX3Psm;
  // legacy 3-pass form: three tiles per SM when the per-slot stash of three channels fits beside the weight image (K <= 13)
  int nch = passes == 3 ? (psm ? kPsmChannels : 3) : 4;
  if (passes == 3 && !psm && decoder_smem_bytes(passes, a->K, 3, false) > 227 * 1024) nch = 2;
  const size_t smem = decoder_smem_bytes(passes, a->K, nch, psm);
  if (smem > 227 * 1024) return UORF_ERR_UNSUPPORTED;
  int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (grid < 1) return UORF_OK;
  auto launch = [&](auto kern, int threads) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UORF_ERR_DEVICE;
    kern<<<grid, threads, smem, stream>>>(p);
    return UORF_OK;
  };
  int rc;
  if (psm) rc = f16 ? launch(decoder_fwd_kernel<3, true, kPsmChannels, true>, Cfg<3, kPsmChannels, true>::THREADS)
                    : launch(decoder_fwd_kernel<3, false, kPsmChannels, true>, Cfg<3, kPsmChannels, true>::THREADS);
  else if (passes == 3 && nch == 3) rc = f16 ? launch(decoder_fwd_kernel<3, true, 3>, Cfg<3, 3>::THREADS) : launch(decoder_fwd_kernel<3, false, 3>, Cfg<3, 3>::THREADS);
  else if (passes == 3) rc = f16 ? launch(decoder_fwd_kernel<3, true, 2>, Cfg<3, 2>::THREADS) : launch(decoder_fwd_kernel<3, false, 2>, Cfg<3, 2>::THREADS);
  else rc = f16 ? launch(decoder_fwd_kernel<1, true, 4>, Cfg<1, 4>::THREADS) : launch(decoder_fwd_kernel<1, false, 4>, Cfg<1, 4>::THREADS);
  if (rc != UORF_OK) return rc;
  return finish_launch(1);
}

}  // namespace uorf

#ifdef UORF_TIMELINE
extern "C" int uorf_debug_timeline(long long* epi, long long* iss, int n) {
  if (n > 8192) n = 8192;
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(epi, g_tl_epi, n * sizeof(long long)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(iss, g_tl_iss, n * sizeof(long long)) != cudaSuccess) return -1;
  return 0;
}
#endif
