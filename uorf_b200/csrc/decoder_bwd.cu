// Kernel 5 (decoder part) — fused K-slot decoder BACKWARD on tcgen05 / TMEM.
//
// Gradient of Decoder.forward (reference models/model.py:106-161) w.r.t. every decoder weight and z_slots, given
// dL/d raws (Px4).  The reference's drivers only differentiate through `raws` (masked/unmasked/masks are detached,
// models/uorf_nogan_model.py:172-196).
//
// Per 128-point tile and slot the kernel recomputes the forward activations (they are never stored in HBM), keeps
// them as fp16 [point][feature] tiles in shared memory, and walks the MLP backwards:
//     dZ_l = dX_{l+1} * [X_{l+1} > 0]                       (epilogue threads, thread <-> point)
//     dX_l = dZ_l . W_l                                      tcgen05.mma  M=128  A = dZ tile (K-major),  B = W tile (MN-major view)
//     dW_l += dZ_l^T . X_l   (contraction over the points)   tcgen05.mma  M=64   A = dZ tile (MN-major), B = X tile (MN-major)
// The SAME shared-memory tile serves as K-major and as MN-major operand (rows of 128 B, 128B swizzle), so no transpose
// is ever materialised.  The eight dW accumulators (fp32) stay resident in TMEM for the whole kernel, two per 64 columns
// (M=64 accumulators use 16 of every 32 lanes; the partner sits at lane offset 16).  The per-slot column sums that carry
// d z_slots ride on spare positional-encoding columns (a one-hot slot indicator in columns 33..47 of the P tile), and the
// bias gradients of the other layers come from the same columns: one extra N=16 MMA group dZ_l^T . P[:, 32..47] per layer
// sums dZ_l over the points on the tensor core (the epilogue used to spend 31 shuffles per thread and layer on it).  Tensor-core operands are fp16 (11-bit significand) with
// fp32 accumulation; activation gradients are kept in range by a global power-of-two scale taken from max |d raw|, which a
// small elementwise pre-kernel (the backward of the slot composition) produces together with d raw itself.
//
// Outputs are per-CTA partial sums; decoder_bwd_finalize reduces them in a fixed order (deterministic) and unfolds the
// folded head (f_color.0 . f_after_latent) and the hoisted slot-latent columns into the reference's parameter shapes.
#include <type_traits>

#include "common.cuh"
#include "decoder_layout.h"
#include "../../include/uorf_b200.h"

// -DUORF_TIMELINE (tuning variant, scripts/timeline_bwd.py): CTA 0 stamps clock64 at the hand-off points of the fg sweep
#ifdef UORF_TIMELINE
__device__ long long g_tlb_epi[8192];
__device__ long long g_tlb_iss[8192];
#define TLB_EPI(id) do { if (tl_on && tl_ne < 8192) g_tlb_epi[tl_ne++] = (clock64() << 8) | (id); } while (0)
#define TLB_ISS(id) do { if (tl_on && tl_ni < 8192) g_tlb_iss[tl_ni++] = (clock64() << 8) | (id); } while (0)
#else
#define TLB_EPI(id)
#define TLB_ISS(id)
#endif

namespace uorf {

constexpr int kTileM128 = 128;
// Epilogue warps per TMEM lane quadrant: warps w, w + 4, ... split the 64 feature columns of a point.  The hand-off chain of a
// tile is serial (one tile per SM: its activation tiles fill shared memory), and an epilogue step is one warp's dependent
// instruction stream per point group.  Measured: 4 warps per quadrant (17 warps: 96 registers, the f_color.2 sums spill) 4.17 ms
// against 3.35 ms with 2 - kept as a switch.
#ifndef UORF_BWD_NQ
#define UORF_BWD_NQ 2
#endif
// Forward-recompute steps 1..6 take their A operand (the previous layer's activations) from TMEM (tcgen05.mma TS form: 34.6 instead
// of 52 cycles per K-step) and hand it over without a proxy fence: the activations are still written to their shared-memory tile
// for the backward half (dW operand, ReLU gate), but nothing of the recompute half waits for those stores - the proxy fence of the
// first backward hand-off covers them.
#ifndef UORF_BWD_TS_FWD
#define UORF_BWD_TS_FWD 1
#endif
// The backward steps do the same for dX_l = dZ_l . W_l: dZ goes to TMEM (operand of the TS MMAs, handed over first) AND to its
// shared-memory tile (operand of the dW / bias-gradient MMAs, which need it contraction-major); the proxy fence for the tile and
// a second barrier `bar_b` follow the first hand-off, so the fence leaves the dependent chain.
#ifndef UORF_BWD_TS_BWD
#define UORF_BWD_TS_BWD 1
#endif
constexpr int kBColA = 352;   // TMEM columns 352..383: 16-bit activation operand of the recompute steps (32 columns = 64 features)
// D (+)= A[tmem, 4 K-steps of 8 columns] . B[smem tile]^T under one elect (A columns + 8 j, B descriptor + 2 j per K-step)
__device__ __forceinline__ void mma_group_ts4(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred e, p, t;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 h1, h2, h3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, %4, %4;\n\t"
      "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
      "add.u64 h1, %2, 2;\n\tadd.u64 h2, %2, 4;\n\tadd.u64 h3, %2, 6;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], h1, %3, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], h2, %3, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], h3, %3, t;\n\t"
      "}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// dX groups: A[tmem] . B, B = MN-major view of a weight tile (descriptor + 128 per K-step), D overwritten by the first MMA
__device__ __forceinline__ void mma_group_ts4_mn(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred e, f, t;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 h1, h2, h3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 f, %3, %3;\n\tsetp.eq.b32 t, %3, %3;\n\t"
      "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
      "add.u64 h1, %2, 128;\n\tadd.u64 h2, %2, 256;\n\tadd.u64 h3, %2, 384;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, f;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], h1, %3, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], h2, %3, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], h3, %3, t;\n\t"
      "}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_group_ts2_mn(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred e, f, t;\n\t.reg .b32 a1;\n\t.reg .b64 h1;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 f, %3, %3;\n\tsetp.eq.b32 t, %3, %3;\n\t"
      "add.u32 a1, %1, 8;\n\tadd.u64 h1, %2, 128;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, f;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], h1, %3, t;\n\t"
      "}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
// one mbarrier arrival per epilogue WARP (after __syncwarp) instead of one per thread
#ifndef UORF_BWD_WARP_ARRIVE
#define UORF_BWD_WARP_ARRIVE 1
#endif
constexpr int kBwdNQ = UORF_BWD_NQ;              // 2 or 4
constexpr int kBwdCols = 64 / kBwdNQ;            // feature columns per thread in the hidden-layer steps
constexpr int kBwdEpiWarps = 4 * kBwdNQ;
constexpr int kBwdThreads = (kBwdEpiWarps + 1) * 32;  // + 1 MMA-issuer warp
constexpr int kAccFloats = 64 * 64;
// accumulators: 0: before.0 (x P)  1: before.2  2: before.4  3: after.0 (x P)  4: after.0 (x tmp)  5: after.2  6: after.4  7: head
constexpr int kNumAcc = 8;
// per-CTA partial record (floats): 8 accumulators, bias sums [7][64] (layers 1,2,4,5 -> rows 0..3, head -> row 4), f_color.2 (48 + 3)
constexpr int kPartBias = kNumAcc * kAccFloats;
constexpr int kPartWc2 = kPartBias + 5 * 64;
constexpr int kPartFloats = kPartWc2 + 64;

struct DecBwdParams {
  const float* coor_bg;
  const float* coor_fg;
  long long fg_slot_stride;
  const float* noise;
  const float4* unmasked;   // [K][P] forward output (rgb in [0,1], sigma)
  const float4* grad_raws;  // [P]
  const unsigned char* image;  // fp16 1-pass image of uorf_decoder_prepare
  float4* scratch_draw;     // [K][P] d raw (rgb_raw x3, sigma_raw), written by the composition-backward pre-kernel
  const float* draw_amax;   // max |d raw| (device scalar, bit pattern written with atomicMax)
  float* partials;          // [2 phases][grid][kPartFloats]
  const float* fg_transform;
  float locality_ratio;
  float dens_noise;
  long long P;
  int K;
  int fixed_locality;
  int locality;
  int num_tiles;
  int* err_flag;
};

// TMEM columns
constexpr int kBColD = 0;
__device__ __forceinline__ uint32_t acc_taddr(uint32_t base, int acc) {
  // pairs share columns: (0,3) 48 cols @64 ; (1,2) @112 ; (4,5) @176 ; (6,7) @240 ; second of each pair at lane 16
  const int col = acc == 0 || acc == 3 ? 64 : acc == 1 || acc == 2 ? 112 : acc == 4 || acc == 5 ? 176 : 240;
  const int hi = (acc == 3 || acc == 2 || acc == 5 || acc == 7) ? 16 : 0;
  return base + col + ((uint32_t)hi << 16);
}

// bias-gradient accumulators (M=64 features x N=16: dZ^T . P[:, 32..47], whose one-hot slot columns sum dZ over the points):
// b = 0..3 for layers 1, 2, 4, 5, b = 4 for the head; pairs share 16 columns from 304 on, the second at lane offset 16
__device__ __forceinline__ uint32_t bacc_taddr(uint32_t base, int b) {
  return base + 304 + 16 * (b >> 1) + ((uint32_t)((b & 1) ? 16 : 0) << 16);
}

// MN-major view of a [rows][64 x bf16] swizzled tile (rows = contraction index)
__device__ __forceinline__ uint64_t sdesc_mn(uint32_t smem_addr) {
  return (static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4)) | (static_cast<uint64_t>(1024) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}
// fp16 x fp16 -> fp32 (tcgen05 rejects mixed fp16 / bf16 operands).  Activation gradients are made fp16-safe by a global
// power-of-two loss scale derived from max |d raw| (see grad_scale()).
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
// power of two that maps max |d raw| into (512, 1024]
__device__ __forceinline__ float grad_scale(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);   // amax = m * 2^e, m in [0.5, 1)
  return ldexpf(1.f, 10 - e);
}

// Each point's 64 feature columns are split between two threads (`half` = 0 / 1 owns columns 32*half .. 32*half+31).
// write 32 features (fp32 regs) as 16-bit values (F16: fp16, else bf16) into this thread's half of row `t` of a swizzled tile
template <bool F16, bool RELU = false>
__device__ __forceinline__ void store_half16(unsigned char* tile, int t, int half, const float (&v)[32]) {
  unsigned char* row = tile + (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 q;
    if (RELU) {   // ReLU folded into the conversion (cvt.rn.relu)
      q.x = pack2_relu<F16>(v[8 * c + 0], v[8 * c + 1]);
      q.y = pack2_relu<F16>(v[8 * c + 2], v[8 * c + 3]);
      q.z = pack2_relu<F16>(v[8 * c + 4], v[8 * c + 5]);
      q.w = pack2_relu<F16>(v[8 * c + 6], v[8 * c + 7]);
    } else {
      q.x = pack2<F16>(v[8 * c + 0], v[8 * c + 1]);
      q.y = pack2<F16>(v[8 * c + 2], v[8 * c + 3]);
      q.z = pack2<F16>(v[8 * c + 4], v[8 * c + 5]);
      q.w = pack2<F16>(v[8 * c + 6], v[8 * c + 7]);
    }
    *reinterpret_cast<uint4*>(row + ((((4 * half + c) ^ (t & 7)) & 7) << 4)) = q;
  }
}
// dZ = dX * [X > 0], converted and stored in one go: dX (fp32) is packed to fp16 pairs and ANDed with a per-half mask taken
// from the packed activations X of the same positions (activations are ReLU outputs: "non-zero" == "positive"), then written
// to this thread's half of row `t` of the dZ tile.  Three instructions per pair (pack, compare-to-mask, and).
__device__ __forceinline__ void gate_store_half(const unsigned char* x_tile, unsigned char* dz_tile, int t, int half, const float (&g)[32]) {
  const int roff = (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int coff = (((4 * half + c) ^ (t & 7)) & 7) << 4;
    const uint4 x = *reinterpret_cast<const uint4*>(x_tile + roff + coff);
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w};
    uint32_t q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t m;
      asm("{\n\t.reg .b32 z;\n\tmov.b32 z, 0;\n\tset.ne.u32.f16x2 %0, %1, z;\n\t}" : "=r"(m) : "r"(xw[i]));   // 0xFFFF per non-zero half
      q[i] = pack2<true>(g[8 * c + 2 * i], g[8 * c + 2 * i + 1]) & m;
    }
    *reinterpret_cast<uint4*>(dz_tile + roff + coff) = make_uint4(q[0], q[1], q[2], q[3]);
  }
}

// NC-column slices (NC = 16 or 32) of a point's row: chunk = 16-byte chunk index of the first column within the 128-byte row
template <int NC, bool RELU>
__device__ __forceinline__ void store_cols16(unsigned char* tile, int t, int chunk0, const float (&v)[NC]) {
  unsigned char* row = tile + (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
  for (int c = 0; c < NC / 8; ++c) {
    uint4 q;
    if (RELU) {
      q.x = pack2_relu<true>(v[8 * c + 0], v[8 * c + 1]); q.y = pack2_relu<true>(v[8 * c + 2], v[8 * c + 3]);
      q.z = pack2_relu<true>(v[8 * c + 4], v[8 * c + 5]); q.w = pack2_relu<true>(v[8 * c + 6], v[8 * c + 7]);
    } else {
      q.x = pack2<true>(v[8 * c + 0], v[8 * c + 1]); q.y = pack2<true>(v[8 * c + 2], v[8 * c + 3]);
      q.z = pack2<true>(v[8 * c + 4], v[8 * c + 5]); q.w = pack2<true>(v[8 * c + 6], v[8 * c + 7]);
    }
    *reinterpret_cast<uint4*>(row + ((((chunk0 + c) ^ (t & 7)) & 7) << 4)) = q;
  }
}
// store_cols16 + the same packed words into `taddr` (TMEM operand columns of this thread's slice)
template <int NC>
__device__ __forceinline__ void store_cols16_tmem(unsigned char* tile, int t, int chunk0, const float (&v)[NC], uint32_t taddr) {
  unsigned char* row = tile + (t >> 3) * 1024 + (t & 7) * 128;
  uint32_t w[NC / 2];
#pragma unroll
  for (int c = 0; c < NC / 8; ++c) {
    uint4 q;
    q.x = pack2_relu<true>(v[8 * c + 0], v[8 * c + 1]); q.y = pack2_relu<true>(v[8 * c + 2], v[8 * c + 3]);
    q.z = pack2_relu<true>(v[8 * c + 4], v[8 * c + 5]); q.w = pack2_relu<true>(v[8 * c + 6], v[8 * c + 7]);
    w[4 * c] = q.x; w[4 * c + 1] = q.y; w[4 * c + 2] = q.z; w[4 * c + 3] = q.w;
    *reinterpret_cast<uint4*>(row + ((((chunk0 + c) ^ (t & 7)) & 7) << 4)) = q;
  }
  if constexpr (NC == 32) tmem_st16(taddr, w); else tmem_st8(taddr, w);
}
template <int NC>
__device__ __forceinline__ void gate_load_cols(const unsigned char* x_tile, int t, int chunk0, uint4 (&x)[NC / 8]) {
  const int roff = (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
  for (int c = 0; c < NC / 8; ++c) {
    const float4 f = lds128(smem_u32(x_tile + roff + ((((chunk0 + c) ^ (t & 7)) & 7) << 4)));   // asm volatile: stays before the wait
    x[c] = make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w));
  }
}
template <int NC>
__device__ __forceinline__ void gate_store_cols(const uint4 (&xs)[NC / 8], unsigned char* dz_tile, int t, int chunk0, const float (&g)[NC]) {
  const int roff = (t >> 3) * 1024 + (t & 7) * 128;
#pragma unroll
  for (int c = 0; c < NC / 8; ++c) {
    const int coff = (((chunk0 + c) ^ (t & 7)) & 7) << 4;
    const uint4 x = xs[c];
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w};
    uint32_t q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t m;
      asm("{\n\t.reg .b32 z;\n\tmov.b32 z, 0;\n\tset.ne.u32.f16x2 %0, %1, z;\n\t}" : "=r"(m) : "r"(xw[i]));   // 0xFFFF per non-zero half
      q[i] = pack2<true>(g[8 * c + 2 * i], g[8 * c + 2 * i + 1]) & m;
    }
    *reinterpret_cast<uint4*>(dz_tile + roff + coff) = make_uint4(q[0], q[1], q[2], q[3]);
  }
}
// gate_store_cols + the same packed words into `taddr` (TMEM operand columns of this thread's slice)
template <int NC>
__device__ __forceinline__ void gate_store_cols_tmem(const uint4 (&xs)[NC / 8], unsigned char* dz_tile, int t, int chunk0, const float (&g)[NC],
                                                     uint32_t taddr) {
  const int roff = (t >> 3) * 1024 + (t & 7) * 128;
  uint32_t w[NC / 2];
#pragma unroll
  for (int c = 0; c < NC / 8; ++c) {
    const int coff = (((chunk0 + c) ^ (t & 7)) & 7) << 4;
    const uint32_t xw[4] = {xs[c].x, xs[c].y, xs[c].z, xs[c].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t m;
      asm("{\n\t.reg .b32 z;\n\tmov.b32 z, 0;\n\tset.ne.u32.f16x2 %0, %1, z;\n\t}" : "=r"(m) : "r"(xw[i]));   // 0xFFFF per non-zero half
      w[4 * c + i] = pack2<true>(g[8 * c + 2 * i], g[8 * c + 2 * i + 1]) & m;
    }
    *reinterpret_cast<uint4*>(dz_tile + roff + coff) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  }
  if constexpr (NC == 32) tmem_st16(taddr, w); else tmem_st8(taddr, w);
}
// store_half16 (fp16, no ReLU) + the same 16 packed words into TMEM
__device__ __forceinline__ void store_half16_tmem(unsigned char* tile, int t, int half, const float (&v)[32], uint32_t taddr) {
  unsigned char* row = tile + (t >> 3) * 1024 + (t & 7) * 128;
  uint32_t w[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[4 * c + i] = pack2<true>(v[8 * c + 2 * i], v[8 * c + 2 * i + 1]);
    *reinterpret_cast<uint4*>(row + ((((4 * half + c) ^ (t & 7)) & 7) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  }
  tmem_st16(taddr, w);
}
template <int NC>
__device__ __forceinline__ void load_dcols(uint32_t taddr, float (&v)[NC]) {
  uint32_t r[NC];
  if constexpr (NC == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
  tc_wait_ld();
#pragma unroll
  for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void load_d32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  tc_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kBwdThreads, 1) decoder_bwd_kernel(const DecBwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  // bar_b[2]: dZ tile of a backward step fenced for the dW MMAs, ALTERNATING between two barriers (step j of a slot uses j & 1).
  // The epilogue's arrival for step j + 1 only needs the commit of step j, which the issuer makes BEFORE it waits on bar_b for step
  // j: with one barrier a stalled issuer could find two phases completed (parity back where it started) and wait for a third that
  // depends on itself - seen as a bounded-wait trap about once in 2,000 launches.  Step j + 2 needs the commit of step j + 1, which
  // follows that wait in program order, and between the last step of a slot and the first of the next (both barrier 0) lie the
  // seven hand-offs of the forward recompute: with two barriers no phase can be skipped.
  __shared__ __align__(8) uint64_t bar_w, bar_a, bar_d, bar_b[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_red[4][64];

  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K;
  // shared-memory map: [weights plane 61440][params + slot biases -> padded to 1024][X tiles 7 x 16384][dZ 2 x 16384]
  const int w_region = kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  const int x_off = (w_region + 1023) & ~1023;
  unsigned char* s_x = smem + x_off;               // tile i: P (i=0), X1..X6
  unsigned char* s_dz = s_x + 7 * 16384;           // two buffers
  const float* s_params = reinterpret_cast<const float*>(smem + kPlaneBytes);
  const float* s_slotb = s_params + kParamFloats;

  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_a, UORF_BWD_WARP_ARRIVE ? kBwdEpiWarps : kTileM128 * kBwdNQ);
    mbar_init(&bar_d, 1);
    mbar_init(&bar_b[0], UORF_BWD_WARP_ARRIVE ? kBwdEpiWarps : kTileM128 * kBwdNQ);
    mbar_init(&bar_b[1], UORF_BWD_WARP_ARRIVE ? kBwdEpiWarps : kTileM128 * kBwdNQ);
    fence_mbar_init();
  }
  if (warp == kBwdEpiWarps) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  float T[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      T[r * 4 + c] = p.fixed_locality ? __ldg(p.fg_transform + r * 4 + c) : (c < 3 ? __ldg(p.fg_transform + r * 3 + c) : 0.f);
  }

  const float gscale = grad_scale(__ldg(p.draw_amax));
  uint32_t n_a = 0, n_d = 0;   // completions consumed so far (parity = count & 1); they run on across the two phases
  uint32_t n_b = 0;            // issuer: slots finished (bar_b[1] completes three phases per slot)
  for (int phase = 0; phase < 2; ++phase) {
    const int nslots = phase == 0 ? 1 : K - 1;
    if (warp == kBwdEpiWarps && lane == 0) {
      fence_proxy_async_smem();
      const unsigned char* src = p.image + (phase ? blob_fg_offset(1) : 0);
      const uint32_t bytes = (uint32_t)blob_bytes_one(1, nslots);
      mbar_arrive_expect_tx(&bar_w, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
        bulk_g2s(smem + off, src + off, n, &bar_w);
      }
    }
    mbar_wait(&bar_w, phase & 1, p.err_flag, 100 + phase);
    __syncthreads();

    float wc2_acc[51];  // f_color.2 weight (3x16) + bias (3) gradient, per thread
#pragma unroll
    for (int i = 0; i < 51; ++i) wc2_acc[i] = 0.f;

    if (warp == kBwdEpiWarps) {
      // =========================== MMA issuer ===========================
      const uint32_t tbu = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t bar_a_addr = smem_u32(&bar_a);
      const uint32_t w_s = smem_u32(smem);
      const uint32_t x_s = smem_u32(s_x);
      const uint32_t dz_s = smem_u32(s_dz);
      const uint32_t id_fwd64 = idesc_f16(128, 64, false, false);          // X . W^T
      const uint32_t id_fwd_head = idesc_f16(128, phase == 0 ? 16 : 32, false, false);
      const uint32_t id_dx = idesc_f16(128, 64, false, true);              // dZ . W
      const uint32_t id_dw64 = idesc_f16(64, 64, true, true);              // dZ^T . X
      const uint32_t id_dw48 = idesc_f16(64, 48, true, true);
      const uint32_t id_dwb = idesc_f16(64, 16, true, true);
      uint32_t acc_live = 0;  // bit a set once accumulator a holds data
#ifdef UORF_TIMELINE
      const bool tl_on = blockIdx.x == 0 && lane == 0 && phase == 1;
      int tl_ni = 0;
#endif
      auto wait_a = [&](int site) {
        TLB_ISS(0x20);                       // previous step fully issued (commit + trailing dW group)
        mbar_wait_lean(bar_a_addr, n_a & 1, p.err_flag, site);
        ++n_a;
        tc_fence_after();
        TLB_ISS(0x10);
      };
      // forward GEMM, both operands K-major: D (+)= X[tile xi] . W[tile wi]^T
      // every group of K-steps is one asm block under a single elect (common.cuh: mma_group_ss); tile indices and K-step
      // counts are compile-time, so the issuer's path between wake-up and commit is short straight-line code
      auto fwd = [&](auto XI, auto WI, auto NKS, uint32_t idesc, bool first) {
        constexpr int xi = decltype(XI)::value, wi = decltype(WI)::value, nk = decltype(NKS)::value;
        if constexpr (UORF_BWD_TS_FWD && xi >= 1) {   // activations of the previous layer: TMEM operand
          static_assert(nk == 4, "hidden layers have four K-steps");
          mma_group_ts4(tbu + kBColD, tbu + kBColA, make_sdesc_sw128(w_s + wi * kTileBytes), idesc, first ? 0u : 1u);
        } else {
          mma_group_ss<nk, 2, 2>(tbu + kBColD, make_sdesc_sw128(x_s + xi * 16384), make_sdesc_sw128(w_s + wi * kTileBytes), idesc, first ? 0u : 1u);
        }
      };
      // dX = dZ[buf] . W[tile wi] : A K-major (K = out features), B = MN-major view of the weight tile
      const uint32_t bar_b_addr = smem_u32(&bar_b[0]);
      // the dZ tile of backward step J (0..6 within the slot) is visible to the tensor core (only with UORF_BWD_TS_BWD).  Step J uses
      // barrier J & 1: barrier 0 completes four phases per slot (parity known at compile time), barrier 1 three (parity follows the
      // slot count) - static addresses, so the issuer's code stays on the uniform datapath
      auto wait_b = [&](auto J) {
        if (UORF_BWD_TS_BWD) {
          constexpr int j = decltype(J)::value;
          const uint32_t par = (j & 1) ? (((j >> 1) + n_b) & 1u) : (uint32_t)((j >> 1) & 1);
          mbar_wait_lean(bar_b_addr + 8 * (j & 1), par, p.err_flag, 230 + j);
          tc_fence_after();
        }
      };
      auto dx = [&](auto BUF, auto WI, auto NKS) {
        constexpr int buf = decltype(BUF)::value, wi = decltype(WI)::value, nk = decltype(NKS)::value;
        if constexpr (UORF_BWD_TS_BWD) {
          if constexpr (nk == 4) mma_group_ts4_mn(tbu + kBColD, tbu + kBColA, sdesc_mn(w_s + wi * kTileBytes), id_dx);
          else mma_group_ts2_mn(tbu + kBColD, tbu + kBColA, sdesc_mn(w_s + wi * kTileBytes), id_dx);
        } else {
          mma_group_ss<nk, 2, 128>(tbu + kBColD, make_sdesc_sw128(dz_s + buf * 16384), sdesc_mn(w_s + wi * kTileBytes), id_dx, 0u);
        }
      };
      // dW[acc] += dZ[buf]^T . X[tile xi] : both MN-major views, contraction over the 128 points (8 K-steps)
      auto dw = [&](auto ACC, auto BUF, auto XI, uint32_t idesc) {
        constexpr int acc = decltype(ACC)::value, buf = decltype(BUF)::value, xi = decltype(XI)::value;
        const bool live = (acc_live >> acc) & 1u;
        mma_group_ss<8, 128, 128>(acc_taddr(tbu, acc), sdesc_mn(dz_s + buf * 16384), sdesc_mn(x_s + xi * 16384), idesc, live ? 1u : 0u);
        acc_live |= 1u << acc;
      };
      // bias gradient of the layer whose dZ sits in buffer `buf`: dZ^T . (columns 32..47 of the encoding tile), N = 16.
      // Columns 33+s are the one-hot slot indicator, so their sum over s is the column sum of dZ over all valid points -
      // the tensor core does the reduction the epilogue used to do with 31 shuffles per thread.
      uint32_t bacc_live = 0;
      auto dwb = [&](auto B, auto BUF) {
        constexpr int b = decltype(B)::value, buf = decltype(BUF)::value;
        const bool live = (bacc_live >> b) & 1u;
        mma_group_ss<8, 128, 128>(bacc_taddr(tbu, b), sdesc_mn(dz_s + buf * 16384), sdesc_mn(x_s) + 4, id_dwb, live ? 1u : 0u);
        bacc_live |= 1u << b;
      };
#define I_(n) std::integral_constant<int, n>{}
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * it;
        if (tile >= p.num_tiles) break;
        for (int s = 0; s < nslots; ++s) {
          // ---- forward recompute ----
          wait_a(200); fwd(I_(0), I_(0), I_(3), id_fwd64, true); tc_commit_elect(&bar_d);
          wait_a(201); fwd(I_(1), I_(1), I_(4), id_fwd64, true); tc_commit_elect(&bar_d);
          wait_a(202); fwd(I_(2), I_(2), I_(4), id_fwd64, true); tc_commit_elect(&bar_d);
          wait_a(203); fwd(I_(0), I_(3), I_(3), id_fwd64, true); fwd(I_(3), I_(4), I_(4), id_fwd64, false); tc_commit_elect(&bar_d);
          wait_a(204); fwd(I_(4), I_(5), I_(4), id_fwd64, true); tc_commit_elect(&bar_d);
          wait_a(205); fwd(I_(5), I_(6), I_(4), id_fwd64, true); tc_commit_elect(&bar_d);
          wait_a(206); fwd(I_(6), I_(7), I_(4), id_fwd_head, true); tc_commit_elect(&bar_d);
          // ---- backward ----   (dZ buffers alternate: head -> 0, layer 5 -> 1, 4 -> 0, 3 -> 1, 2 -> 0, 1 -> 1, 0 -> 0)
          wait_a(207); dx(I_(0), I_(7), I_(2)); tc_commit_elect(&bar_d); wait_b(I_(0)); dw(I_(7), I_(0), I_(6), id_dw64); dwb(I_(4), I_(0));
          wait_a(208); dx(I_(1), I_(6), I_(4)); tc_commit_elect(&bar_d); wait_b(I_(1)); dw(I_(6), I_(1), I_(5), id_dw64); dwb(I_(3), I_(1));
          wait_a(209); dx(I_(0), I_(5), I_(4)); tc_commit_elect(&bar_d); wait_b(I_(2)); dw(I_(5), I_(0), I_(4), id_dw64); dwb(I_(2), I_(0));
          wait_a(210); dx(I_(1), I_(4), I_(4)); tc_commit_elect(&bar_d); wait_b(I_(3)); dw(I_(4), I_(1), I_(3), id_dw64); dw(I_(3), I_(1), I_(0), id_dw48);
          wait_a(211); dx(I_(0), I_(2), I_(4)); tc_commit_elect(&bar_d); wait_b(I_(4)); dw(I_(2), I_(0), I_(2), id_dw64); dwb(I_(1), I_(0));
          wait_a(212); dx(I_(1), I_(1), I_(4)); tc_commit_elect(&bar_d); wait_b(I_(5)); dw(I_(1), I_(1), I_(1), id_dw64); dwb(I_(0), I_(1));
          wait_a(213); wait_b(I_(6)); dw(I_(0), I_(0), I_(0), id_dw48); tc_commit_elect(&bar_d);   // all tiles free for the next slot once this lands
          ++n_b;                                                              // slots done (parity of barrier 1)
        }
      }
      __syncwarp();
    } else {
      // =========================== epilogue warps (two threads <-> point) ===========================
      const int t = (warp & 3) * 32 + lane;   // point within the tile == TMEM lane
      const int quarter = warp >> 2;          // hidden-layer steps: this thread owns feature columns kBwdCols * quarter ...
      const int c0 = kBwdCols * quarter;
      const int chunk0 = c0 / 8;
      // encoding and head steps keep the two-way column split; with four warps per quadrant the last two only take part in the hand-off
      const int half = quarter;
      const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#ifdef UORF_TIMELINE
      const bool tl_on = blockIdx.x == 0 && warp == 0 && lane == 0 && phase == 1;
      int tl_ne = 0;
#endif
      // barrier / parameter addresses as opaque 32-bit shared addresses (the compiler otherwise re-derives the shared window
      // from special registers at every use) and a single-loop wait: the general mbar_wait() and the select chain that picked
      // the bias pointer cost ~300 of the ~1100 cycles of every hand-off
      const uint32_t bar_d_addr = opaque(smem_u32(&bar_d)), bar_a_addr = opaque(smem_u32(&bar_a));
      const uint32_t prma = opaque(smem_u32(s_params));
      auto wait_d = [&](int site) {
        mbar_wait_lean(bar_d_addr, n_d & 1, p.err_flag, site);
        ++n_d;
        tc_fence_after();
        TLB_EPI(0x30);
      };
      auto publish = [&]() {  // make this thread's shared-memory tile writes visible to the tensor core, then signal
        TLB_EPI(0x50);
        fence_proxy_async_smem();
        tc_fence_before();
#if UORF_BWD_WARP_ARRIVE
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(bar_a_addr);
#else
        mbar_arrive_addr(bar_a_addr);
#endif
        TLB_EPI(0x40);
      };
      auto publish_tmem = [&]() {  // hand-off of a TMEM operand: the tcgen05.st has to be complete, no proxy fence
        TLB_EPI(0x50);
        tc_wait_st();
        tc_fence_before();
#if UORF_BWD_WARP_ARRIVE
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(bar_a_addr);
#else
        mbar_arrive_addr(bar_a_addr);
#endif
        TLB_EPI(0x40);
      };
      const uint32_t bar_b_addr = opaque(smem_u32(&bar_b[0]));
      auto publish_tile = [&](int j) {  // second hand-off of backward step j (0..6): the shared-memory dZ tile is visible to the tensor core
        fence_proxy_async_smem();
#if UORF_BWD_WARP_ARRIVE
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(bar_b_addr + 8 * (j & 1));
#else
        mbar_arrive_addr(bar_b_addr + 8 * (j & 1));
#endif
      };
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * it;
        if (tile >= p.num_tiles) break;
        const long long pt = tile * kTileM128 + t;
        const bool valid = pt < p.P;
        // fg slots that share their coordinates (the reference's expand() view) share the encoding: the first column half of
        // the P tile is written once per tile, only the one-hot slot indicator in the second half changes from slot to slot
        const bool shared_enc = phase == 1 && p.fg_slot_stride == 0;
        float x = 0.f, y = 0.f, z = 0.f;
        bool outside = false;
        float cos16z = 0.f;
        for (int s = 0; s < nslots; ++s) {
          const int kslot = phase == 0 ? 0 : s + 1;
          const bool fresh = s == 0 || !shared_enc;
          // d raw of this slot is needed seven hand-offs from now: fetch it here, off the critical path
          float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && half == 0) dr = p.scratch_draw[(long long)kslot * p.P + pt];
          // ---- stage 0: coordinates -> positional encoding tile (column 33 + s carries the per-slot column sums) ----
          if (fresh) {
          x = y = z = 0.f;
          outside = false;
          if (valid) {
            const float* src = phase == 0 ? p.coor_bg + pt * 3 : p.coor_fg + (long long)s * p.fg_slot_stride + pt * 3;
            x = __ldg(src); y = __ldg(src + 1); z = __ldg(src + 2);
          }
          if (phase == 1) {
            if (p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
            const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[3]);
            const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z)), T[7]);
            const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z)), T[11]);
            x = tx; y = ty; z = tz;
            if (!p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
          }
          }
          // wait until the previous slot's last weight-gradient MMAs have finished reading the tiles
          if (it > 0 || s > 0) wait_d(300);
          if (half == 0) {
            if (fresh) {
            // columns 0..31: x y z, then [sin, cos](2^i xyz) up to sin(16 xyz) and cos(16 x), cos(16 y)
            float v[32];
            float sn[3], cs[3];
            sincosf(x, &sn[0], &cs[0]);
            sincosf(y, &sn[1], &cs[1]);
            sincosf(z, &sn[2], &cs[2]);
            v[0] = x; v[1] = y; v[2] = z;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                v[3 + 6 * i + j] = sn[j];
                if (6 + 6 * i + j < 32) v[6 + 6 * i + j] = cs[j];
              }
              if (i < 4) {
#pragma unroll
                for (int j = 0; j < 3; ++j) { const float s2 = 2.f * sn[j] * cs[j], c2 = 1.f - 2.f * sn[j] * sn[j]; sn[j] = s2; cs[j] = c2; }
              }
            }
            if (!valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;   // padded points contribute nothing to any gradient
            }
            store_half16<true>(s_x, t, 0, v);
            }
          } else if (half == 1) {
            // columns 32..63: cos(16 z), the one-hot slot indicator, zero padding
            float v[32];
            if (fresh) {
              float sn, cs;
              sincosf(z, &sn, &cs);
#pragma unroll
              for (int i = 0; i < 4; ++i) { const float s2 = 2.f * sn * cs, c2 = 1.f - 2.f * sn * sn; sn = s2; cs = c2; }
              cos16z = cs;
            }
            v[0] = valid ? cos16z : 0.f;
#pragma unroll
            for (int i = 1; i < 32; ++i) v[i] = (valid && i == 1 + (s < 15 ? s : 14)) ? 1.f : 0.f;
            store_half16<true>(s_x, t, 1, v);
          }
          publish();
          // ---- stages 1..6: hidden activations ----
          const uint32_t sba = prma + 4 * (kParamFloats + s * kSlotBiasFloats);
#pragma unroll 1
          for (int st = 1; st <= 6; ++st) {
            // bias of layer st - 1: slot biases for the two wide layers (st 1, 4), else b_before.2, .4, b_after.2, .4
            static_assert(kPB1 == 0 && kPB2 == 64 && kPA1 == 128 && kPA2 == 192, "parameter block layout");
            const uint32_t bias = (st == 1 ? sba : st == 4 ? sba + 256 : prma + 256 * (st - 2 - (st > 4 ? 1 : 0))) + 4 * c0;
            float4 b[kBwdCols / 4];               // fetched before the wait: no shared-memory latency behind it
#pragma unroll
            for (int q = 0; q < kBwdCols / 4; ++q) b[q] = lds128(bias + 16 * q);
            wait_d(300 + st);
            float v[kBwdCols];
            load_dcols<kBwdCols>(tb + kBColD + c0, v);
            // rows of padded points need no masking here: their d raw is zero, so every dZ row derived from it is zero and
            // nothing they hold reaches a gradient
#pragma unroll
            for (int q = 0; q < kBwdCols / 4; ++q) {
              v[4 * q + 0] += b[q].x; v[4 * q + 1] += b[q].y; v[4 * q + 2] += b[q].z; v[4 * q + 3] += b[q].w;
            }
            if (UORF_BWD_TS_FWD) {
              store_cols16_tmem<kBwdCols>(s_x + st * 16384, t, chunk0, v, tb + kBColA + c0 / 2);
              publish_tmem();
            } else {
              store_cols16<kBwdCols, true>(s_x + st * 16384, t, chunk0, v);
              publish();
            }
          }
          // ---- head forward + backward seed (the 17 head rows live in the first column half) ----
          wait_d(307);
          {
            float dzh[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) dzh[i] = 0.f;
            if (half == 0) {
              uint32_t r[32];
              tmem_ld32(tb + kBColD, r);
              tc_wait_ld();
              dr.x *= gscale; dr.y *= gscale; dr.z *= gscale; dr.w *= gscale;
              const float* hb = s_params + kPHeadB;
              if (phase == 0) {
                dzh[0] = dr.x; dzh[1] = dr.y; dzh[2] = dr.z; dzh[3] = dr.w;
              } else {
                const float* wc2 = s_params + kPWc2;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float pre = __uint_as_float(r[i]) + hb[i];
                  const float h = fmaxf(pre, 0.f);
                  wc2_acc[i] += dr.x * h;
                  wc2_acc[16 + i] += dr.y * h;
                  wc2_acc[32 + i] += dr.z * h;
                  const float dh = wc2[i] * dr.x + wc2[16 + i] * dr.y + wc2[32 + i] * dr.z;
                  dzh[i] = pre > 0.f ? dh : 0.f;
                }
                wc2_acc[48] += dr.x; wc2_acc[49] += dr.y; wc2_acc[50] += dr.z;
                dzh[kHeadShapeRow] = (p.locality && outside) ? 0.f : dr.w;   // model.py:147-148 zeroes sigma outside the box
              }
            }
            if (UORF_BWD_TS_BWD) {
              if (half < 2) store_half16_tmem(s_dz, t, half, dzh, tb + kBColA + 16 * half);
            } else {
              if (half < 2) store_half16<true>(s_dz, t, half, dzh);   // the second half writes the zero padding of the head's dZ tile
            }
          }
          if (UORF_BWD_TS_BWD) { publish_tmem(); publish_tile(0); } else publish();
          // ---- layers 5..0 ----
#pragma unroll 1
          for (int l = 5; l >= 0; --l) {
            uint4 xg[kBwdCols / 8];                         // activations that gate this layer: fetched before the wait
            gate_load_cols<kBwdCols>(s_x + (l + 1) * 16384, t, chunk0, xg);
            wait_d(320 + l);
            float g[kBwdCols];
            load_dcols<kBwdCols>(tb + kBColD + c0, g);      // dX_{l+1}
            if (UORF_BWD_TS_BWD) {
              gate_store_cols_tmem<kBwdCols>(xg, s_dz + ((l & 1) ? 16384 : 0), t, chunk0, g, tb + kBColA + c0 / 2);
              publish_tmem();
              publish_tile(6 - l);
            } else {
              gate_store_cols<kBwdCols>(xg, s_dz + ((l & 1) ? 16384 : 0), t, chunk0, g);   // dZ_l = dX_{l+1} * [X_{l+1} > 0]
              publish();
            }
          }
        }
      }
    }
    // ---- end of phase: drain, dump accumulators + bias sums as this CTA's partial record ----
    if (warp < kBwdEpiWarps) {
      // the issuer's last commit (dW of layer 0) has not been consumed yet
      bool any = (long long)blockIdx.x < p.num_tiles;
      if (any) {
        mbar_wait(&bar_d, n_d & 1, p.err_flag, 399);
        ++n_d;
        tc_fence_after();
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    float* part = p.partials + ((long long)phase * gridDim.x + blockIdx.x) * kPartFloats;
    if (warp < 4) {
      const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int row = warp * 16 + (lane & 15);
      const int second = lane >> 4;
#pragma unroll 1
      for (int pair = 0; pair < 4; ++pair) {
        const int col = pair == 0 ? 64 : pair == 1 ? 112 : pair == 2 ? 176 : 240;
        const int acc = pair == 0 ? (second ? 3 : 0) : pair == 1 ? (second ? 2 : 1) : pair == 2 ? (second ? 5 : 4) : (second ? 7 : 6);
        const int ncol = pair == 0 ? 48 : 64;
        for (int c0 = 0; c0 < ncol; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tb + col + c0, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) part[acc * kAccFloats + row * 64 + c0 + i] = __uint_as_float(r[i]);
        }
        if (ncol == 48)
          for (int c = 48; c < 64; ++c) part[acc * kAccFloats + row * 64 + c] = 0.f;
      }
      // bias gradients: sum of the one-hot slot columns (local columns 1..15) of the N=16 accumulators
#pragma unroll 1
      for (int pb = 0; pb < 3; ++pb) {
        uint32_t r[16];
        tmem_ld16(tb + 304 + 16 * pb, r);
        tc_wait_ld();
        const int b = 2 * pb + second;
        float sum = 0.f;
#pragma unroll
        for (int i = 1; i < 16; ++i) sum += __uint_as_float(r[i]);
        if (b < 5) part[kPartBias + b * 64 + row] = sum;
      }
      // f_color.2 gradient: block reduction of the per-thread sums (fixed order)
#pragma unroll
      for (int base = 0; base < 51; base += 1) {
        float v = wc2_acc[base];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][base] = v;
      }
    }
    __syncthreads();
    if (threadIdx.x < 64) part[kPartWc2 + threadIdx.x] = threadIdx.x < 51 ? s_red[0][threadIdx.x] + s_red[1][threadIdx.x] + s_red[2][threadIdx.x] + s_red[3][threadIdx.x] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == kBwdEpiWarps) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// backward of the slot composition (model.py:151-159), elementwise, one thread per point:
//   d raw_k = (d rgb_raw x3, d sigma_raw) for every slot, and max |d raw| for the loss scale
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decoder_compose_bwd_kernel(const float4* __restrict__ grad_raws, const float4* __restrict__ unmasked,
                                                                  const float* __restrict__ noise, float dens_noise, long long P, int K,
                                                                  float4* __restrict__ draw, float* __restrict__ amax) {
  const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float local = 0.f;
  if (pt < P) {
    const float4 g = __ldg(grad_raws + pt);
    float S = 0.f, qm = 0.f;
    for (int k = 0; k < K; ++k) {
      const float4 u = __ldg(unmasked + (long long)k * P + pt);
      S += fmaxf(noise ? u.w - dens_noise * __ldg(noise + (long long)k * P + pt) : u.w, 0.f);
    }
    const float denom = S + 1e-5f;
    for (int k = 0; k < K; ++k) {
      const float4 u = __ldg(unmasked + (long long)k * P + pt);
      const float dens = fmaxf(noise ? u.w - dens_noise * __ldg(noise + (long long)k * P + pt) : u.w, 0.f);
      qm += (g.x * u.x + g.y * u.y + g.z * u.z + g.w * u.w) * (dens / denom);
    }
    for (int k = 0; k < K; ++k) {
      const float4 u = __ldg(unmasked + (long long)k * P + pt);
      const float dens = fmaxf(noise ? u.w - dens_noise * __ldg(noise + (long long)k * P + pt) : u.w, 0.f);
      const float m = dens / denom;
      const float q = g.x * u.x + g.y * u.y + g.z * u.z + g.w * u.w;
      float4 d;
      // rgb = (tanh(raw) + 1) / 2  ->  d raw = d rgb * (1 - tanh^2) / 2 = d rgb * 2 rgb (1 - rgb)
      d.x = g.x * m * 2.f * u.x * (1.f - u.x);
      d.y = g.y * m * 2.f * u.y * (1.f - u.y);
      d.z = g.z * m * 2.f * u.z * (1.f - u.z);
      // m_k = dens_k / (sum dens + 1e-5), sigma_k = dens_k + noise;  sigma_raw gated by its ReLU
      d.w = dens > 0.f ? g.w * m + (q - qm) / denom : 0.f;
      draw[(long long)k * P + pt] = d;
      local = fmaxf(local, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local = fmaxf(local, __shfl_xor_sync(0xffffffffu, local, o));
  if ((threadIdx.x & 31) == 0 && local > 0.f && isfinite(local))
    atomicMax(reinterpret_cast<int*>(amax), __float_as_int(local));   // non-negative floats order like their bit patterns
}

// ---------------------------------------------------------------------------------------------------
// finalize: fixed-order reduction over the CTAs' partial records, then unfold into the reference's parameter shapes
// ---------------------------------------------------------------------------------------------------
struct FinalizeParams {
  uorf_decoder_weights w;       // forward weights (needed to unfold the folded head and the hoisted latent columns)
  uorf_decoder_grads g;         // gradient outputs, same shapes
  const float* z_slots;
  float* grad_z_slots;          // [K][Z]
  const float* partials;
  float* reduced;               // [2][kPartFloats] scratch
  int grid, K, Z;
};

__global__ void decoder_bwd_reduce_kernel(const float* partials, float* reduced, int grid, const float* draw_amax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * kPartFloats) return;
  const int phase = i / kPartFloats, e = i % kPartFloats;
  float acc = 0.f;
  for (int b = 0; b < grid; ++b) acc += partials[((long long)phase * grid + b) * kPartFloats + e];
  reduced[i] = acc / grad_scale(__ldg(draw_amax));   // undo the loss scale (exact: power of two)
}

// Every output element is independent (<= 64x161 outputs per tensor, dot products of <= 64 terms): grid = (CTAs, 2) with
// blockIdx.y = background / foreground and a grid-stride over the elements (as ONE CTA this kernel took 120 us of pure
// load latency).
__global__ void decoder_bwd_unfold_kernel(const FinalizeParams f) {
  const int Z = f.Z, K = f.K, I = kPosenc + Z, CH = Z / 4;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  {
    const int fg = blockIdx.y;
    const float* R = f.reduced + fg * kPartFloats;   // phase 0 = background, 1 = foreground
    const float* A0 = R + 0 * kAccFloats;   // [out][P cols]   before.0
    const float* A1 = R + 1 * kAccFloats;
    const float* A2 = R + 2 * kAccFloats;
    const float* A3 = R + 3 * kAccFloats;   // after.0 x P
    const float* A4 = R + 4 * kAccFloats;   // after.0 x tmp
    const float* A5 = R + 5 * kAccFloats;
    const float* A6 = R + 6 * kAccFloats;
    const float* A7 = R + 7 * kAccFloats;   // head
    const float* DB = R + kPartBias;        // [5][64]: before.2, before.4, after.2, after.4, head
    float* const* gbw = fg ? f.g.f_before_w : f.g.b_before_w;
    float* const* gbb = fg ? f.g.f_before_b : f.g.b_before_b;
    float* const* gaw = fg ? f.g.f_after_w : f.g.b_after_w;
    float* const* gab = fg ? f.g.f_after_b : f.g.b_after_b;
    const float* const* bw = fg ? f.w.f_before_w : f.w.b_before_w;
    const float* const* aw = fg ? f.w.f_after_w : f.w.b_after_w;
    const int nslots = fg ? K - 1 : 1;
    // square layers + biases
    for (int e = tid; e < Z * Z; e += nt) {
      const int n = e / Z, k = e % Z;
      gbw[1][e] = A1[n * 64 + k];
      gbw[2][e] = A2[n * 64 + k];
      gaw[1][e] = A5[n * 64 + k];
      gaw[2][e] = A6[n * 64 + k];
    }
    for (int n = tid; n < Z; n += nt) {
      gbb[1][n] = DB[0 * 64 + n];
      gbb[2][n] = DB[1 * 64 + n];
      gab[1][n] = DB[2 * 64 + n];
      gab[2][n] = DB[3 * 64 + n];
      // wide layers: bias = sum over slots of the per-slot column sums (one-hot columns 33 + s, slot >= 14 share column 47)
      float b0 = 0.f, b3 = 0.f;
      for (int c = kPosenc; c < 48; ++c) { b0 += A0[n * 64 + c]; b3 += A3[n * 64 + c]; }
      gbb[0][n] = b0;
      gab[0][n] = b3;
    }
    // wide layers: posenc columns, tmp columns, latent columns (outer product of per-slot column sums with z)
    for (int e = tid; e < Z * (I + Z); e += nt) {
      const int n = e / (I + Z), c = e % (I + Z);
      float v;
      if (c < kPosenc) v = A3[n * 64 + c];
      else if (c < I) {
        v = 0.f;
        for (int s = 0; s < nslots && s < 15; ++s) v += A3[n * 64 + kPosenc + s] * f.z_slots[(fg ? s + 1 : 0) * Z + (c - kPosenc)];
      } else v = A4[n * 64 + (c - I)];
      gaw[0][e] = v;
    }
    for (int e = tid; e < Z * I; e += nt) {
      const int n = e / I, c = e % I;
      float v;
      if (c < kPosenc) v = A0[n * 64 + c];
      else {
        v = 0.f;
        for (int s = 0; s < nslots && s < 15; ++s) v += A0[n * 64 + kPosenc + s] * f.z_slots[(fg ? s + 1 : 0) * Z + (c - kPosenc)];
      }
      gbw[0][e] = v;
    }
    // d z_slots: W0z^T colsum0_s + W3z^T colsum3_s
    for (int e = tid; e < nslots * Z; e += nt) {
      const int s = e / Z, c = e % Z;
      float v = 0.f;
      if (s < 15) {
        for (int n = 0; n < Z; ++n)
          v += bw[0][n * I + kPosenc + c] * A0[n * 64 + kPosenc + s] + aw[0][n * (I + Z) + kPosenc + c] * A3[n * 64 + kPosenc + s];
      }
      f.grad_z_slots[(fg ? s + 1 : 0) * Z + c] = v;
    }
    // heads
    if (fg) {
      const float* dbf = DB + 4 * 64;  // [0..CH) colour-hidden bias sums, [16] shape bias sum
      const float* WC2 = R + kPartWc2; // [3][16] + [3]
      for (int e = tid; e < 3 * CH; e += nt) f.g.f_color2_w[e] = WC2[(e / CH) * 16 + (e % CH)];
      for (int e = tid; e < 3; e += nt) f.g.f_color2_b[e] = WC2[48 + e];
      for (int e = tid; e < Z; e += nt) f.g.f_after_shape_w[e] = A7[kHeadShapeRow * 64 + e];
      if (tid == 0) f.g.f_after_shape_b[0] = dbf[kHeadShapeRow];
      // folded head: Zh = Wc0 (Wlat x + blat) + bc0 ;  dWf = A7[0..CH) , dbf
      for (int e = tid; e < CH * Z; e += nt) {   // d Wc0 = dWf . Wlat^T + dbf (x) blat
        const int n = e / Z, m = e % Z;
        float v = dbf[n] * f.w.f_after_latent_b[m];
        for (int k = 0; k < Z; ++k) v += A7[n * 64 + k] * f.w.f_after_latent_w[m * Z + k];
        f.g.f_color0_w[e] = v;
      }
      for (int e = tid; e < CH; e += nt) f.g.f_color0_b[e] = dbf[e];
      for (int e = tid; e < Z * Z; e += nt) {    // d Wlat = Wc0^T . dWf
        const int m = e / Z, k = e % Z;
        float v = 0.f;
        for (int n = 0; n < CH; ++n) v += f.w.f_color0_w[n * Z + m] * A7[n * 64 + k];
        f.g.f_after_latent_w[e] = v;
      }
      for (int m = tid; m < Z; m += nt) {        // d blat = Wc0^T dbf
        float v = 0.f;
        for (int n = 0; n < CH; ++n) v += f.w.f_color0_w[n * Z + m] * dbf[n];
        f.g.f_after_latent_b[m] = v;
      }
    } else {
      const float* dbh = DB + 4 * 64;
      for (int e = tid; e < 4 * Z; e += nt) f.g.b_after_w[3][e] = A7[(e / Z) * 64 + (e % Z)];
      for (int e = tid; e < 4; e += nt) f.g.b_after_b[3][e] = dbh[e];
    }
  }
}

size_t decoder_bwd_workspace_floats(long long P, int K, int num_sms) {
  const long long tiles = (P + 127) / 128;
  const long long grid = tiles < num_sms ? tiles : num_sms;
  return (size_t)K * P * 4 + (size_t)2 * grid * kPartFloats + (size_t)2 * kPartFloats + 64;   // d raw | partials | reduced | amax
}

static size_t decoder_bwd_smem(int K) {
  const int w_region = kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  return 1024 + ((w_region + 1023) & ~1023) + 9 * 16384;
}

int launch_decoder_bwd(const uorf_decoder_weights* w, const uorf_decoder_grads* g, const uorf_decoder_bwd_args* a, int num_sms,
                       int* err_flag, cudaStream_t stream) {
  DecBwdParams p;
  p.coor_bg = a->coor_bg;
  p.coor_fg = a->coor_fg;
  p.fg_slot_stride = a->fg_slot_stride;
  p.noise = a->dens_noise != 0.f ? a->noise : nullptr;
  p.unmasked = reinterpret_cast<const float4*>(a->unmasked_raws);
  p.grad_raws = reinterpret_cast<const float4*>(a->grad_raws);
  p.image = static_cast<const unsigned char*>(a->image);
  p.fg_transform = a->fg_transform;
  p.locality_ratio = a->locality_ratio;
  p.dens_noise = a->dens_noise;
  p.P = a->P;
  p.K = a->K;
  p.fixed_locality = a->fixed_locality;
  p.locality = a->locality;
  p.num_tiles = (int)((a->P + 127) / 128);
  p.err_flag = err_flag;
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  float* ws = a->workspace;
  p.scratch_draw = reinterpret_cast<float4*>(ws);
  p.partials = ws + (size_t)a->K * a->P * 4;
  float* reduced = p.partials + (size_t)2 * grid * kPartFloats;
  float* amax = reduced + (size_t)2 * kPartFloats;
  p.draw_amax = amax;
  cudaMemsetAsync(amax, 0, sizeof(float), stream);
  decoder_compose_bwd_kernel<<<(unsigned)((a->P + 255) / 256), 256, 0, stream>>>(p.grad_raws, p.unmasked, p.noise, p.dens_noise, a->P, a->K,
                                                                               p.scratch_draw, amax);
  const size_t smem = decoder_bwd_smem(a->K);
  if (smem > 227 * 1024) return UORF_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(decoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UORF_ERR_DEVICE;
  decoder_bwd_kernel<<<grid, kBwdThreads, smem, stream>>>(p);
  decoder_bwd_reduce_kernel<<<(2 * kPartFloats + 255) / 256, 256, 0, stream>>>(p.partials, reduced, grid, amax);
  FinalizeParams f;
  f.w = *w;
  f.g = *g;
  f.z_slots = a->z_slots;
  f.grad_z_slots = a->grad_z_slots;
  f.partials = p.partials;
  f.reduced = reduced;
  f.grid = grid;
  f.K = a->K;
  f.Z = a->Z;
  decoder_bwd_unfold_kernel<<<dim3(32, 2), 256, 0, stream>>>(f);
  return finish_launch(4);
}

}  // namespace uorf

#ifdef UORF_TIMELINE
extern "C" int uorf_debug_timeline_bwd(long long* epi, long long* iss, int n) {
  if (n > 8192) n = 8192;
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(epi, g_tlb_epi, n * sizeof(long long)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(iss, g_tlb_iss, n * sizeof(long long)) != cudaSuccess) return -1;
  return 0;
}
#endif
