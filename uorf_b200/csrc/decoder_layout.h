// Layout of the prepared decoder weight image ("blob") shared by decoder_prep.cu (writer) and
// decoder_fwd.cu (reader).  One image per MLP (background, foreground); both use the same layout so the
// forward kernel can bulk-copy either into the same shared-memory region.
//
//   plane 0 (bf16 "hi"), plane 1 (bf16 "lo" residual; only when passes == 3)   kPlaneBytes each
//     tile t (t = 0..6) at t*kTileBytes : 64 output rows x 64 K columns, bf16, K-major rows of 128 B,
//                                         128-byte swizzle (16 B chunk c of row n stored at chunk c ^ (n&7))
//        t0 = before.0[:, posenc]   t1 = before.2   t2 = before.4
//        t3 = after.0[:, posenc]    t4 = after.0[:, tmp]    t5 = after.2    t6 = after.4
//     tile 7 at 7*kTileBytes        : 32 head rows x 64 K
//        fg: rows 0..Z/4-1 = f_color.0 . f_after_latent (folded), row 16 = f_after_shape
//        bg: rows 0..3     = b_after.6
//     bias tile at kBiasTileOff     : 64 feature rows x 64 K; COLUMN j is bias vector j, added to the accumulator by one
//        extra MMA whose A operand is a one-hot "selector" K-step (forward kernel):
//        j = 0..3 : before.2, before.4, after.2, after.4 ;  j = 16+2s / 17+2s : the hoisted biases of slot s (before.0 / after.0)
//   params (fp32, kParamFloats) : b_before.2, b_before.4, b_after.2, b_after.4 (64 each, zero padded),
//        head bias [32], f_color.2 weight [3][16], f_color.2 bias [4 (3 used)],
//        before.0[:, 32] and after.0[:, 32] (64 each): the weight column of the LAST positional-encoding element cos(16 z).
//        The 3-pass forward feeds only elements 0..31 (two K-steps) through the tensor core and adds this column with one
//        fp32 FMA per output, which frees 16 TMEM columns per tile - enough for a third tile in flight.
//   slot biases (fp32) : per slot s: bias0[64] = before.0.bias + before.0[:, 33:] . z_s
//                                     bias3[64] = after.0.bias  + after.0[:, 33:33+Z] . z_s
//        (the slot-latent columns of the two wide layers are hoisted out of the per-point GEMM)
#pragma once

namespace uorf {
constexpr int kPosenc = 33;        // 3 + 6*n_freq, n_freq = 5
constexpr int kPosencK = 48;       // padded to a multiple of the MMA K (16)
constexpr int kZPad = 64;          // hidden width the kernels run at (z_dim <= 64, zero padded)
constexpr int kTileBytes = 64 * 128;
constexpr int kHeadTileBytes = 32 * 128;
constexpr int kBiasTileOff = 7 * kTileBytes + kHeadTileBytes;              // 61440
constexpr int kPlaneBytes = kBiasTileOff + kTileBytes;                     // 69632
__host__ __device__ constexpr int bias_id_fixed(int which) { return which; }              // 0..3
__host__ __device__ constexpr int bias_id_slot(int s, int which) { return 16 + 2 * s + which; }   // which: 0 before.0, 1 after.0
constexpr int kParamFloats = 480;
constexpr int kParamBytes = kParamFloats * 4;                 // 1920
constexpr int kSlotBiasFloats = 128;
constexpr int kSlotBiasBytes = kSlotBiasFloats * 4;           // 512
// float offsets inside params
constexpr int kPB1 = 0, kPB2 = 64, kPA1 = 128, kPA2 = 192, kPHeadB = 256, kPWc2 = 288, kPBc2 = 336, kPW0x = 352, kPW3x = 416;
constexpr int kPosencMma3 = 32;    // positional-encoding elements the 3-pass forward sends through the tensor core
constexpr int kHeadShapeRow = 16;

__host__ __device__ inline long long blob_bytes_one(int passes, int nslots) {
  return (long long)(passes == 3 ? 2 : 1) * kPlaneBytes + kParamBytes + (long long)nslots * kSlotBiasBytes;
}
// total image: [bg image][fg image]; both 1024-byte aligned inside the buffer
__host__ __device__ inline long long blob_align(long long x) { return (x + 1023) & ~1023LL; }
__host__ __device__ inline long long blob_fg_offset(int passes) { return blob_align(blob_bytes_one(passes, 1)); }
__host__ __device__ inline long long blob_bytes_total(int passes, int K) {
  return blob_fg_offset(passes) + blob_align(blob_bytes_one(passes, K - 1));
}
}  // namespace uorf
