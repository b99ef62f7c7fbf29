// Decoder weight preparation: fp32 state-dict tensors -> tensor-core image (decoder_layout.h).
// Runs once per scene (weights may have moved after an optimizer step, slots are new every scene).
// Reference: Decoder.__init__ / forward, models/model.py:65-104,130-146 (which columns feed which layer).
#include "common.cuh"
#include "decoder_layout.h"
#include "../../include/uorf_b200.h"

namespace uorf {

struct PrepArgs {
  uorf_decoder_weights w;
  const float* z_slots;
  int K, Z, passes, f16;
  unsigned char* image;
};

// byte offset of element (row n, col k) inside a swizzled K-major tile
__device__ __forceinline__ int sw128_off(int n, int k) {
  return (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2;
}

__device__ __forceinline__ void put(unsigned char* plane0, int passes, int f16, int off, float v) {
  const unsigned short hi = f16 ? to16<true>(v) : to16<false>(v);
  *reinterpret_cast<unsigned short*>(plane0 + off) = hi;
  if (passes == 3) {
    const float r = v - (f16 ? from16<true>(hi) : from16<false>(hi));
    *reinterpret_cast<unsigned short*>(plane0 + kPlaneBytes + off) = f16 ? to16<true>(r) : to16<false>(r);
  }
}

// grid: x = 0..7 tile, y = 0 (bg) / 1 (fg);   x = 8: params;   x = 9: bias tile;   x = 10..: slot biases;
//       z = 0..kPrepSplit-1: interleaved slices of a block's element range (the work is a handful of dependent global loads per
//       element: spreading it over more CTAs shortens the serial chain, the kernel is latency-bound)
constexpr int kPrepSplit = 4;

__global__ void decoder_prep_kernel(PrepArgs a) {
  const int e_first = blockIdx.z * blockDim.x + threadIdx.x, e_step = kPrepSplit * blockDim.x;
  const int fg = blockIdx.y;
  const int Z = a.Z;
  const int I = kPosenc + Z;
  unsigned char* img = a.image + (fg ? blob_fg_offset(a.passes) : 0);
  const int nplanes = a.passes == 3 ? 2 : 1;
  const float* const* before_w = fg ? a.w.f_before_w : a.w.b_before_w;
  const float* const* before_b = fg ? a.w.f_before_b : a.w.b_before_b;
  const float* const* after_w = fg ? a.w.f_after_w : a.w.b_after_w;
  const float* const* after_b = fg ? a.w.f_after_b : a.w.b_after_b;
  const int bx = blockIdx.x;

  if (bx < 7) {
    unsigned char* tile = img + bx * kTileBytes;
    for (int e = e_first; e < 64 * 64; e += e_step) {
      const int n = e >> 6, k = e & 63;
      float v = 0.f;
      if (n < Z) {
        switch (bx) {
          case 0: if (k < kPosenc) v = before_w[0][n * I + k]; break;
          case 1: if (k < Z) v = before_w[1][n * Z + k]; break;
          case 2: if (k < Z) v = before_w[2][n * Z + k]; break;
          case 3: if (k < kPosenc) v = after_w[0][n * (I + Z) + k]; break;
          case 4: if (k < Z) v = after_w[0][n * (I + Z) + I + k]; break;
          case 5: if (k < Z) v = after_w[1][n * Z + k]; break;
          case 6: if (k < Z) v = after_w[2][n * Z + k]; break;
        }
      }
      put(tile, a.passes, a.f16, sw128_off(n, k), v);
    }
  } else if (bx == 7) {
    unsigned char* tile = img + 7 * kTileBytes;
    const int CH = Z / 4;
    for (int e = e_first; e < 32 * 64; e += e_step) {
      const int n = e >> 6, k = e & 63;
      float v = 0.f;
      if (k < Z) {
        if (fg) {
          if (n < CH) {  // f_color.0 . f_after_latent
            double acc = 0.0;
            for (int m = 0; m < Z; ++m) acc += (double)a.w.f_color0_w[n * Z + m] * (double)a.w.f_after_latent_w[m * Z + k];
            v = (float)acc;
          } else if (n == kHeadShapeRow) {
            v = a.w.f_after_shape_w[k];
          }
        } else if (n < 4) {
          v = a.w.b_after_w[3][n * Z + k];
        }
      }
      put(tile, a.passes, a.f16, sw128_off(n, k), v);
    }
  } else if (bx == 8) {
    float* prm = reinterpret_cast<float*>(img + nplanes * kPlaneBytes);
    const int CH = Z / 4;
    for (int e = e_first; e < kParamFloats; e += e_step) {
      float v = 0.f;
      if (e < kPHeadB) {
        const int which = e >> 6, n = e & 63;
        if (n < Z) v = which == 0 ? before_b[1][n] : which == 1 ? before_b[2][n] : which == 2 ? after_b[1][n] : after_b[2][n];
      } else if (e < kPWc2) {
        const int n = e - kPHeadB;
        if (fg) {
          if (n < CH) {
            double acc = a.w.f_color0_b[n];
            for (int m = 0; m < Z; ++m) acc += (double)a.w.f_color0_w[n * Z + m] * (double)a.w.f_after_latent_b[m];
            v = (float)acc;
          } else if (n == kHeadShapeRow) {
            v = a.w.f_after_shape_b[0];
          }
        } else if (n < 4) {
          v = a.w.b_after_b[3][n];
        }
      } else if (e < kPBc2) {
        const int c = (e - kPWc2) / 16, i = (e - kPWc2) % 16;
        if (fg && i < CH) v = a.w.f_color2_w[c * CH + i];
      } else if (e < kPBc2 + 3) {
        if (fg) v = a.w.f_color2_b[e - kPBc2];
      } else if (e >= kPW0x) {
        const int which = (e - kPW0x) >> 6, n = (e - kPW0x) & 63;
        if (n < Z) v = which == 0 ? before_w[0][n * I + kPosencMma3] : after_w[0][n * (I + Z) + kPosencMma3];
      }
      prm[e] = v;
    }
  } else if (bx == 9) {
    // bias tile: column j holds bias vector j (see decoder_layout.h)
    unsigned char* tile = img + kBiasTileOff;
    const int nslots = fg ? a.K - 1 : 1;
    for (int e = e_first; e < 64 * 64; e += e_step) {
      const int n = e >> 6, j = e & 63;
      float v = 0.f;
      if (n < Z) {
        if (j < 4) {
          v = j == 0 ? before_b[1][n] : j == 1 ? before_b[2][n] : j == 2 ? after_b[1][n] : after_b[2][n];
        } else if (j >= 16 && (j - 16) / 2 < nslots && (j - 16) / 2 < 15) {
          const int s = (j - 16) / 2, which = (j - 16) & 1;
          const float* z = a.z_slots + (fg ? (s + 1) : 0) * Z;
          const float* wrow = which == 0 ? before_w[0] + n * I + kPosenc : after_w[0] + n * (I + Z) + kPosenc;
          float acc = which == 0 ? before_b[0][n] : after_b[0][n];
          for (int c = 0; c < Z; ++c) acc = fmaf(wrow[c], z[c], acc);
          v = acc;
        }
      }
      put(tile, a.passes, a.f16, sw128_off(n, j), v);
    }
  } else {
    // slot biases (fp32, read by the backward kernel): block (bx - 10) = slot index within this image
    const int s = bx - 10;
    const int nslots = fg ? a.K - 1 : 1;
    if (s >= nslots) return;
    const float* z = a.z_slots + (fg ? (s + 1) : 0) * Z;
    float* sb = reinterpret_cast<float*>(img + nplanes * kPlaneBytes + kParamBytes) + s * kSlotBiasFloats;
    for (int e = e_first; e < kSlotBiasFloats; e += e_step) {
      const int which = e >> 6, n = e & 63;
      float v = 0.f;
      if (n < Z) {
        const float* wrow = which == 0 ? before_w[0] + n * I + kPosenc : after_w[0] + n * (I + Z) + kPosenc;
        float acc = which == 0 ? before_b[0][n] : after_b[0][n];
        for (int c = 0; c < Z; ++c) acc = fmaf(wrow[c], z[c], acc);
        v = acc;
      }
      sb[e] = v;
    }
  }
}

int launch_decoder_prep(const uorf_decoder_weights* w, const float* z_slots, int K, int Z, int passes, int f16, void* image,
                        cudaStream_t stream) {
  PrepArgs a;
  a.w = *w;
  a.z_slots = z_slots;
  a.K = K;
  a.Z = Z;
  a.passes = passes;
  a.f16 = f16;
  a.image = static_cast<unsigned char*>(image);
  dim3 grid(10 + (K - 1 > 1 ? K - 1 : 1), 2, kPrepSplit);
  decoder_prep_kernel<<<grid, 256, 0, stream>>>(a);
  return finish_launch(1);
}

}  // namespace uorf
