// Kernel 3 — alpha compositing along each ray as a warp-level scan (forward and backward).
// Replaces raw2outputs, reference models/model.py:279-313:
//   delta_i = (z_{i+1} - z_i | 1e-2 for the last sample) * |rays_d|
//   alpha_i = 1 - exp(-sigma_i * delta_i);  T_i = prod_{j<i} (1 - alpha_j + 1e-10);  w_i = alpha_i * T_i
//   rgb_map = sum_i w_i rgb_i;  weights_norm = (w + 1e-5) / sum(w + 1e-5) [detached];  depth = sum wn_i z_i
//   mask_map = sum_i w_i sigma_i   (render_mask)
// A ray is owned by a group of L lanes (L = power of two <= 32, several rays per warp when D is small); every lane holds S
// CONSECUTIVE samples (S = 4, or 8 for D > 128), which it loads as S contiguous float4 (64-128 B per lane, 1-4 KB per
// warp).  The transmittance is an exclusive product scan: sequential inside a lane, a log2(L)-step segmented shuffle scan
// across the group; the ray sums are segmented butterfly reductions.  HBM-bound: 24 B per sample + 28 B per ray.
#include "common.cuh"
#include <cstdlib>

// Resident CTAs per SM the register allocation aims for.  Measured on the 524,288-ray x 64-sample shape (fwd / bwd fraction
// of the HBM peak): unconstrained (54 registers, 4 CTAs) 0.68-0.79 / 0.70, 5 CTAs 0.86 / 0.81, 6 CTAs 0.86 / 0.74.
// Rejected: staging the raw loads through shared memory so each warp load is 512 contiguous bytes (0.44-0.78), and 8
// samples per lane at D = 64 (0.55).
#ifndef UORF_COMPOSITE_MINB
#define UORF_COMPOSITE_MINB 5
#endif

namespace uorf {

template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// loads the S samples [i0, i0+S) of ray r and the geometry they need; out-of-range samples get neutral values
template <int S>
__device__ __forceinline__ void load_samples(const float4* __restrict__ raw, const float* __restrict__ zz, long long roff, int i0, int D,
                                             bool active, bool vec_z, float4 (&v)[S], float (&z)[S + 1]) {
#pragma unroll
  for (int j = 0; j < S; ++j) {
    const int i = i0 + j;
    v[j] = (active && i < D) ? __ldg(raw + roff + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (vec_z && active && i0 + S <= D) {
#pragma unroll
    for (int q = 0; q < S / 4; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(zz + i0) + q);
      z[4 * q] = t.x; z[4 * q + 1] = t.y; z[4 * q + 2] = t.z; z[4 * q + 3] = t.w;
    }
    z[S] = (i0 + S < D) ? __ldg(zz + i0 + S) : 0.f;
  } else {
#pragma unroll
    for (int j = 0; j <= S; ++j) z[j] = (active && i0 + j < D) ? __ldg(zz + i0 + j) : 0.f;
  }
}

template <int S, int L>
__global__ void __launch_bounds__(256, UORF_COMPOSITE_MINB) composite_fwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                                                            const float* __restrict__ rays_d, long long R, long long Rg, int D, int vec,
                                                            float* __restrict__ rgb_map, float* __restrict__ depth_map,
                                                            float* __restrict__ weights_norm, float* __restrict__ mask_map) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (L - 1);          // lane within the ray's group
  constexpr int rpw = 32 / L;              // rays per warp
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool vec_z = vec != 0;   // D % 4 == 0 and 16-byte aligned z_vals / weights_norm: float4 accesses
  const int i0 = sub * S;
  for (long long wi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi * rpw < R; wi += warps) {
    const long long r = wi * rpw + lane / L;
    const bool active = r < R;
    const long long rg = active ? r % Rg : 0;   // geometry (z_vals, rays_d) may be shared by several stacked ray sets
    float nrm = 0.f;
    if (active) {
      const float dx = __ldg(rays_d + rg * 3), dy = __ldg(rays_d + rg * 3 + 1), dz = __ldg(rays_d + rg * 3 + 2);
      nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    float4 v[S];
    float z[S + 1];
    load_samples<S>(raw, z_vals + rg * D, active ? r * D : 0, i0, D, active, vec_z, v, z);
    float w[S], texc[S];
    float lane_prod = 1.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const int i = i0 + j;
      const bool in = active && i < D;
      const float delta = ((i + 1 < D) ? (z[j + 1] - z[j]) : 1e-2f) * nrm;
      const float alpha = in ? 1.f - expf(-v[j].w * delta) : 0.f;
      texc[j] = lane_prod;                 // product of the terms before sample j inside this lane
      lane_prod *= in ? (1.f - alpha + 1e-10f) : 1.f;
      w[j] = alpha;
    }
    // exclusive product over the previous lanes of the group
    float inc = lane_prod;
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
      const float u = __shfl_up_sync(0xffffffffu, inc, o, L);
      if (sub >= o) inc *= u;
    }
    float exc = __shfl_up_sync(0xffffffffu, inc, 1, L);
    if (sub == 0) exc = 1.f;
    float ar = 0.f, ag = 0.f, ab = 0.f, am = 0.f, sw = 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const float wi_ = w[j] * (exc * texc[j]);
      w[j] = wi_;
      ar += wi_ * v[j].x;
      ag += wi_ * v[j].y;
      ab += wi_ * v[j].z;
      am += wi_ * v[j].w;
      if (active && i0 + j < D) sw += wi_ + 1e-5f;
    }
    ar = group_sum<L>(ar);
    ag = group_sum<L>(ag);
    ab = group_sum<L>(ab);
    sw = group_sum<L>(sw);
    if (mask_map) am = group_sum<L>(am);
    float dep = 0.f;
    float wn[S];
    const float inv_sw = 1.f / sw;         // one division per lane; the per-sample products differ from w / sum by <= 1 ulp
#pragma unroll
    for (int j = 0; j < S; ++j) {
      wn[j] = (active && i0 + j < D) ? (w[j] + 1e-5f) * inv_sw : 0.f;
      dep += wn[j] * z[j];
    }
    dep = group_sum<L>(dep);
    if (weights_norm && active) {
      float* dst = weights_norm + r * D + i0;
      if (vec_z && i0 + S <= D) {
#pragma unroll
        for (int q = 0; q < S / 4; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(wn[4 * q], wn[4 * q + 1], wn[4 * q + 2], wn[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < S; ++j)
          if (i0 + j < D) dst[j] = wn[j];
      }
    }
    if (sub == 0 && active) {
      rgb_map[r * 3] = ar;
      rgb_map[r * 3 + 1] = ag;
      rgb_map[r * 3 + 2] = ab;
      depth_map[r] = dep;
      if (mask_map) mask_map[r] = am;
    }
  }
}

// grad of rgb_map w.r.t. raw:  d rgb_i = g * w_i ;  G_i = g . rgb_i ;  S_i = sum_{m>i} G_m w_m ;
//   d alpha_i = G_i T_i - S_i / (1 - alpha_i + 1e-10) ;  d sigma_i = d alpha_i * delta_i * exp(-sigma_i delta_i)
template <int S, int L>
__global__ void __launch_bounds__(256, UORF_COMPOSITE_MINB) composite_bwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                                                            const float* __restrict__ rays_d, const float* __restrict__ grad_rgb,
                                                            long long R, long long Rg, int D, int vec, float4* __restrict__ grad_raw) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (L - 1);
  constexpr int rpw = 32 / L;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool vec_z = vec != 0;   // D % 4 == 0 and 16-byte aligned z_vals / weights_norm: float4 accesses
  const int i0 = sub * S;
  for (long long wi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi * rpw < R; wi += warps) {
    const long long r = wi * rpw + lane / L;
    const bool active = r < R;
    const long long rg = active ? r % Rg : 0;
    float nrm = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (active) {
      const float dx = __ldg(rays_d + rg * 3), dy = __ldg(rays_d + rg * 3 + 1), dz = __ldg(rays_d + rg * 3 + 2);
      nrm = sqrtf(dx * dx + dy * dy + dz * dz);
      g0 = __ldg(grad_rgb + r * 3); g1 = __ldg(grad_rgb + r * 3 + 1); g2 = __ldg(grad_rgb + r * 3 + 2);
    }
    float4 v[S];
    float z[S + 1];
    load_samples<S>(raw, z_vals + rg * D, active ? r * D : 0, i0, D, active, vec_z, v, z);
    float alpha[S], term[S], texc[S], de[S];
    float lane_prod = 1.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const int i = i0 + j;
      const bool in = active && i < D;
      const float delta = ((i + 1 < D) ? (z[j + 1] - z[j]) : 1e-2f) * nrm;
      const float e = in ? expf(-v[j].w * delta) : 1.f;
      alpha[j] = 1.f - e;
      term[j] = in ? (1.f - alpha[j] + 1e-10f) : 1.f;
      de[j] = delta * e;
      texc[j] = lane_prod;
      lane_prod *= term[j];
    }
    float inc = lane_prod;
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
      const float u = __shfl_up_sync(0xffffffffu, inc, o, L);
      if (sub >= o) inc *= u;
    }
    float exc = __shfl_up_sync(0xffffffffu, inc, 1, L);
    if (sub == 0) exc = 1.f;
    // suffix sums of G*w: sequential inside the lane (from the last sample), segmented reverse scan across the group
    float T[S], w[S], G[S], gw_after[S];
    float lane_gw = 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      T[j] = exc * texc[j];
      w[j] = (active && i0 + j < D) ? alpha[j] * T[j] : 0.f;
      G[j] = g0 * v[j].x + g1 * v[j].y + g2 * v[j].z;
    }
#pragma unroll
    for (int j = S - 1; j >= 0; --j) {
      gw_after[j] = lane_gw;               // sum over the later samples of this lane
      lane_gw += G[j] * w[j];
    }
    float rinc = lane_gw;                  // inclusive suffix sum over lanes >= sub of the group
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
      const float u = __shfl_down_sync(0xffffffffu, rinc, o, L);
      if (sub + o < L) rinc += u;
    }
    const float later = rinc - lane_gw;    // lanes after this one
    if (active) {
#pragma unroll
      for (int j = 0; j < S; ++j) {
        if (i0 + j < D) {
          const float Ssum = later + gw_after[j];
          const float dalpha = G[j] * T[j] - Ssum / term[j];
          grad_raw[r * D + i0 + j] = make_float4(g0 * w[j], g1 * w[j], g2 * w[j], dalpha * de[j]);
        }
      }
    }
  }
}

static void composite_shape(int D, int& S, int& L) {
  static const int forced = [] { const char* e = getenv("UORF_COMPOSITE_S"); return e ? atoi(e) : 0; }();   // tuning override
  S = D > 128 ? 8 : 4;
  if (forced == 8 && D > 32) S = 8;        // 8 consecutive samples per lane: 128 B of raw per lane in flight, one scan step fewer
  const int need = (D + S - 1) / S;
  L = 1;
  while (L < need) L <<= 1;
}
static int composite_vec(int D, const void* a, const void* b) {
  return ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0) ? 1 : 0;
}
static int composite_grid(long long R, int L, int num_sms) {
  const long long rays_per_block = 8LL * (32 / L);
  long long blocks = (R + rays_per_block - 1) / rays_per_block;
  const long long cap = (long long)num_sms * 64;   // 8 resident CTAs per SM x 8 waves: enough rays in flight to cover HBM latency
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// (S, L) are compile-time: the shuffle scans / butterflies unroll and interleave
#define UORF_COMPOSITE_DISPATCH(S_, L_, CALL)                                   \
  do {                                                                          \
    if (S_ == 4) {                                                              \
      switch (L_) {                                                             \
        case 1: { constexpr int kS = 4, kL = 1; CALL; } break;                  \
        case 2: { constexpr int kS = 4, kL = 2; CALL; } break;                  \
        case 4: { constexpr int kS = 4, kL = 4; CALL; } break;                  \
        case 8: { constexpr int kS = 4, kL = 8; CALL; } break;                  \
        case 16: { constexpr int kS = 4, kL = 16; CALL; } break;                \
        default: { constexpr int kS = 4, kL = 32; CALL; } break;                \
      }                                                                         \
    } else {                                                                    \
      switch (L_) {                                                             \
        case 8: { constexpr int kS = 8, kL = 8; CALL; } break;                  \
        case 16: { constexpr int kS = 8, kL = 16; CALL; } break;                \
        default: { constexpr int kS = 8, kL = 32; CALL; } break;                \
      }                                                                         \
    }                                                                           \
  } while (0)

int launch_composite_fwd(const float* raw, const float* z_vals, const float* rays_d, long long R, long long Rg, int D, float* rgb_map,
                         float* depth_map, float* weights_norm, float* mask_map, int num_sms, cudaStream_t stream) {
  if (R == 0) return UORF_OK;
  int S, L;
  composite_shape(D, S, L);
  if (L > 32) return UORF_ERR_UNSUPPORTED;       // D > 256
  const int grid = composite_grid(R, L, num_sms);
  const int vec = composite_vec(D, z_vals, weights_norm);
  UORF_COMPOSITE_DISPATCH(S, L, (composite_fwd_kernel<kS, kL><<<grid, 256, 0, stream>>>(
      reinterpret_cast<const float4*>(raw), z_vals, rays_d, R, Rg, D, vec, rgb_map, depth_map, weights_norm, mask_map)));
  return finish_launch(1);
}

int launch_composite_bwd(const float* raw, const float* z_vals, const float* rays_d, const float* grad_rgb, long long R, long long Rg, int D,
                         float* grad_raw, int num_sms, cudaStream_t stream) {
  if (R == 0) return UORF_OK;
  int S, L;
  composite_shape(D, S, L);
  if (L > 32) return UORF_ERR_UNSUPPORTED;
  const int grid = composite_grid(R, L, num_sms);
  const int vec = composite_vec(D, z_vals, nullptr);
  UORF_COMPOSITE_DISPATCH(S, L, (composite_bwd_kernel<kS, kL><<<grid, 256, 0, stream>>>(
      reinterpret_cast<const float4*>(raw), z_vals, rays_d, grad_rgb, R, Rg, D, vec, reinterpret_cast<float4*>(grad_raw))));
  return finish_launch(1);
}

}  // namespace uorf
