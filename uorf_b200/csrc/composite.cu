// Kernel 3 — alpha compositing along each ray as a warp-level scan (forward and backward).
// Replaces raw2outputs, reference models/model.py:279-313:
//   delta_i = (z_{i+1} - z_i | 1e-2 for the last sample) * |rays_d|
//   alpha_i = 1 - exp(-sigma_i * delta_i);  T_i = prod_{j<i} (1 - alpha_j + 1e-10);  w_i = alpha_i * T_i
//   rgb_map = sum_i w_i rgb_i;  weights_norm = (w + 1e-5) / sum(w + 1e-5) [detached];  depth = sum wn_i z_i
//   mask_map = sum_i w_i sigma_i   (render_mask)
// One warp per ray; lane l owns samples l, l+32, ...; the transmittance is an exclusive product scan
// (shuffle-based, carried across the 32-sample chunks).  HBM-bound: 24 B per sample + 28 B per ray.
#include "common.cuh"

namespace uorf {

constexpr int kMaxChunks = 8;  // D <= 256

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive product scan across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= u;
  }
  return v;
}
// inclusive suffix-sum scan across the warp (lane l gets sum over lanes >= l)
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float u = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += u;
  }
  return v;
}

__global__ void __launch_bounds__(256) composite_fwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                                                            const float* __restrict__ rays_d, long long R, long long Rg, int D,
                                                            float* __restrict__ rgb_map, float* __restrict__ depth_map,
                                                            float* __restrict__ weights_norm, float* __restrict__ mask_map) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nchunks = (D + 31) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
    const long long rg = r % Rg;  // geometry (z_vals, rays_d) may be shared by several stacked ray sets
    const float dx = __ldg(rays_d + rg * 3), dy = __ldg(rays_d + rg * 3 + 1), dz = __ldg(rays_d + rg * 3 + 2);
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float4* rr = raw + r * D;
    const float* zz = z_vals + rg * D;
    float w[kMaxChunks], zv[kMaxChunks];
    float carry = 1.f;  // product of (1 - alpha + 1e-10) over all previous chunks
    float ar = 0.f, ag = 0.f, ab = 0.f, am = 0.f, sw = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      w[c] = 0.f;
      zv[c] = 0.f;
      if (c < nchunks) {
        const int i = c * 32 + lane;
        const bool in = i < D;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float z0 = 0.f, z1 = 0.f;
        if (in) {
          v = __ldg(rr + i);
          z0 = __ldg(zz + i);
          z1 = (i + 1 < D) ? __ldg(zz + i + 1) : 0.f;
        }
        const float delta = ((i + 1 < D) ? (z1 - z0) : 1e-2f) * nrm;
        const float alpha = in ? 1.f - expf(-v.w * delta) : 0.f;
        const float term = in ? (1.f - alpha + 1e-10f) : 1.f;
        const float inc = warp_scan_mul(term, lane);
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 1.f;
        const float T = carry * exc;
        carry *= __shfl_sync(0xffffffffu, inc, 31);
        const float wi = alpha * T;
        w[c] = wi;
        zv[c] = z0;
        ar += wi * v.x;
        ag += wi * v.y;
        ab += wi * v.z;
        am += wi * v.w;
        if (in) sw += wi + 1e-5f;
      }
    }
    ar = warp_sum(ar);
    ag = warp_sum(ag);
    ab = warp_sum(ab);
    sw = warp_sum(sw);
    float dep = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      if (c < nchunks) {
        const int i = c * 32 + lane;
        if (i < D) {
          const float wn = (w[c] + 1e-5f) / sw;
          dep += wn * zv[c];
          if (weights_norm) weights_norm[r * D + i] = wn;
        }
      }
    }
    dep = warp_sum(dep);
    if (mask_map) am = warp_sum(am);
    if (lane == 0) {
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
__global__ void __launch_bounds__(256) composite_bwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                                                            const float* __restrict__ rays_d, const float* __restrict__ grad_rgb,
                                                            long long R, long long Rg, int D, float4* __restrict__ grad_raw) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nchunks = (D + 31) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
    const long long rg = r % Rg;  // geometry (z_vals, rays_d) may be shared by several stacked ray sets
    const float dx = __ldg(rays_d + rg * 3), dy = __ldg(rays_d + rg * 3 + 1), dz = __ldg(rays_d + rg * 3 + 2);
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float g0 = __ldg(grad_rgb + r * 3), g1 = __ldg(grad_rgb + r * 3 + 1), g2 = __ldg(grad_rgb + r * 3 + 2);
    const float4* rr = raw + r * D;
    const float* zz = z_vals + rg * D;
    float w[kMaxChunks], T[kMaxChunks], term[kMaxChunks], G[kMaxChunks], de[kMaxChunks];
    float carry = 1.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      w[c] = T[c] = G[c] = de[c] = 0.f;
      term[c] = 1.f;
      if (c < nchunks) {
        const int i = c * 32 + lane;
        const bool in = i < D;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float z0 = 0.f, z1 = 0.f;
        if (in) {
          v = __ldg(rr + i);
          z0 = __ldg(zz + i);
          z1 = (i + 1 < D) ? __ldg(zz + i + 1) : 0.f;
        }
        const float delta = ((i + 1 < D) ? (z1 - z0) : 1e-2f) * nrm;
        const float e = in ? expf(-v.w * delta) : 1.f;
        const float alpha = 1.f - e;
        const float tm = in ? (1.f - alpha + 1e-10f) : 1.f;
        const float inc = warp_scan_mul(tm, lane);
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 1.f;
        T[c] = carry * exc;
        carry *= __shfl_sync(0xffffffffu, inc, 31);
        w[c] = in ? alpha * T[c] : 0.f;
        term[c] = tm;
        G[c] = g0 * v.x + g1 * v.y + g2 * v.z;
        de[c] = delta * e;
      }
    }
    float tail = 0.f;  // sum of G*w over all later chunks
#pragma unroll
    for (int c = kMaxChunks - 1; c >= 0; --c) {
      if (c < nchunks) {
        const int i = c * 32 + lane;
        const float gw = G[c] * w[c];
        const float incl = warp_rscan_add(gw, lane);
        const float S = tail + incl - gw;
        tail += __shfl_sync(0xffffffffu, incl, 0);
        if (i < D) {
          const float dalpha = G[c] * T[c] - S / term[c];
          grad_raw[r * D + i] = make_float4(g0 * w[c], g1 * w[c], g2 * w[c], dalpha * de[c]);
        }
      }
    }
  }
}

static int composite_grid(long long R, int num_sms) {
  long long blocks = (R + 7) / 8;  // 8 warps (rays) per 256-thread block
  const long long cap = (long long)num_sms * 64;   // 8 resident CTAs per SM x 8 waves: enough rays in flight to cover HBM latency
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

int launch_composite_fwd(const float* raw, const float* z_vals, const float* rays_d, long long R, long long Rg, int D, float* rgb_map,
                         float* depth_map, float* weights_norm, float* mask_map, int num_sms, cudaStream_t stream) {
  if (R == 0) return UORF_OK;
  composite_fwd_kernel<<<composite_grid(R, num_sms), 256, 0, stream>>>(reinterpret_cast<const float4*>(raw), z_vals, rays_d, R, Rg, D,
                                                                       rgb_map, depth_map, weights_norm, mask_map);
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

int launch_composite_bwd(const float* raw, const float* z_vals, const float* rays_d, const float* grad_rgb, long long R, long long Rg, int D,
                         float* grad_raw, int num_sms, cudaStream_t stream) {
  if (R == 0) return UORF_OK;
  composite_bwd_kernel<<<composite_grid(R, num_sms), 256, 0, stream>>>(reinterpret_cast<const float4*>(raw), z_vals, rays_d, grad_rgb,
                                                                       R, Rg, D, reinterpret_cast<float4*>(grad_raw));
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

}  // namespace uorf
