// Kernel 1 — frustum sample generation.
// Replaces Projection.construct_frus_coor / construct_sampling_coor, reference models/projection.py:33-113:
//   pixel grid (x, y, depth index) -> [x*z, y*z, z, 1] -> spixel2cam -> cam2world[n] -> world2nss (1/nss_scale),
//   z_vals = (z_cam - near) / (far - near), ray_dir = spixel2cam[:3,:3] . [x, y, 1],
//   optional strided partition into scale^2 ray chunks (projection.py:71-77) or a crop window
//   (uorf_nogan_model.py:153-165).  HBM-bound: ~16 B written per sample, inputs are a few hundred bytes.
// One thread per sample; all three outputs are written with unit-stride across the warp.
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

struct CoordParams {
  float S[16];   // spixel2cam
  const float* z_cam;      // [D]
  const float* cam2world;  // [N][16]
  float inv_nss;           // world2nss diagonal
  float near_plane, inv_range_den;  // z_vals = (z - near) / den
  int W, H, D, N;
  int scale, crop_h, crop_w, Hc, Wc, ray_major;
  float* coords;
  float* z_vals;
  float* ray_dir;
};

// IdxT = unsigned for every realistic frustum (< 2^32 samples): 32-bit div/mod is several times cheaper than 64-bit
template <typename IdxT>
__global__ void __launch_bounds__(256) sample_coords_kernel(const CoordParams p) {
  const IdxT per_chunk = (IdxT)p.N * p.D * p.Hc * p.Wc;
  const IdxT rays_per_chunk = (IdxT)p.N * p.Hc * p.Wc;
  const IdxT total = per_chunk * p.scale * p.scale;
  const IdxT idx = (IdxT)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;

  // ---- coords ----
  {
    const int chunk = (int)(idx / per_chunk);
    IdxT q = idx - (IdxT)chunk * per_chunk;
    int n, d, hc, wc;
    if (p.ray_major) {
      d = (int)(q % p.D); q /= p.D;
      wc = (int)(q % p.Wc); q /= p.Wc;
      hc = (int)(q % p.Hc); n = (int)(q / p.Hc);
    } else {
      wc = (int)(q % p.Wc); q /= p.Wc;
      hc = (int)(q % p.Hc); q /= p.Hc;
      d = (int)(q % p.D); n = (int)(q / p.D);
    }
    const int h0 = p.crop_h >= 0 ? p.crop_h : chunk / p.scale;
    const int w0 = p.crop_w >= 0 ? p.crop_w : chunk % p.scale;
    const int h = h0 + hc * p.scale, w = w0 + wc * p.scale;
    const float zc = __ldg(p.z_cam + d);
    const float p0 = (float)w * zc, p1 = (float)h * zc, p2 = zc;
    float cam[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
      cam[r] = fmaf(p.S[r * 4 + 3], 1.f, fmaf(p.S[r * 4 + 2], p2, fmaf(p.S[r * 4 + 1], p1, p.S[r * 4] * p0)));
    const float* C = p.cam2world + n * 16;
    float out[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float wv = fmaf(__ldg(C + r * 4 + 3), cam[3],
                            fmaf(__ldg(C + r * 4 + 2), cam[2], fmaf(__ldg(C + r * 4 + 1), cam[1], __ldg(C + r * 4) * cam[0])));
      out[r] = wv * p.inv_nss;
    }
    float* dst = p.coords + (long long)idx * 3;
    dst[0] = out[0];
    dst[1] = out[1];
    dst[2] = out[2];
  }
  // ---- z_vals: [chunk][ray][D]; identical for every ray, so the flat index idx maps onto it directly ----
  if (p.z_vals) {
    const int d = (int)(idx % p.D);
    p.z_vals[idx] = (__ldg(p.z_cam + d) - p.near_plane) / p.inv_range_den;
  }
  // ---- ray_dir: [chunk][ray][3]; the first rays_total threads each write one ray ----
  if (p.ray_dir && idx < rays_per_chunk * p.scale * p.scale) {
    const int chunk = (int)(idx / rays_per_chunk);
    IdxT q = idx - (IdxT)chunk * rays_per_chunk;
    const int wc = (int)(q % p.Wc); q /= p.Wc;
    const int hc = (int)(q % p.Hc);
    const int h0 = p.crop_h >= 0 ? p.crop_h : chunk / p.scale;
    const int w0 = p.crop_w >= 0 ? p.crop_w : chunk % p.scale;
    const float x = (float)(w0 + wc * p.scale), y = (float)(h0 + hc * p.scale);
    float* dst = p.ray_dir + (long long)idx * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) dst[r] = fmaf(p.S[r * 4 + 2], 1.f, fmaf(p.S[r * 4 + 1], y, p.S[r * 4] * x));
  }
}

int launch_sample_coords(const uorf_frustum* f, const float* z_cam, const float* cam2world, int n_views, int scale, int crop_h,
                         int crop_w, int Hc, int Wc, int ray_major, float* coords, float* z_vals, float* ray_dir,
                         cudaStream_t stream) {
  CoordParams p;
  for (int i = 0; i < 16; ++i) p.S[i] = f->spixel2cam[i];
  p.z_cam = z_cam;
  p.cam2world = cam2world;
  p.inv_nss = 1.0f / f->nss_scale;
  p.near_plane = f->near_plane;
  p.inv_range_den = f->far_plane - f->near_plane;
  p.W = f->W; p.H = f->H; p.D = f->D; p.N = n_views;
  p.scale = scale; p.crop_h = crop_h; p.crop_w = crop_w; p.Hc = Hc; p.Wc = Wc; p.ray_major = ray_major;
  p.coords = coords; p.z_vals = z_vals; p.ray_dir = ray_dir;
  const long long total = (long long)n_views * f->D * Hc * Wc * scale * scale;
  if (total == 0) return UORF_OK;
  const long long blocks = (total + 255) / 256;
  if (total < (1LL << 32) - 256) sample_coords_kernel<unsigned><<<(unsigned)blocks, 256, 0, stream>>>(p);
  else sample_coords_kernel<unsigned long long><<<(unsigned)blocks, 256, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

}  // namespace uorf
