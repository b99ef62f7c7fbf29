// Kernel 1 — frustum sample generation.
// Replaces Projection.construct_frus_coor / construct_sampling_coor, reference models/projection.py:33-113:
//   pixel grid (x, y, depth index) -> [x*z, y*z, z, 1] -> spixel2cam -> cam2world[n] -> world2nss (1/nss_scale),
//   z_vals = (z_cam - near) / (far - near), ray_dir = spixel2cam[:3,:3] . [x, y, 1],
//   optional strided partition into scale^2 ray chunks (projection.py:71-77) or a crop window
//   (uorf_nogan_model.py:153-165).  HBM-bound (write-only): 16 B per sample + 12 B per ray; inputs are a few hundred bytes.
// Persistent CTAs (a multiple of the SM count) loop over groups of 256 consecutive samples: the per-depth tables and the
// camera matrices are staged in shared memory once per CTA, and the xyz triples of a group are staged through shared
// memory so that global stores are full 16-byte, unit-stride vectors (a direct 12-byte-stride store would touch every
// 32-byte sector three times).
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kMaxStagedViews = 64;

struct CoordParams {
  float S[16];   // spixel2cam
  const float* z_cam;      // [D]
  const float* cam2world;  // [N][16]
  float inv_nss;           // world2nss diagonal
  float near_plane, inv_range_den;  // z_vals = (z - near) / den
  int W, H, D, N;
  int scale, crop_h, crop_w, Hc, Wc, ray_major;
  const int* crop_origin;  // device (h0, w0) overriding crop_h / crop_w (clamped to the frustum), or null
  int sh_d, sh_w, sh_h;   // log2 of D, Wc, Hc when they are powers of two (index math by shifts), else -1
  float* coords;
  float* z_vals;
  float* ray_dir;
};

// A frustum point is affine in the camera depth of its sample: with a_k = S[k][0] w + S[k][1] h + S[k][2],
//   cam_k = a_k z + S[k][3],  world_r = z (sum_k C[r][k] a_k) + sum_k C[r][k] S[k][3],
// so a ray needs one direction and one origin (both pre-scaled by 1/nss_scale) and a sample is ONE FMA per coordinate
// instead of the reference's 33-operation matmul chain (projection.py:64-67).  The result differs from the reference's
// fp32 matmuls by rounding only (<= 5e-7 on the |x| <= 2 NSS range; tests: <= 2e-6).  Both kernels use this function, so
// their outputs are bit-equal re-indexings of each other.
__device__ __forceinline__ void ray_affine(const float (&S)[16], const float (&c)[12], float inv_nss, float w, float h,
                                           float (&dirn)[3], float (&orgn)[3]) {
  float a[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = fmaf(S[k * 4 + 1], h, fmaf(S[k * 4], w, S[k * 4 + 2]));
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float d = fmaf(c[r * 4 + 3], a[3], fmaf(c[r * 4 + 2], a[2], fmaf(c[r * 4 + 1], a[1], __fmul_rn(c[r * 4], a[0]))));
    const float o = fmaf(c[r * 4 + 3], S[15], fmaf(c[r * 4 + 2], S[11], fmaf(c[r * 4 + 1], S[7], __fmul_rn(c[r * 4], S[3]))));
    dirn[r] = __fmul_rn(d, inv_nss);
    orgn[r] = __fmul_rn(o, inv_nss);
  }
}

// IdxT = unsigned for every realistic frustum (< 2^32 samples): 32-bit index math is several times cheaper than 64-bit
template <typename IdxT>
__global__ void __launch_bounds__(256) sample_coords_kernel(const CoordParams p) {
  // crop origin: host scalars, or (fine-stage training inside a CUDA graph) two device ints, clamped to the frustum
  int crop_h = p.crop_h, crop_w = p.crop_w;
  if (p.crop_origin != nullptr) {
    crop_h = min(max(__ldg(p.crop_origin), 0), p.H - p.Hc);
    crop_w = min(max(__ldg(p.crop_origin + 1), 0), p.W - p.Wc);
  }
  __shared__ __align__(16) float s_xyz[2][256 * 3];
  __shared__ float s_zc[256], s_zv[256];                 // z_cam and normalised z_vals per depth index (D <= 256)
  __shared__ float s_c2w[kMaxStagedViews][12];           // rows 0..2 of every view's cam2world
  const IdxT per_chunk = (IdxT)p.N * p.D * p.Hc * p.Wc;
  const IdxT rays_per_chunk = (IdxT)p.N * p.Hc * p.Wc;
  const IdxT total = per_chunk * p.scale * p.scale;
  const IdxT rays_total = rays_per_chunk * p.scale * p.scale;
  const bool ztab = p.D <= 256;
  const bool cstage = p.N <= kMaxStagedViews;
  const bool pow2 = p.sh_d >= 0 && p.sh_w >= 0 && p.sh_h >= 0;
  if (ztab && (int)threadIdx.x < p.D) {
    const float zc = __ldg(p.z_cam + threadIdx.x);
    s_zc[threadIdx.x] = zc;
    s_zv[threadIdx.x] = (zc - p.near_plane) / p.inv_range_den;
  }
  if (cstage)
    for (int e = threadIdx.x; e < p.N * 12; e += blockDim.x) s_c2w[e / 12][e % 12] = __ldg(p.cam2world + (e / 12) * 16 + e % 12);
  __syncthreads();

  const IdxT ngroups = (total + 255) / 256;
  int buf = 0;
  for (IdxT g = blockIdx.x; g < ngroups; g += gridDim.x, buf ^= 1) {
    const IdxT idx = g * 256 + threadIdx.x;
    const bool live = idx < total;
    if (live) {
      // ---- coords ----
      const int chunk = p.scale == 1 ? 0 : (int)(idx / per_chunk);
      IdxT q = idx - (IdxT)chunk * per_chunk;
      int n, d, hc, wc;
      if (p.ray_major && pow2) {
        d = (int)(q & (IdxT)(p.D - 1)); q >>= p.sh_d;
        wc = (int)(q & (IdxT)(p.Wc - 1)); q >>= p.sh_w;
        hc = (int)(q & (IdxT)(p.Hc - 1)); n = (int)(q >> p.sh_h);
      } else if (pow2) {
        wc = (int)(q & (IdxT)(p.Wc - 1)); q >>= p.sh_w;
        hc = (int)(q & (IdxT)(p.Hc - 1)); q >>= p.sh_h;
        d = (int)(q & (IdxT)(p.D - 1)); n = (int)(q >> p.sh_d);
      } else if (p.ray_major) {
        d = (int)(q % p.D); q /= p.D;
        wc = (int)(q % p.Wc); q /= p.Wc;
        hc = (int)(q % p.Hc); n = (int)(q / p.Hc);
      } else {
        wc = (int)(q % p.Wc); q /= p.Wc;
        hc = (int)(q % p.Hc); q /= p.Hc;
        d = (int)(q % p.D); n = (int)(q / p.D);
      }
      const int h0 = crop_h >= 0 ? crop_h : chunk / p.scale;
      const int w0 = crop_w >= 0 ? crop_w : chunk % p.scale;
      const int h = h0 + hc * p.scale, w = w0 + wc * p.scale;
      const float zc = ztab ? s_zc[d] : __ldg(p.z_cam + d);
      const float* Cg = p.cam2world + n * 16;
      float c[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) c[e] = cstage ? s_c2w[n][e] : __ldg(Cg + e);
      float dirn[3], orgn[3];
      ray_affine(p.S, c, p.inv_nss, (float)w, (float)h, dirn, orgn);
#pragma unroll
      for (int r = 0; r < 3; ++r) s_xyz[buf][threadIdx.x * 3 + r] = fmaf(zc, dirn[r], orgn[r]);   // stride 3 is coprime with the 32 banks: conflict-free
      // ---- z_vals: [chunk][ray][D]; identical for every ray, so the flat index idx maps onto it directly ----
      if (p.z_vals) {
        const int dz = p.sh_d >= 0 ? (int)(idx & (IdxT)(p.D - 1)) : (int)(idx % p.D);
        p.z_vals[idx] = ztab ? s_zv[dz] : (__ldg(p.z_cam + dz) - p.near_plane) / p.inv_range_den;
      }
      // ---- ray_dir: [chunk][ray][3]; the first rays_total flat indices each write one ray ----
      if (p.ray_dir && idx < rays_total) {
        const int rchunk = p.scale == 1 ? 0 : (int)(idx / rays_per_chunk);
        IdxT qr = idx - (IdxT)rchunk * rays_per_chunk;
        int rwc, rhc;
        if (pow2) { rwc = (int)(qr & (IdxT)(p.Wc - 1)); rhc = (int)((qr >> p.sh_w) & (IdxT)(p.Hc - 1)); }
        else { rwc = (int)(qr % p.Wc); qr /= p.Wc; rhc = (int)(qr % p.Hc); }
        const int rh0 = crop_h >= 0 ? crop_h : rchunk / p.scale;
        const int rw0 = crop_w >= 0 ? crop_w : rchunk % p.scale;
        const float x = (float)(rw0 + rwc * p.scale), y = (float)(rh0 + rhc * p.scale);
        float* dst = p.ray_dir + (long long)idx * 3;
#pragma unroll
        for (int r = 0; r < 3; ++r) dst[r] = fmaf(p.S[r * 4 + 2], 1.f, fmaf(p.S[r * 4 + 1], y, p.S[r * 4] * x));
      }
    }
    __syncthreads();   // s_xyz[buf] complete; the other buffer was drained before the previous barrier
    {
      const IdxT first = g * 256;                                  // 768 floats = 3072 B per group: 16 B aligned
      const IdxT n_here = total - first < (IdxT)256 ? total - first : (IdxT)256;
      float* dst = p.coords + (long long)first * 3;
      if (n_here == (IdxT)256 && (reinterpret_cast<uintptr_t>(p.coords) & 15) == 0) {
        if (threadIdx.x < 192) reinterpret_cast<float4*>(dst)[threadIdx.x] = reinterpret_cast<const float4*>(s_xyz[buf])[threadIdx.x];
      } else {
        for (unsigned e = threadIdx.x; e < (unsigned)n_here * 3; e += blockDim.x) dst[e] = s_xyz[buf][e];
      }
    }
  }
}

// Fast path for the layout the B200 drivers use (ray-major, D % 4 == 0, unpartitioned or cropped): one thread produces FOUR
// consecutive depth samples of one ray, so the pixel / view decode and the 12 camera-matrix reads are shared by 4 samples and
// every store is a 16-byte vector (3 x float4 of xyz, 1 x float4 of z_vals) - no shared-memory staging needed.
// Arithmetic per sample is the general kernel's (ray_affine + one FMA per coordinate).
// POW2: D/4, Wc, Hc are powers of two and there are < 2^32 quads: the index decode is shifts and masks on 32-bit values (the
// generic 64-bit divisions were ~40 % of the kernel's instructions, which made it issue-bound rather than HBM-bound).
template <bool POW2>
__global__ void __launch_bounds__(256) sample_coords_rm4_kernel(const CoordParams p) {
  // crop origin: host scalars, or (fine-stage training inside a CUDA graph) two device ints, clamped to the frustum
  int crop_h = p.crop_h, crop_w = p.crop_w;
  if (p.crop_origin != nullptr) {
    crop_h = min(max(__ldg(p.crop_origin), 0), p.H - p.Hc);
    crop_w = min(max(__ldg(p.crop_origin + 1), 0), p.W - p.Wc);
  }
  __shared__ __align__(16) float4 s_strip[8][96];
  __shared__ float s_zc[256], s_zv[256];
  __shared__ float s_c2w[kMaxStagedViews][12];
  const bool ztab = p.D <= 256;
  const bool cstage = p.N <= kMaxStagedViews;
  if (ztab && (int)threadIdx.x < p.D) {
    const float zc = __ldg(p.z_cam + threadIdx.x);
    s_zc[threadIdx.x] = zc;
    s_zv[threadIdx.x] = (zc - p.near_plane) / p.inv_range_den;
  }
  if (cstage)
    for (int e = threadIdx.x; e < p.N * 12; e += blockDim.x) s_c2w[e / 12][e % 12] = __ldg(p.cam2world + (e / 12) * 16 + e % 12);
  __syncthreads();
  const int dq = p.D >> 2;                                    // quads per ray
  const long long rays = (long long)p.N * p.Hc * p.Wc;
  const long long nquads = rays * dq;
  const unsigned lane = threadIdx.x & 31;
  // warp-uniform loop over the first quad of each warp (the shared-memory transpose below needs whole warps)
  for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < nquads; base += (long long)gridDim.x * blockDim.x) {
    const long long qi = base + lane;
    const bool full = base + 32 <= nquads;
    if (qi >= nquads) continue;                               // ragged last warp only (then full == false: no warp-level sync below)
    long long ray;
    int d0, wc, hc, n;
    if (POW2) {
      const unsigned q32 = (unsigned)qi, r32 = q32 >> (p.sh_d - 2);
      ray = r32;
      d0 = (int)(q32 & (unsigned)(dq - 1)) * 4;
      wc = (int)(r32 & (unsigned)(p.Wc - 1));
      hc = (int)((r32 >> p.sh_w) & (unsigned)(p.Hc - 1));
      n = (int)(r32 >> (p.sh_w + p.sh_h));
    } else {
      ray = qi / dq;
      d0 = (int)(qi - ray * dq) * 4;
      wc = (int)(ray % p.Wc);
      const long long t2 = ray / p.Wc;
      hc = (int)(t2 % p.Hc); n = (int)(t2 / p.Hc);
    }
    const int h = (crop_h >= 0 ? crop_h : 0) + hc, w = (crop_w >= 0 ? crop_w : 0) + wc;
    float c[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) c[e] = cstage ? s_c2w[n][e] : __ldg(p.cam2world + n * 16 + e);
    float dirn[3], orgn[3];
    ray_affine(p.S, c, p.inv_nss, (float)w, (float)h, dirn, orgn);
    float xyz[12], zv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + j;
      const float zc = ztab ? s_zc[d] : __ldg(p.z_cam + d);
      zv[j] = ztab ? s_zv[d] : (zc - p.near_plane) / p.inv_range_den;
#pragma unroll
      for (int r = 0; r < 3; ++r) xyz[3 * j + r] = fmaf(zc, dirn[r], orgn[r]);
    }
    // the 48 B a thread produces are transposed through a per-warp shared-memory strip so that each of the three store
    // instructions of the warp writes 512 contiguous bytes (a direct store puts 16 B at a 48 B stride: every 32 B sector
    // is written in two halves by different instructions)
    float4* strip = s_strip[threadIdx.x >> 5];
    if (full) {
      strip[3 * lane] = make_float4(xyz[0], xyz[1], xyz[2], xyz[3]);
      strip[3 * lane + 1] = make_float4(xyz[4], xyz[5], xyz[6], xyz[7]);
      strip[3 * lane + 2] = make_float4(xyz[8], xyz[9], xyz[10], xyz[11]);
      __syncwarp();
      float4* dst = reinterpret_cast<float4*>(p.coords) + (qi - lane) * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) dst[k * 32 + lane] = strip[k * 32 + lane];
      __syncwarp();
    } else {
      float4* dst = reinterpret_cast<float4*>(p.coords + qi * 12);
      dst[0] = make_float4(xyz[0], xyz[1], xyz[2], xyz[3]);
      dst[1] = make_float4(xyz[4], xyz[5], xyz[6], xyz[7]);
      dst[2] = make_float4(xyz[8], xyz[9], xyz[10], xyz[11]);
    }
    if (p.z_vals) reinterpret_cast<float4*>(p.z_vals)[qi] = make_float4(zv[0], zv[1], zv[2], zv[3]);
    if (p.ray_dir && d0 == 0) {
      const float x = (float)w, y = (float)h;
#pragma unroll
      for (int r = 0; r < 3; ++r) p.ray_dir[ray * 3 + r] = fmaf(p.S[r * 4 + 2], 1.f, fmaf(p.S[r * 4 + 1], y, p.S[r * 4] * x));
    }
  }
}

int launch_sample_coords(const uorf_frustum* f, const float* z_cam, const float* cam2world, int n_views, int scale, int crop_h,
                         int crop_w, const int* crop_origin, int Hc, int Wc, int ray_major, float* coords, float* z_vals, float* ray_dir,
                         int num_sms, cudaStream_t stream) {
  CoordParams p;
  for (int i = 0; i < 16; ++i) p.S[i] = f->spixel2cam[i];
  p.z_cam = z_cam;
  p.cam2world = cam2world;
  p.inv_nss = 1.0f / f->nss_scale;
  p.near_plane = f->near_plane;
  p.inv_range_den = f->far_plane - f->near_plane;
  p.W = f->W; p.H = f->H; p.D = f->D; p.N = n_views;
  p.scale = scale; p.crop_h = crop_h; p.crop_w = crop_w; p.crop_origin = crop_origin; p.Hc = Hc; p.Wc = Wc; p.ray_major = ray_major;
  p.coords = coords; p.z_vals = z_vals; p.ray_dir = ray_dir;
  auto lg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  p.sh_d = lg(f->D); p.sh_w = lg(Wc); p.sh_h = lg(Hc);
  const long long total = (long long)n_views * f->D * Hc * Wc * scale * scale;
  if (total == 0) return UORF_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms * 8;   // persistent: 8 resident CTAs per SM
  if (blocks > cap) blocks = cap;
  const bool aligned = ((reinterpret_cast<uintptr_t>(coords) | reinterpret_cast<uintptr_t>(z_vals)) & 15) == 0;
  if (ray_major && scale == 1 && (f->D & 3) == 0 && aligned) {
    long long b4 = (total / 4 + 255) / 256;
    if (b4 > cap) b4 = cap;
    if (p.sh_d >= 2 && p.sh_w >= 0 && p.sh_h >= 0 && total / 4 < (1LL << 32)) sample_coords_rm4_kernel<true><<<(unsigned)b4, 256, 0, stream>>>(p);
    else sample_coords_rm4_kernel<false><<<(unsigned)b4, 256, 0, stream>>>(p);
  } else if (total < (1LL << 32) - 512) sample_coords_kernel<unsigned><<<(unsigned)blocks, 256, 0, stream>>>(p);
  else sample_coords_kernel<unsigned long long><<<(unsigned)blocks, 256, 0, stream>>>(p);
  return finish_launch(1);
}

}  // namespace uorf
