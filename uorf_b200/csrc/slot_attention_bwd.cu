// Kernel 5 (slot-attention part) — backward of SlotAttention.forward (reference models/model.py:205-258), batch 1.
//
// Given d slots [K][C] (the gradient arriving from the decoder's z_slots; `attn` is detached by the reference's training
// driver, models/uorf_nogan_model.py:186) it produces the gradient of every SlotAttention parameter and of `feat`.
// All three iterations are unrolled backwards; the forward pass saved, per iteration, only slot_in / q / upd / sum(attn)
// (3KC + K floats) - attention maps, GRU gates, LayerNorm statistics and MLP activations are recomputed here.
//
// Launch structure (fp32 FFMA; the problem is ~0.3 GFLOP and latency-bound like the forward):
//   per iteration, last to first:
//     sa_bwd_slot   one CTA per slot : [query path of the later iteration: d q -> to_q, LN -> d slot]  +
//                                       residual MLP, LN, GRUCell backward -> d upd, d slot_in (part)
//     sa_bwd_token  token-parallel   : recompute softmax-over-slots, d w, d attn, d dots -> d k, d v (accumulated in place),
//                                       per-CTA partial d q
//   then  sa_bwd_slot (query path of iteration 0 + slot initialisation), sa_bwd_feat (d k, d v -> to_k, to_v, LN(feat)),
//   sa_bwd_reduce (fixed-order reductions of the per-slot / per-CTA partial parameter gradients).
// Parameter gradients are accumulated without atomics: each slot CTA owns a private region, each token CTA a private
// partial; the final reduction adds them in a fixed order, so results are deterministic.
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kBTok = 32;
constexpr int kBThreads = 256;
constexpr int kBSlotThreads = 1024;   // per-slot CTAs: latency-bound matvecs, more warps = more weight rows in flight
constexpr int kBMaxC = 64;
constexpr int kBMaxK = 16;
constexpr int kBMaxHid = 128;
__host__ __device__ inline int sab_saved_stride(int K, int C) { return 3 * K * C + ((K + 3) & ~3); }

// per-slot gradient region (floats)
// The weight gradients of the per-slot matrices are sums of outer products a (x) b over iterations and slots.  The per-slot kernel
// only STORES the factor vectors (12 C + 2 hidden floats per iteration and slot); the reduce kernel forms every weight element as
// a short sum of products.  (Accumulating the outer products into a per-slot global region cost five dependent read-modify-write
// sweeps of up to 12,288 elements per launch - two thirds of the per-slot kernel's time.)
struct SabVec {
  int q_a, q_b;        // d q [C], LN_q(slot) [C]                   -> to_q weight
  int ds, m;           // d slot_out [C], relu hidden [hidden]      -> residual MLP second weight
  int dm, ln;          // d hidden (gated) [hidden], LN_res(h') [C] -> residual MLP first weight
  int dg;              // d gates: input part [3C], hidden part [3C]
  int upd, h;          // GRU input [C], GRU state [C]              -> gru_w_ih = dg[0:3C] (x) upd, gru_w_hh = dg[3C:6C] (x) h
  int total;
};
__host__ __device__ inline SabVec sab_vec(int C, int Hd) {
  SabVec v;
  v.q_a = 0; v.q_b = C; v.ds = 2 * C; v.m = 3 * C; v.dm = 3 * C + Hd; v.ln = 3 * C + 2 * Hd; v.dg = 4 * C + 2 * Hd;
  v.upd = 10 * C + 2 * Hd; v.h = 11 * C + 2 * Hd; v.total = 12 * C + 2 * Hd;
  return v;
}
constexpr int kBMaxIters = 8;

struct SlotRegion {
  int wih, whh, bih, bhh, lnr_w, lnr_b, w1, b1, w2, b2, lnq_w, lnq_b, wq, total;
};
__host__ __device__ inline SlotRegion slot_region(int C, int Hd) {
  SlotRegion r;
  int o = 0;
  r.wih = o; o += 3 * C * C;
  r.whh = o; o += 3 * C * C;
  r.bih = o; o += 3 * C;
  r.bhh = o; o += 3 * C;
  r.lnr_w = o; o += C;
  r.lnr_b = o; o += C;
  r.w1 = o; o += Hd * C;
  r.b1 = o; o += Hd;
  r.w2 = o; o += C * Hd;
  r.b2 = o; o += C;
  r.lnq_w = o; o += C;
  r.lnq_b = o; o += C;
  r.wq = o; o += C * C;
  r.total = (o + 3) & ~3;
  return r;
}

struct SabParams {
  uorf_slot_attention_weights w;
  uorf_slot_attention_grads g;
  const float* feat; long long feat_ts, feat_cs;
  const float* noise_fg; const float* noise_bg;
  int ntok, C, K, hidden, iters;
  float eps;
  const float* kbuf; const float* vbuf; const float* saved;   // from the forward workspace
  const float* grad_slots;   // [K][C]
  float* dS;        // [K][C] running d slot
  float* dU;        // [K][C] d upd of the current iteration
  float* dK;        // [ntok][C]
  float* dV;        // [ntok][C]
  float* dq_part;   // [nblk][K][C]
  float* slot_reg;  // [K][region]
  float* tok_part;  // [nblk][2*C*C + 2*C]   d to_k | d to_v | d norm_feat w | b
  float* vec_store; // [iters][K][12 C + 2 hidden]: the factor pairs of every weight-gradient outer product (see SabVec)
  float* grad_feat; // [ntok][C]
  int nblk;
};

// y[o] = (bias ? bias[o] : 0) + sum_c W[o][c] x[c]; one warp per output row (coalesced weight reads)
__device__ __forceinline__ void b_matvec(const float* W, const float* bias, const float* x, float* y, int n_out, int n_in) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < n_out; o += nw) {
    float acc = 0.f;
    for (int c = lane; c < n_in; c += 32) acc = fmaf(W[(long long)o * n_in + c], x[c], acc);   // generic load: global or staged shared memory
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) y[o] = acc + (bias ? bias[o] : 0.f);
  }
}
// y[c] = sum_o W[o][c] x[o]  (transposed product; threads over c -> coalesced)
__device__ __forceinline__ void b_matvec_t(const float* W, const float* x, float* y, int n_out, int n_in) {
  for (int c = threadIdx.x; c < n_in; c += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (int o = 0; o < n_out; ++o) acc = fmaf(W[(long long)o * n_in + c], x[o], acc);
    y[c] = acc;
  }
}
__device__ __forceinline__ void b_vec_acc(float* G, const float* a, int n) {
  for (int e = threadIdx.x; e < n; e += blockDim.x) G[e] += a[e];
}
// LayerNorm forward by warp 0: xhat, rstd (broadcast through smem), y = xhat*w + b
__device__ __forceinline__ void b_ln_fwd(const float* x, const float* __restrict__ w, const float* __restrict__ b, float* xhat, float* y,
                                         float* stat, int C) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += x[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = x[c] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + 1e-5f);
    for (int c = lane; c < C; c += 32) {
      const float xh = (x[c] - mean) * rstd;
      xhat[c] = xh;
      y[c] = xh * w[c] + b[c];
    }
    if (lane == 0) stat[0] = rstd;
  }
}
// LayerNorm backward by warp 0: dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), dxhat = dy * w
__device__ __forceinline__ void b_ln_bwd(const float* dy, const float* xhat, const float* __restrict__ w, float rstd, float* dx, int C) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = dy[c] * w[c]; s1 += d; s2 += d * xhat[c]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    s1 /= C; s2 /= C;
    for (int c = lane; c < C; c += 32) dx[c] = rstd * (dy[c] * w[c] - s1 - xhat[c] * s2);
  }
}

// ---- per-slot kernel --------------------------------------------------------------------------------
// do_q : query-path backward of iteration it_q (needs dq_part from the token pass of that iteration); adds into dS
// do_a : MLP / GRU backward of iteration it_a, consuming dS (= d slot_out of it_a) -> dU, dS (= d slot_in via the GRU state)
// order inside one launch: do_q first (it completes d slot_in of iteration it_q = d slot_out of it_a = it_q - 1)
__global__ void __launch_bounds__(kBSlotThreads) sa_bwd_slot_kernel(const SabParams p, int do_q, int it_q, int do_a, int it_a, int first,
                                                                          int staged) {
  __shared__ float s_a[kBMaxC], s_b[kBMaxC], s_c[kBMaxC], s_d[kBMaxC], s_e[kBMaxC], s_f[kBMaxC], s_xh[kBMaxC], s_ds[kBMaxC];
  __shared__ float s_g[6 * kBMaxC], s_dg[6 * kBMaxC], s_m[kBMaxHid], s_dm[kBMaxHid];
  __shared__ float s_stat[2];
  const int C = p.C, K = p.K, Hd = p.hidden, k = blockIdx.x;
  const SlotRegion R = slot_region(C, Hd);
  float* G = p.slot_reg + (long long)k * R.total;
  const SabVec VO = sab_vec(C, Hd);
  const bool bg = k == 0;
  // The kernel is a chain of ~10 dependent matvec phases over the slot's weights: with `staged` they are pulled into shared
  // memory once, by TMA bulk copies that run while the first phases read their inputs (as in the forward update kernel).
  extern __shared__ __align__(128) float s_dyn[];
  __shared__ uint64_t s_bar;
  const float* w_ih = bg ? p.w.gru_bg_w_ih : p.w.gru_w_ih;
  const float* w_hh = bg ? p.w.gru_bg_w_hh : p.w.gru_w_hh;
  const float* w_1 = bg ? p.w.to_res_bg_w1 : p.w.to_res_w1;
  const float* w_2 = bg ? p.w.to_res_bg_w2 : p.w.to_res_w2;
  const float* w_q = bg ? p.w.to_q_bg_w : p.w.to_q_w;
  if (staged) {
    float* d_q = s_dyn;
    float* d_ih = d_q + C * C;
    float* d_hh = d_ih + 3 * C * C;
    float* d_1 = d_hh + 3 * C * C;
    float* d_2 = d_1 + Hd * C;
    if (threadIdx.x == 0) {
      mbar_init(&s_bar, 1);
      fence_mbar_init();
      const uint32_t bytes = (uint32_t)(((do_q ? C * C : 0) + (do_a ? 2 * 3 * C * C + 2 * Hd * C : 0)) * sizeof(float));
      mbar_arrive_expect_tx(&s_bar, bytes);
      if (do_q) bulk_g2s(d_q, w_q, C * C * 4, &s_bar);
      if (do_a) {
        bulk_g2s(d_ih, w_ih, 3 * C * C * 4, &s_bar);
        bulk_g2s(d_hh, w_hh, 3 * C * C * 4, &s_bar);
        bulk_g2s(d_1, w_1, Hd * C * 4, &s_bar);
        bulk_g2s(d_2, w_2, C * Hd * 4, &s_bar);
      }
    }
    w_q = d_q; w_ih = d_ih; w_hh = d_hh; w_1 = d_1; w_2 = d_2;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_ds[c] = first ? p.grad_slots[k * C + c] : p.dS[k * C + c];
  __syncthreads();                                   // also publishes the mbarrier initialisation
  bool landed = !staged;                             // block-uniform: the first consumer of a staged weight waits once

  if (do_q) {
    const float* sv = p.saved + (long long)it_q * sab_saved_stride(K, C);
    // d q = sum of the token CTAs' partials (fixed order)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float acc = 0.f;
#pragma unroll 8
      for (int b = 0; b < p.nblk; ++b) acc += p.dq_part[((long long)b * K + k) * C + c];
      s_a[c] = acc;               // d q
      s_b[c] = sv[k * C + c];     // slot_in
    }
    __syncthreads();
    b_ln_fwd(s_b, bg ? p.w.to_q_bg_ln_w : p.w.to_q_ln_w, bg ? p.w.to_q_bg_ln_b : p.w.to_q_ln_b, s_xh, s_c, s_stat, C);   // s_c = LN_q(slot)
    __syncthreads();
    {                                                                        // factors of d to_q.1.weight
      float* V = p.vec_store + ((long long)it_q * K + k) * VO.total;
      for (int c = threadIdx.x; c < C; c += blockDim.x) { V[VO.q_a + c] = s_a[c]; V[VO.q_b + c] = s_c[c]; }
    }
    if (!landed) { mbar_wait(&s_bar, 0); landed = true; }
    b_matvec_t(w_q, s_a, s_d, C, C);                                         // s_d = d LN_q output
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      G[R.lnq_w + c] += s_d[c] * s_xh[c];
      G[R.lnq_b + c] += s_d[c];
    }
    b_ln_bwd(s_d, s_xh, bg ? p.w.to_q_bg_ln_w : p.w.to_q_ln_w, s_stat[0], s_e, C);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_ds[c] += s_e[c];
    __syncthreads();
  }

  if (do_a) {
    const float* sv = p.saved + (long long)it_a * sab_saved_stride(K, C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      s_a[c] = sv[k * C + c];               // h = slot_in
      s_b[c] = sv[2 * K * C + k * C + c];   // upd
    }
    __syncthreads();
    // ---- recompute the forward of this iteration for slot k ----
    if (!landed) { mbar_wait(&s_bar, 0); landed = true; }
    b_matvec(w_ih, bg ? p.w.gru_bg_b_ih : p.w.gru_b_ih, s_b, s_g, 3 * C, C);
    b_matvec(w_hh, bg ? p.w.gru_bg_b_hh : p.w.gru_b_hh, s_a, s_g + 3 * C, 3 * C, C);
    __syncthreads();
    // s_c = r, s_d = z, s_e = n, s_f = h'
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float r = 1.f / (1.f + expf(-(s_g[c] + s_g[3 * C + c])));
      const float z = 1.f / (1.f + expf(-(s_g[C + c] + s_g[4 * C + c])));
      const float n = tanhf(s_g[2 * C + c] + r * s_g[5 * C + c]);
      s_c[c] = r; s_d[c] = z; s_e[c] = n;
      s_f[c] = (1.f - z) * n + z * s_a[c];
    }
    __syncthreads();
    float* s_ln = s_dg;  // reuse: LN_res output (C floats) lives in s_dg[0..C) until the GRU gate gradients are formed
    b_ln_fwd(s_f, bg ? p.w.to_res_bg_ln_w : p.w.to_res_ln_w, bg ? p.w.to_res_bg_ln_b : p.w.to_res_ln_b, s_xh, s_ln, s_stat, C);
    __syncthreads();
    b_matvec(w_1, bg ? p.w.to_res_bg_b1 : p.w.to_res_b1, s_ln, s_m, Hd, C);
    __syncthreads();
    for (int j = threadIdx.x; j < Hd; j += blockDim.x) s_m[j] = fmaxf(s_m[j], 0.f);
    __syncthreads();
    // ---- residual MLP backward: slot_out = h' + W2 relu(W1 LN(h') + b1) + b2 ----
    float* V = p.vec_store + ((long long)it_a * K + k) * VO.total;           // factors of this iteration's weight gradients
    for (int c = threadIdx.x; c < C; c += blockDim.x) { V[VO.ds + c] = s_ds[c]; V[VO.ln + c] = s_ln[c]; V[VO.upd + c] = s_b[c]; V[VO.h + c] = s_a[c]; }
    for (int j = threadIdx.x; j < Hd; j += blockDim.x) V[VO.m + j] = s_m[j];
    b_vec_acc(G + R.b2, s_ds, C);
    b_matvec_t(w_2, s_ds, s_dm, C, Hd);                                      // d relu output
    __syncthreads();
    for (int j = threadIdx.x; j < Hd; j += blockDim.x) s_dm[j] = s_m[j] > 0.f ? s_dm[j] : 0.f;
    __syncthreads();
    for (int j = threadIdx.x; j < Hd; j += blockDim.x) V[VO.dm + j] = s_dm[j];
    b_vec_acc(G + R.b1, s_dm, Hd);
    float* s_dln = s_g;   // reuse s_g[0..C): the gate pre-activations are no longer needed (r, z, n are in s_c/s_d/s_e)
    // keep h_n = W_hn h + b_hn, needed for d r: copy before s_g is overwritten
    __shared__ float s_hn[kBMaxC];
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_hn[c] = s_g[5 * C + c];
    __syncthreads();
    b_matvec_t(w_1, s_dm, s_dln, Hd, C);                                     // d LN_res output
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      G[R.lnr_w + c] += s_dln[c] * s_xh[c];
      G[R.lnr_b + c] += s_dln[c];
    }
    float* s_dx = s_g + kBMaxC;
    b_ln_bwd(s_dln, s_xh, bg ? p.w.to_res_bg_ln_w : p.w.to_res_ln_w, s_stat[0], s_dx, C);
    __syncthreads();
    // ---- GRUCell backward: h' = (1-z) n + z h ----
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float dh2 = s_ds[c] + s_dx[c];            // d h'
      const float r = s_c[c], z = s_d[c], n = s_e[c], h = s_a[c];
      const float dn = dh2 * (1.f - z);
      const float dz = dh2 * (h - n);
      const float dpn = dn * (1.f - n * n);
      const float dpz = dz * z * (1.f - z);
      const float dr = dpn * s_hn[c];
      const float dpr = dr * r * (1.f - r);
      s_dg[c] = dpr; s_dg[C + c] = dpz; s_dg[2 * C + c] = dpn;                 // d gi
      s_dg[3 * C + c] = dpr; s_dg[4 * C + c] = dpz; s_dg[5 * C + c] = dpn * r;  // d gh
      s_f[c] = dh2 * z;                                // direct path to h
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 6 * C; j += blockDim.x) V[VO.dg + j] = s_dg[j];
    b_vec_acc(G + R.bih, s_dg, 3 * C);
    b_vec_acc(G + R.bhh, s_dg + 3 * C, 3 * C);
    b_matvec_t(w_ih, s_dg, s_xh, 3 * C, C);             // s_xh = d upd
    b_matvec_t(w_hh, s_dg + 3 * C, s_hn, 3 * C, C);     // s_hn = W_hh^T d gh
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      p.dU[k * C + c] = s_xh[c];
      s_ds[c] = s_f[c] + s_hn[c];     // d slot_in through the GRU state (the query path is added by the next launch)
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) p.dS[k * C + c] = s_ds[c];
}

// ---- token-parallel kernel of iteration `it` ---------------------------------------------------------
__global__ void __launch_bounds__(kBThreads) sa_bwd_token_kernel(const SabParams p, int it) {
  __shared__ float s_q[kBMaxK][kBMaxC + 1], s_du[kBMaxK][kBMaxC + 1];
  __shared__ float s_k[kBTok][kBMaxC + 1], s_v[kBTok][kBMaxC + 1];
  __shared__ float s_p[kBMaxK][kBTok];    // softmax over slots
  __shared__ float s_dd[kBMaxK][kBTok];   // d dots (incl. scale)
  __shared__ float s_w[kBMaxK][kBTok];    // normalised attention weights
  __shared__ float s_asum[kBMaxK], s_duu[kBMaxK];
  const int C = p.C, K = p.K;
  const int t0 = blockIdx.x * kBTok;
  const float* sv = p.saved + (long long)it * sab_saved_stride(K, C);
  for (int e = threadIdx.x; e < kBTok * C; e += kBThreads) {
    const int t = e / C, c = e % C;
    const bool in = t0 + t < p.ntok;
    s_k[t][c] = in ? p.kbuf[(long long)(t0 + t) * C + c] : 0.f;
    s_v[t][c] = in ? p.vbuf[(long long)(t0 + t) * C + c] : 0.f;
  }
  for (int e = threadIdx.x; e < K * C; e += kBThreads) {
    s_q[e / C][e % C] = sv[K * C + e];
    s_du[e / C][e % C] = p.dU[e];
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    s_asum[k] = sv[3 * K * C + k];
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(s_du[k][c], sv[2 * K * C + k * C + c], acc);   // d upd_k . upd_k
    s_duu[k] = acc;
  }
  const float scale = rsqrtf((float)C);
  for (int e = threadIdx.x; e < K * kBTok; e += kBThreads) {
    const int k = e / kBTok, t = e % kBTok;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(s_q[k][c], s_k[t][c], acc);
    s_p[k][t] = acc * scale;
  }
  __syncthreads();
  if (threadIdx.x < kBTok) {
    const int t = threadIdx.x;
    const bool in = t0 + t < p.ntok;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, s_p[k][t]);
    float den = 0.f;
    for (int k = 0; k < K; ++k) { const float e = expf(s_p[k][t] - m); s_p[k][t] = e; den += e; }
    float sum_pda = 0.f;
    for (int k = 0; k < K; ++k) {
      const float pk = s_p[k][t] / den;
      s_p[k][t] = pk;
      float dw = 0.f;
      for (int c = 0; c < C; ++c) dw = fmaf(s_du[k][c], s_v[t][c], dw);     // d w[k,t] = d upd_k . v_t
      const float da = in ? (dw - s_duu[k]) / s_asum[k] : 0.f;              // d attn[k,t]
      s_dd[k][t] = da;
      s_w[k][t] = in ? (pk + p.eps) / s_asum[k] : 0.f;
      sum_pda += pk * da;
    }
    for (int k = 0; k < K; ++k) s_dd[k][t] = in ? s_p[k][t] * (s_dd[k][t] - sum_pda) * scale : 0.f;   // d (q.k)
  }
  __syncthreads();
  // d k_t, d v_t (accumulated over the iterations; this CTA owns its tokens)
  for (int e = threadIdx.x; e < kBTok * C; e += kBThreads) {
    const int t = e / C, c = e % C;
    if (t0 + t >= p.ntok) continue;
    float dk = 0.f, dv = 0.f;
    for (int k = 0; k < K; ++k) { dk = fmaf(s_dd[k][t], s_q[k][c], dk); dv = fmaf(s_w[k][t], s_du[k][c], dv); }
    p.dK[(long long)(t0 + t) * C + c] += dk;
    p.dV[(long long)(t0 + t) * C + c] += dv;
  }
  // partial d q
  float* part = p.dq_part + (long long)blockIdx.x * K * C;
  for (int e = threadIdx.x; e < K * C; e += kBThreads) {
    const int k = e / C, c = e % C;
    float acc = 0.f;
    for (int t = 0; t < kBTok; ++t) acc = fmaf(s_dd[k][t], s_k[t][c], acc);
    part[e] = acc;
  }
}

// ---- k / v projections and LayerNorm(feat) backward ----------------------------------------------------
__global__ void __launch_bounds__(kBThreads) sa_bwd_feat_kernel(const SabParams p) {
  __shared__ float s_x[kBTok][kBMaxC + 1];    // feat, then xhat
  __shared__ float s_f[kBTok][kBMaxC + 1];    // LN(feat)
  __shared__ float s_dk[kBTok][kBMaxC + 1], s_dv[kBTok][kBMaxC + 1];
  __shared__ float s_df[kBTok][kBMaxC + 1];
  __shared__ float s_rstd[kBTok];
  const int C = p.C;
  const int t0 = blockIdx.x * kBTok;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < kBTok * C; e += kBThreads) {
    const int c = e / kBTok, t = e % kBTok;
    const bool in = t0 + t < p.ntok;
    s_x[t][c] = in ? __ldg(p.feat + (long long)(t0 + t) * p.feat_ts + (long long)c * p.feat_cs) : 0.f;
  }
  for (int e = threadIdx.x; e < kBTok * C; e += kBThreads) {
    const int t = e / C, c = e % C;
    const bool in = t0 + t < p.ntok;
    s_dk[t][c] = in ? p.dK[(long long)(t0 + t) * C + c] : 0.f;
    s_dv[t][c] = in ? p.dV[(long long)(t0 + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int t = warp; t < kBTok; t += kBThreads / 32) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += s_x[t][c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = s_x[t][c] - mean; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + 1e-5f);
    for (int c = lane; c < C; c += 32) {
      const float xh = (s_x[t][c] - mean) * rstd;
      s_x[t][c] = xh;
      s_f[t][c] = xh * p.w.norm_feat_w[c] + p.w.norm_feat_b[c];
    }
    if (lane == 0) s_rstd[t] = rstd;
  }
  __syncthreads();
  // d f[t][c] = sum_o dK[t][o] Wk[o][c] + dV[t][o] Wv[o][c]
  for (int e = threadIdx.x; e < kBTok * C; e += kBThreads) {
    const int t = e / C, c = e % C;
    float acc = 0.f;
    for (int o = 0; o < C; ++o) acc = fmaf(s_dk[t][o], __ldg(p.w.to_k_w + o * C + c), fmaf(s_dv[t][o], __ldg(p.w.to_v_w + o * C + c), acc));
    s_df[t][c] = acc;
  }
  __syncthreads();
  // partial parameter gradients of this CTA
  float* part = p.tok_part + (long long)blockIdx.x * (2 * C * C + 2 * C);
  for (int e = threadIdx.x; e < C * C; e += kBThreads) {
    const int o = e / C, c = e % C;
    float ak = 0.f, av = 0.f;
    for (int t = 0; t < kBTok; ++t) { ak = fmaf(s_dk[t][o], s_f[t][c], ak); av = fmaf(s_dv[t][o], s_f[t][c], av); }
    part[e] = ak;
    part[C * C + e] = av;
  }
  for (int c = threadIdx.x; c < C; c += kBThreads) {
    float gw = 0.f, gb = 0.f;
    for (int t = 0; t < kBTok; ++t) { gw = fmaf(s_df[t][c], s_x[t][c], gw); gb += s_df[t][c]; }
    part[2 * C * C + c] = gw;
    part[2 * C * C + C + c] = gb;
  }
  // LayerNorm backward -> d feat
  for (int t = warp; t < kBTok; t += kBThreads / 32) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = s_df[t][c] * p.w.norm_feat_w[c]; s1 += d; s2 += d * s_x[t][c]; }
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    s1 /= C; s2 /= C;
    if (t0 + t < p.ntok)
      for (int c = lane; c < C; c += 32)
        p.grad_feat[(long long)(t0 + t) * C + c] = s_rstd[t] * (s_df[t][c] * p.w.norm_feat_w[c] - s1 - s_x[t][c] * s2);
  }
}

// ---- fixed-order reductions into the parameter-gradient tensors ----------------------------------------------
__global__ void sa_bwd_reduce_kernel(const SabParams p, int staged) {
  const int C = p.C, K = p.K, Hd = p.hidden;
  const SlotRegion R = slot_region(C, Hd);
  const int gid = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
  auto fg_sum = [&](int off) {
    float a = 0.f;
#pragma unroll 4
    for (int k = 1; k < K; ++k) a += p.slot_reg[(long long)k * R.total + off];
    return a;
  };
  auto bg_val = [&](int off) { return p.slot_reg[off]; };
  // weight element (o, c) of slot k: sum over the iterations (last first, the order the backward runs) of a[o] * b[c]
  const SabVec VO = sab_vec(C, Hd);
  // the factor vectors (iters x K x (12 C + 2 hidden) floats: 96 KB at the reference sizes) are read ~40 times per weight element:
  // staged in shared memory when they fit (staged != 0)
  extern __shared__ __align__(16) float s_vec[];
  const float* vec = p.vec_store;
  if (staged) {
    const int n4 = p.iters * K * VO.total / 4;              // C, hidden multiples of 4 (checked by the launcher)
    for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<float4*>(s_vec)[i] = reinterpret_cast<const float4*>(p.vec_store)[i];
    __syncthreads();
    vec = s_vec;
  }
  auto outer = [&](int off_a, int off_b, int o, int c, int k) {
    float acc = 0.f;
    for (int it = p.iters - 1; it >= 0; --it) {
      const float* V = vec + (it * K + k) * VO.total;
      acc = fmaf(V[off_a + o], V[off_b + c], acc);
    }
    return acc;
  };
  auto outer_fg = [&](int off_a, int off_b, int o, int c) {
    float a = 0.f;
    for (int k = 1; k < K; ++k) a += outer(off_a, off_b, o, c, k);
    return a;
  };
  for (int e = gid; e < 3 * C * C; e += gn) {
    const int o = e / C, c = e - o * C;
    p.g.gru_w_ih[e] = outer_fg(VO.dg, VO.upd, o, c); p.g.gru_w_hh[e] = outer_fg(VO.dg + 3 * C, VO.h, o, c);
    p.g.gru_bg_w_ih[e] = outer(VO.dg, VO.upd, o, c, 0); p.g.gru_bg_w_hh[e] = outer(VO.dg + 3 * C, VO.h, o, c, 0);
  }
  for (int e = gid; e < 3 * C; e += gn) { p.g.gru_b_ih[e] = fg_sum(R.bih + e); p.g.gru_b_hh[e] = fg_sum(R.bhh + e);
                                            p.g.gru_bg_b_ih[e] = bg_val(R.bih + e); p.g.gru_bg_b_hh[e] = bg_val(R.bhh + e); }
  for (int e = gid; e < Hd * C; e += gn) {
    const int j = e / C, c = e - j * C;          // first weight [hidden][C]
    p.g.to_res_w1[e] = outer_fg(VO.dm, VO.ln, j, c); p.g.to_res_bg_w1[e] = outer(VO.dm, VO.ln, j, c, 0);
    const int o = e / Hd, jj = e - o * Hd;       // second weight [C][hidden]
    p.g.to_res_w2[e] = outer_fg(VO.ds, VO.m, o, jj); p.g.to_res_bg_w2[e] = outer(VO.ds, VO.m, o, jj, 0);
  }
  for (int e = gid; e < Hd; e += gn) { p.g.to_res_b1[e] = fg_sum(R.b1 + e); p.g.to_res_bg_b1[e] = bg_val(R.b1 + e); }
  for (int e = gid; e < C * C; e += gn) {
    const int o = e / C, c = e - o * C;
    p.g.to_q_w[e] = outer_fg(VO.q_a, VO.q_b, o, c); p.g.to_q_bg_w[e] = outer(VO.q_a, VO.q_b, o, c, 0);
  }
  for (int e = gid; e < C; e += gn) {
    p.g.to_res_b2[e] = fg_sum(R.b2 + e); p.g.to_res_bg_b2[e] = bg_val(R.b2 + e);
    p.g.to_res_ln_w[e] = fg_sum(R.lnr_w + e); p.g.to_res_ln_b[e] = fg_sum(R.lnr_b + e);
    p.g.to_res_bg_ln_w[e] = bg_val(R.lnr_w + e); p.g.to_res_bg_ln_b[e] = bg_val(R.lnr_b + e);
    p.g.to_q_ln_w[e] = fg_sum(R.lnq_w + e); p.g.to_q_ln_b[e] = fg_sum(R.lnq_b + e);
    p.g.to_q_bg_ln_w[e] = bg_val(R.lnq_w + e); p.g.to_q_bg_ln_b[e] = bg_val(R.lnq_b + e);
    // slot initialisation: slot = mu + exp(logsigma) * noise   (model.py:214-219); dS now holds d slot_0
    float dmu = 0.f, dls = 0.f;
    for (int k = 1; k < K; ++k) { const float d = p.dS[k * C + e]; dmu += d; dls += d * expf(p.w.slots_logsigma[e]) * p.noise_fg[(k - 1) * C + e]; }
    p.g.slots_mu[e] = dmu; p.g.slots_logsigma[e] = dls;
    p.g.slots_mu_bg[e] = p.dS[e];
    p.g.slots_logsigma_bg[e] = p.dS[e] * expf(p.w.slots_logsigma_bg[e]) * p.noise_bg[e];
  }
  // token-pass partials [nblk][2 C C + 2 C] -> to_k_w | to_v_w | norm_feat_w | norm_feat_b (the same order).  A block takes 32
  // consecutive elements; warp j adds blocks j, j + 8, ... in order (coalesced rows), warp 0 then adds the eight sums in order:
  // bit-reproducible.  (One thread per element walked all 128 partial blocks alone: 4096 busy threads, 128 dependent round trips.)
  __shared__ float s_part[8][32];
  const int row = 2 * C * C + 2 * C;
  const int lane = threadIdx.x & 31, wsub = threadIdx.x >> 5;       // launched with 256 threads
  for (int e0 = blockIdx.x * 32; e0 < row; e0 += gridDim.x * 32) {
    const int e = e0 + lane;
    float a = 0.f;
    if (e < row) {
#pragma unroll 4
      for (int b = wsub; b < p.nblk; b += 8) a += p.tok_part[(long long)b * row + e];
    }
    s_part[wsub][lane] = a;
    __syncthreads();
    if (wsub == 0 && e < row) {
      float t = s_part[0][lane];
#pragma unroll
      for (int j = 1; j < 8; ++j) t += s_part[j][lane];
      if (e < C * C) p.g.to_k_w[e] = t;
      else if (e < 2 * C * C) p.g.to_v_w[e - C * C] = t;
      else if (e < 2 * C * C + C) p.g.norm_feat_w[e - 2 * C * C] = t;
      else p.g.norm_feat_b[e - 2 * C * C - C] = t;
    }
    __syncthreads();
  }
}

size_t slot_attention_bwd_ws_floats(int ntok, int C, int K, int hidden) {
  const size_t nblk = (ntok + kBTok - 1) / kBTok;
  const SlotRegion R = slot_region(C, hidden);
  return (size_t)2 * K * C + (size_t)2 * ntok * C + nblk * K * C + (size_t)K * R.total + nblk * (2 * C * C + 2 * C) +
         (size_t)kBMaxIters * K * sab_vec(C, hidden).total + 64;
}

int launch_slot_attention_bwd(const uorf_slot_attention_weights* w, const uorf_slot_attention_grads* g, const float* feat, long long feat_ts,
                              long long feat_cs, const float* noise_fg, const float* noise_bg, int ntok, int C, int K, int hidden, int iters,
                              float eps, const float* fwd_workspace, const float* grad_slots, float* workspace, float* grad_feat,
                              cudaStream_t stream) {
  if (C > kBMaxC || K > kBMaxK || hidden > kBMaxHid || K < 2 || iters > kBMaxIters || iters < 1) return UORF_ERR_UNSUPPORTED;
  SabParams p;
  p.w = *w; p.g = *g;
  p.feat = feat; p.feat_ts = feat_ts; p.feat_cs = feat_cs;
  p.noise_fg = noise_fg; p.noise_bg = noise_bg;
  p.ntok = ntok; p.C = C; p.K = K; p.hidden = hidden; p.iters = iters; p.eps = eps;
  const int nblk = (ntok + kBTok - 1) / kBTok;
  p.nblk = nblk;
  // forward workspace layout (slot_attention.cu): k | v | partials [nblk][K][C+1] | q [K][C] | saved
  p.kbuf = fwd_workspace;
  p.vbuf = fwd_workspace + (size_t)ntok * C;
  p.saved = fwd_workspace + (size_t)2 * ntok * C + (size_t)nblk * K * (C + 1) + (size_t)K * C;
  p.grad_slots = grad_slots;
  const SlotRegion R = slot_region(C, hidden);
  float* ws = workspace;
  p.dS = ws; ws += K * C;
  p.dU = ws; ws += K * C;
  p.dK = ws; ws += (size_t)ntok * C;
  p.dV = ws; ws += (size_t)ntok * C;
  p.dq_part = ws; ws += (size_t)nblk * K * C;
  p.slot_reg = ws; ws += (size_t)K * R.total;
  p.tok_part = ws; ws += (size_t)nblk * (2 * C * C + 2 * C);
  p.vec_store = ws;
  p.grad_feat = grad_feat;
  // weight staging of the per-slot kernel (TMA bulk copies: 16-byte aligned tensors, sizes multiples of 16 bytes)
  const size_t stage_bytes = (size_t)(C * C + 2 * 3 * C * C + 2 * hidden * C) * sizeof(float);
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int staged = (stage_bytes <= 200 * 1024 && C % 4 == 0 && hidden % 4 == 0 && al(w->gru_w_ih) && al(w->gru_w_hh) && al(w->gru_bg_w_ih) && al(w->gru_bg_w_hh) &&
                      al(w->to_res_w1) && al(w->to_res_w2) && al(w->to_res_bg_w1) && al(w->to_res_bg_w2) && al(w->to_q_w) && al(w->to_q_bg_w)) ? 1 : 0;
  // set on every launch, like the decoder kernels: function attributes are per device, a process may drive several
  if (staged && cudaFuncSetAttribute(sa_bwd_slot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return UORF_ERR_DEVICE;
  const size_t dyn = staged ? stage_bytes : 0;
  cudaMemsetAsync(p.dK, 0, sizeof(float) * 2 * (size_t)ntok * C, stream);
  cudaMemsetAsync(p.slot_reg, 0, sizeof(float) * (size_t)K * R.total, stream);
  for (int it = iters - 1; it >= 0; --it) {
    const bool last = it == iters - 1;
    // query path of iteration it+1 (if any) + MLP/GRU of iteration it
    sa_bwd_slot_kernel<<<K, kBSlotThreads, dyn, stream>>>(p, last ? 0 : 1, it + 1, 1, it, last ? 1 : 0, staged);
    sa_bwd_token_kernel<<<nblk, kBThreads, 0, stream>>>(p, it);
  }
  sa_bwd_slot_kernel<<<K, kBSlotThreads, dyn, stream>>>(p, 1, 0, 0, 0, 0, staged);   // query path of iteration 0 -> d slot_0
  sa_bwd_feat_kernel<<<nblk, kBThreads, 0, stream>>>(p);
  const size_t vec_bytes = (size_t)iters * K * sab_vec(C, hidden).total * sizeof(float);
  const int red_staged = (vec_bytes <= 160 * 1024 && C % 4 == 0 && hidden % 4 == 0 && (reinterpret_cast<uintptr_t>(p.vec_store) & 15) == 0) ? 1 : 0;
  if (red_staged && cudaFuncSetAttribute(sa_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) return UORF_ERR_DEVICE;
  sa_bwd_reduce_kernel<<<148, 256, red_staged ? vec_bytes : 0, stream>>>(p, red_staged);
  return finish_launch(2 * iters + 3);
}

}  // namespace uorf
