// Kernel 4 — background-aware slot attention (forward), batch 1.
// Replaces SlotAttention.forward, reference models/model.py:205-258:
//   slots_fg = mu + exp(logsigma) * noise_fg ; slot_bg likewise                      (:214-219)
//   k, v = to_k(LN(feat)), to_v(LN(feat))                                             (:221-223)
//   iters x { q = to_q(LN(slot)) (separate fg / bg) ; dots = q.k * C^-0.5 ;           (:229-233)
//             attn = softmax over SLOTS + eps ; per-slot renormalise over tokens ;     (:235-238)
//             updates = attn_w . v ; slot = GRUCell(updates, slot) ; slot += MLP(LN(slot)) }  (:240-255)
// The work is ~0.1 GFLOP and latency-bound, so it is fp32 FFMA, organised to need only two kinds of
// launch per iteration: a token-parallel pass over all SMs (dots, softmax, partial weighted sums) and a
// one-CTA-per-slot update (reduction of the partials, GRU, residual MLP, next query).  Reductions run in a fixed order,
// so results are deterministic.
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kSaTok = 32;      // tokens per CTA in the token-parallel passes
constexpr int kSaThreads = 256;
constexpr int kSaSlotThreads = 1024;   // per-slot CTAs: more warps = more weight rows in flight (latency-bound matvecs)
constexpr int kSaMaxC = 64;
constexpr int kSaMaxK = 16;
constexpr int kSaMaxHid = 128;
__host__ __device__ inline int sa_saved_stride(int K, int C) { return 3 * K * C + ((K + 3) & ~3); }

struct SaParams {
  uorf_slot_attention_weights w;
  const float* feat;
  long long feat_ts, feat_cs;  // address of (token t, channel c) = feat + t*ts + c*cs
  const float* noise_fg;
  const float* noise_bg;
  int ntok, C, K, hidden, iters;
  float eps;
  float* kbuf;     // [ntok][C]
  float* vbuf;     // [ntok][C]
  float* partial;  // [nblk][K][C+1]
  float* qbuf;     // [K][C] queries of the current iteration
  float* saved;    // per iteration: slot_in [K][C] | q [K][C] | upd [K][C] | asum [K]   (re-read by the backward kernels)
  float* slots;    // [K][C] state (slot 0 = bg)
  float* attn;     // [K][ntok]
};

// ---- pass A: k, v projections --------------------------------------------------------------------
__global__ void __launch_bounds__(kSaThreads) sa_kv_kernel(const SaParams p) {
  __shared__ float s_f[kSaTok][kSaMaxC + 1];
  __shared__ float s_w[2][kSaMaxC][kSaMaxC + 1];
  const int C = p.C;
  const int t0 = blockIdx.x * kSaTok;
  for (int e = threadIdx.x; e < C * C; e += kSaThreads) {
    s_w[0][e / C][e % C] = p.w.to_k_w[e];
    s_w[1][e / C][e % C] = p.w.to_v_w[e];
  }
  for (int e = threadIdx.x; e < kSaTok * C; e += kSaThreads) {
    const int c = e / kSaTok, t = e % kSaTok;  // consecutive threads -> consecutive tokens (unit stride for NCHW input)
    s_f[t][c] = (t0 + t < p.ntok) ? __ldg(p.feat + (long long)(t0 + t) * p.feat_ts + (long long)c * p.feat_cs) : 0.f;
  }
  __syncthreads();
  // LayerNorm over channels: one warp handles 4 tokens
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < kSaTok; t += kSaThreads / 32) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += s_f[t][c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = s_f[t][c] - mean; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + 1e-5f);
    for (int c = lane; c < C; c += 32) s_f[t][c] = (s_f[t][c] - mean) * rstd * p.w.norm_feat_w[c] + p.w.norm_feat_b[c];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kSaTok * C * 2; e += kSaThreads) {
    const int which = e / (kSaTok * C);
    const int r = e % (kSaTok * C);
    const int t = r / C, o = r % C;
    if (t0 + t >= p.ntok) continue;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(s_f[t][c], s_w[which][o][c], acc);
    (which ? p.vbuf : p.kbuf)[(long long)(t0 + t) * C + o] = acc;
  }
}

// ---- helpers for the per-slot CTAs (256 threads = 8 warps) -----------------------------------------
// y[o] = bias[o] + sum_c W[o][c] * x[c] for o < n_out: one warp per output row, lanes stride over c, so the
// weight rows are read with unit stride (coalesced) and reduced with shuffles.
__device__ __forceinline__ void warp_matvec(const float* W, const float* bias, const float* x, float* y,
                                            int n_out, int n_in, bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // kRows rows of a warp are loaded before any of them is reduced: the weight reads (L2 latency ~1 us on a cold line) of a
  // warp's rows overlap instead of forming one round trip per row
  constexpr int kRows = 6, kPer = kSaMaxHid / 32;
  for (int o0 = warp; o0 < n_out; o0 += nw * kRows) {
    float w[kRows][kPer];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int o = o0 + r * nw;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int c = lane + 32 * j;
        w[r][j] = (o < n_out && c < n_in) ? W[(long long)o * n_in + c] : 0.f;   // generic load: global or staged shared memory
      }
    }
    float xv[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) xv[j] = lane + 32 * j < n_in ? x[lane + 32 * j] : 0.f;
    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      acc[r] = 0.f;
#pragma unroll
      for (int j = 0; j < kPer; ++j) acc[r] = fmaf(w[r][j], xv[j], acc[r]);   // same order as a lane-strided loop over c
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1)
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int o = o0 + r * nw;
        if (o < n_out) {
          const float a = acc[r] + (bias ? bias[o] : 0.f);
          y[o] = relu ? fmaxf(a, 0.f) : a;
        }
      }
    }
  }
}
// LayerNorm of x[0..C) by warp 0 -> y
__device__ __forceinline__ void warp0_layernorm(const float* x, const float* __restrict__ w, const float* __restrict__ b, float* y, int C) {
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
    for (int c = lane; c < C; c += 32) y[c] = (x[c] - mean) * rstd * w[c] + b[c];
  }
}
// q_k = to_q(LN(slot_k)) -> global q buffer
__device__ __forceinline__ void slot_query(const SaParams& p, int k, const float* slot, float* s_ln, float* s_q, float* q_save,
                                           const float* wq_staged = nullptr) {
  warp0_layernorm(slot, k == 0 ? p.w.to_q_bg_ln_w : p.w.to_q_ln_w, k == 0 ? p.w.to_q_bg_ln_b : p.w.to_q_ln_b, s_ln, p.C);
  __syncthreads();
  warp_matvec(wq_staged ? wq_staged : (k == 0 ? p.w.to_q_bg_w : p.w.to_q_w), nullptr, s_ln, s_q, p.C, p.C, false);
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    p.qbuf[k * p.C + c] = s_q[c];
    if (q_save) q_save[k * p.C + c] = s_q[c];
  }
}

// ---- slot initialisation + first query; one CTA per slot -------------------------------------------
__global__ void __launch_bounds__(kSaSlotThreads) sa_init_kernel(const SaParams p) {
  __shared__ float s_slot[kSaMaxC], s_ln[kSaMaxC], s_q[kSaMaxC];
  const int k = blockIdx.x, C = p.C, K = p.K;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = k == 0 ? p.w.slots_mu_bg[c] + expf(p.w.slots_logsigma_bg[c]) * p.noise_bg[c]
                           : p.w.slots_mu[c] + expf(p.w.slots_logsigma[c]) * p.noise_fg[(k - 1) * C + c];
    s_slot[c] = v;
    p.slots[k * C + c] = v;
    p.saved[k * C + c] = v;                       // slot_in of iteration 0
  }
  __syncthreads();
  slot_query(p, k, s_slot, s_ln, s_q, p.saved + K * C);
}

// ---- pass B (per iteration): dots, softmax over slots, partial weighted sums -------------------------
__global__ void __launch_bounds__(kSaThreads) sa_attn_kernel(const SaParams p) {
  __shared__ float s_q[kSaMaxK][kSaMaxC + 1];
  __shared__ float s_k[kSaTok][kSaMaxC + 1];
  __shared__ float s_v[kSaTok][kSaMaxC + 1];
  __shared__ float s_a[kSaMaxK][kSaTok];
  const int C = p.C, K = p.K;
  const int t0 = blockIdx.x * kSaTok;
  for (int e = threadIdx.x; e < kSaTok * C; e += kSaThreads) {
    const int t = e / C, c = e % C;
    const bool in = t0 + t < p.ntok;
    s_k[t][c] = in ? p.kbuf[(long long)(t0 + t) * C + c] : 0.f;
    s_v[t][c] = in ? p.vbuf[(long long)(t0 + t) * C + c] : 0.f;
  }
  for (int e = threadIdx.x; e < K * C; e += kSaThreads) s_q[e / C][e % C] = p.qbuf[e];
  __syncthreads();
  const float scale = rsqrtf((float)C);
  for (int e = threadIdx.x; e < K * kSaTok; e += kSaThreads) {
    const int k = e / kSaTok, t = e % kSaTok;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(s_q[k][c], s_k[t][c], acc);
    s_a[k][t] = acc * scale;
  }
  __syncthreads();
  if (threadIdx.x < kSaTok) {
    const int t = threadIdx.x;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, s_a[k][t]);
    float den = 0.f;
    for (int k = 0; k < K; ++k) { const float e = expf(s_a[k][t] - m); s_a[k][t] = e; den += e; }
    const bool in = t0 + t < p.ntok;
    for (int k = 0; k < K; ++k) {
      const float a = in ? s_a[k][t] / den + p.eps : 0.f;
      s_a[k][t] = a;
      if (in) p.attn[(long long)k * p.ntok + t0 + t] = a;
    }
  }
  __syncthreads();
  // partial[b][k][c] = sum_t a[k][t] v[t][c] ; partial[b][k][C] = sum_t a[k][t]
  float* part = p.partial + (long long)blockIdx.x * K * (C + 1);
  for (int e = threadIdx.x; e < K * (C + 1); e += kSaThreads) {
    const int k = e / (C + 1), c = e % (C + 1);
    float acc = 0.f;
    if (c < C) { for (int t = 0; t < kSaTok; ++t) acc = fmaf(s_a[k][t], s_v[t][c], acc); }
    else       { for (int t = 0; t < kSaTok; ++t) acc += s_a[k][t]; }
    part[e] = acc;
  }
}

// ---- pass C (per iteration): reduce partials, GRUCell, residual MLP, next query; one CTA per slot ------
// All of a slot's weights (GRU 2 x 3C x C, residual MLP 2 x Hd x C, next query C x C, their biases: <= 180 KB) are pulled
// into shared memory by five TMA bulk copies issued before the partial sums are reduced, so the six dependent matvec phases
// read shared memory instead of paying one L2 round trip each (`staged`; needs 16-byte aligned tensors and C, Hd % 4 == 0).
__global__ void __launch_bounds__(kSaSlotThreads) sa_update_kernel(const SaParams p, int nblk, int it, int staged) {
  extern __shared__ __align__(128) float s_dyn[];
  __shared__ uint64_t s_bar;
  __shared__ float s_u[kSaMaxC], s_h[kSaMaxC], s_g[6 * kSaMaxC], s_n[kSaMaxC], s_ln[kSaMaxC], s_m[kSaMaxHid], s_o[kSaMaxC];
  __shared__ float s_red[kSaSlotThreads / 32][kSaMaxC + 1];
  const int C = p.C, K = p.K, Hd = p.hidden, k = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool bg = k == 0;
  const float* w_ih = bg ? p.w.gru_bg_w_ih : p.w.gru_w_ih;
  const float* w_hh = bg ? p.w.gru_bg_w_hh : p.w.gru_w_hh;
  const float* w_1 = bg ? p.w.to_res_bg_w1 : p.w.to_res_w1;
  const float* w_2 = bg ? p.w.to_res_bg_w2 : p.w.to_res_w2;
  const float* w_q = bg ? p.w.to_q_bg_w : p.w.to_q_w;
  const float* b_ih = bg ? p.w.gru_bg_b_ih : p.w.gru_b_ih;
  const float* b_hh = bg ? p.w.gru_bg_b_hh : p.w.gru_b_hh;
  const float* b_1 = bg ? p.w.to_res_bg_b1 : p.w.to_res_b1;
  const float* b_2 = bg ? p.w.to_res_bg_b2 : p.w.to_res_b2;
  if (staged) {
    float* d_ih = s_dyn;
    float* d_hh = d_ih + 3 * C * C;
    float* d_1 = d_hh + 3 * C * C;
    float* d_2 = d_1 + Hd * C;
    float* d_q = d_2 + C * Hd;
    float* d_bih = d_q + C * C;
    float* d_bhh = d_bih + 3 * C;
    float* d_b1 = d_bhh + 3 * C;
    float* d_b2 = d_b1 + Hd;
    if (threadIdx.x == 0) {
      mbar_init(&s_bar, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(&s_bar, (uint32_t)((2 * 3 * C * C + 2 * Hd * C + C * C + 6 * C + Hd + C) * sizeof(float)));
      bulk_g2s(d_ih, w_ih, 3 * C * C * 4, &s_bar);
      bulk_g2s(d_hh, w_hh, 3 * C * C * 4, &s_bar);
      bulk_g2s(d_1, w_1, Hd * C * 4, &s_bar);
      bulk_g2s(d_2, w_2, C * Hd * 4, &s_bar);
      bulk_g2s(d_q, w_q, C * C * 4, &s_bar);
      bulk_g2s(d_bih, b_ih, 3 * C * 4, &s_bar);
      bulk_g2s(d_bhh, b_hh, 3 * C * 4, &s_bar);
      bulk_g2s(d_b1, b_1, Hd * 4, &s_bar);
      bulk_g2s(d_b2, b_2, C * 4, &s_bar);
    }
    w_ih = d_ih; w_hh = d_hh; w_1 = d_1; w_2 = d_2; w_q = d_q; b_ih = d_bih; b_hh = d_bhh; b_1 = d_b1; b_2 = d_b2;
  }
  // fixed-order reduction of the per-CTA partials: warp w sums blocks w, w+32, ...; then the 32 warp sums in order
  for (int c = lane; c <= C; c += 32) {
    float acc = 0.f;
    for (int b = warp; b < nblk; b += kSaSlotThreads / 32) acc += p.partial[((long long)b * K + k) * (C + 1) + c];
    s_red[warp][c] = acc;
  }
  __syncthreads();                                   // also publishes the mbarrier initialisation to the waiting threads
  if (threadIdx.x <= C) {
    float acc = 0.f;
    for (int w = 0; w < kSaSlotThreads / 32; ++w) acc += s_red[w][threadIdx.x];
    s_red[0][threadIdx.x] = acc;
  }
  __syncthreads();
  float* sv = p.saved + (long long)it * sa_saved_stride(K, C);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_u[c] = s_red[0][c] / s_red[0][C];
    s_h[c] = p.slots[k * C + c];
    sv[2 * K * C + k * C + c] = s_u[c];          // upd
  }
  if (threadIdx.x == 0) sv[3 * K * C + k] = s_red[0][C];   // sum over tokens of attn
  __syncthreads();
  if (staged) mbar_wait(&s_bar, 0);
  warp_matvec(w_ih, b_ih, s_u, s_g, 3 * C, C, false);
  warp_matvec(w_hh, b_hh, s_h, s_g + 3 * C, 3 * C, C, false);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float r = 1.f / (1.f + expf(-(s_g[c] + s_g[3 * C + c])));
    const float z = 1.f / (1.f + expf(-(s_g[C + c] + s_g[4 * C + c])));
    const float n = tanhf(s_g[2 * C + c] + r * s_g[5 * C + c]);
    s_n[c] = (1.f - z) * n + z * s_h[c];  // new hidden state
  }
  __syncthreads();
  warp0_layernorm(s_n, bg ? p.w.to_res_bg_ln_w : p.w.to_res_ln_w, bg ? p.w.to_res_bg_ln_b : p.w.to_res_ln_b, s_ln, C);
  __syncthreads();
  warp_matvec(w_1, b_1, s_ln, s_m, Hd, C, true);
  __syncthreads();
  warp_matvec(w_2, b_2, s_m, s_o, C, Hd, false);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = s_n[c] + s_o[c];
    s_n[c] = v;
    p.slots[k * C + c] = v;
    if (it + 1 < p.iters) sv[sa_saved_stride(K, C) + k * C + c] = v;   // slot_in of the next iteration
  }
  __syncthreads();
  slot_query(p, k, s_n, s_ln, s_o, it + 1 < p.iters ? sv + sa_saved_stride(K, C) + K * C : nullptr, staged ? w_q : nullptr);   // query for the next iteration
}

size_t slot_attention_ws_floats(int ntok, int C, int K) {
  const size_t nblk = (ntok + kSaTok - 1) / kSaTok;
  return (size_t)2 * ntok * C + nblk * K * (C + 1) + (size_t)K * C + (size_t)8 * sa_saved_stride(K, C);   // k | v | partials | q | saved (<= 8 iterations)
}

int launch_slot_attention(const uorf_slot_attention_weights* w, const float* feat, long long feat_ts, long long feat_cs,
                          const float* noise_fg, const float* noise_bg, int ntok, int C, int K, int hidden, int iters, float eps,
                          float* workspace, float* slots, float* attn, cudaStream_t stream) {
  if (C > kSaMaxC || K > kSaMaxK || hidden > kSaMaxHid || K < 2 || iters > 8) return UORF_ERR_UNSUPPORTED;
  SaParams p;
  p.w = *w;
  p.feat = feat; p.feat_ts = feat_ts; p.feat_cs = feat_cs;
  p.noise_fg = noise_fg; p.noise_bg = noise_bg;
  p.ntok = ntok; p.C = C; p.K = K; p.hidden = hidden; p.iters = iters; p.eps = eps;
  const int nblk = (ntok + kSaTok - 1) / kSaTok;
  p.kbuf = workspace;
  p.vbuf = workspace + (size_t)ntok * C;
  p.partial = workspace + (size_t)2 * ntok * C;
  p.qbuf = p.partial + (size_t)nblk * K * (C + 1);
  p.saved = p.qbuf + (size_t)K * C;
  p.slots = slots;
  p.attn = attn;
  // weight staging of the per-slot update kernel: every tensor it bulk-copies must be 16-byte aligned and a multiple of 16 bytes
  const size_t stage_bytes = (size_t)(2 * 3 * C * C + 2 * hidden * C + C * C + 6 * C + hidden + C) * sizeof(float);
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int staged = (stage_bytes <= 200 * 1024 && C % 4 == 0 && hidden % 4 == 0 && al(w->gru_w_ih) && al(w->gru_w_hh) && al(w->gru_bg_w_ih) && al(w->gru_bg_w_hh) &&
                      al(w->to_res_w1) && al(w->to_res_w2) && al(w->to_res_bg_w1) && al(w->to_res_bg_w2) && al(w->to_q_w) && al(w->to_q_bg_w) &&
                      al(w->gru_b_ih) && al(w->gru_b_hh) && al(w->gru_bg_b_ih) && al(w->gru_bg_b_hh) && al(w->to_res_b1) && al(w->to_res_b2) &&
                      al(w->to_res_bg_b1) && al(w->to_res_bg_b2)) ? 1 : 0;
  // set on every launch, like the decoder kernels: function attributes are per device, a process may drive several
  if (staged && cudaFuncSetAttribute(sa_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return UORF_ERR_DEVICE;
  sa_kv_kernel<<<nblk, kSaThreads, 0, stream>>>(p);
  sa_init_kernel<<<K, kSaSlotThreads, 0, stream>>>(p);
  for (int it = 0; it < iters; ++it) {
    sa_attn_kernel<<<nblk, kSaThreads, 0, stream>>>(p);
    sa_update_kernel<<<K, kSaSlotThreads, staged ? stage_bytes : 0, stream>>>(p, nblk, it, staged);
  }
  return finish_launch(2 + 2 * iters);
}

}  // namespace uorf
