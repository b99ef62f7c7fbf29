// Kernel 4 — background-aware slot attention (forward), batch 1.
// Replaces SlotAttention.forward, reference models/model.py:205-258:
//   slots_fg = mu + exp(logsigma) * noise_fg ; slot_bg likewise                      (:214-219)
//   k, v = to_k(LN(feat)), to_v(LN(feat))                                             (:221-223)
//   iters x { q = to_q(LN(slot)) (separate fg / bg) ; dots = q.k * C^-0.5 ;           (:229-233)
//             attn = softmax over SLOTS + eps ; per-slot renormalise over tokens ;     (:235-238)
//             updates = attn_w . v ; slot = GRUCell(updates, slot) ; slot += MLP(LN(slot)) }  (:240-255)
// The work is ~0.1 GFLOP and latency-bound, so it is fp32 FFMA, organised to need only two kinds of
// launch per iteration: a token-parallel pass over all SMs (q, dots, softmax, partial weighted sums) and a
// single-CTA slot update (reduction of the partials, GRU, residual MLP).  Reductions run in a fixed order,
// so results are deterministic.
#include "common.cuh"
#include "../../include/uorf_b200.h"

namespace uorf {

constexpr int kSaTok = 32;      // tokens per CTA in the token-parallel passes
constexpr int kSaThreads = 256;
constexpr int kSaMaxC = 64;
constexpr int kSaMaxK = 16;
constexpr int kSaMaxHid = 128;

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

// ---- slot initialisation --------------------------------------------------------------------------
__global__ void sa_init_kernel(const SaParams p) {
  const int C = p.C;
  for (int e = threadIdx.x; e < p.K * C; e += blockDim.x) {
    const int k = e / C, c = e % C;
    p.slots[e] = k == 0 ? p.w.slots_mu_bg[c] + expf(p.w.slots_logsigma_bg[c]) * p.noise_bg[c]
                        : p.w.slots_mu[c] + expf(p.w.slots_logsigma[c]) * p.noise_fg[(k - 1) * C + c];
  }
}

// ---- pass B (per iteration): q, dots, softmax over slots, partial sums ----------------------------
__global__ void __launch_bounds__(kSaThreads) sa_attn_kernel(const SaParams p) {
  __shared__ float s_q[kSaMaxK][kSaMaxC + 1];
  __shared__ float s_ln[kSaMaxK][kSaMaxC + 1];
  __shared__ float s_k[kSaTok][kSaMaxC + 1];
  __shared__ float s_v[kSaTok][kSaMaxC + 1];
  __shared__ float s_a[kSaMaxK][kSaTok];
  const int C = p.C, K = p.K;
  const int t0 = blockIdx.x * kSaTok;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < kSaTok * C; e += kSaThreads) {
    const int t = e / C, c = e % C;
    const bool in = t0 + t < p.ntok;
    s_k[t][c] = in ? p.kbuf[(long long)(t0 + t) * C + c] : 0.f;
    s_v[t][c] = in ? p.vbuf[(long long)(t0 + t) * C + c] : 0.f;
  }
  // LN of each slot (warp per slot)
  for (int k = warp; k < K; k += kSaThreads / 32) {
    const float* lw = k == 0 ? p.w.to_q_bg_ln_w : p.w.to_q_ln_w;
    const float* lb = k == 0 ? p.w.to_q_bg_ln_b : p.w.to_q_ln_b;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += p.slots[k * C + c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = p.slots[k * C + c] - mean; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + 1e-5f);
    for (int c = lane; c < C; c += 32) s_ln[k][c] = (p.slots[k * C + c] - mean) * rstd * lw[c] + lb[c];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * C; e += kSaThreads) {
    const int k = e / C, o = e % C;
    const float* wq = (k == 0 ? p.w.to_q_bg_w : p.w.to_q_w) + o * C;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(s_ln[k][c], __ldg(wq + c), acc);
    s_q[k][o] = acc;
  }
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

// ---- pass C (per iteration): reduce partials, GRUCell, residual MLP; one CTA ----------------------
__global__ void __launch_bounds__(kSaThreads) sa_update_kernel(const SaParams p, int nblk) {
  __shared__ float s_u[kSaMaxK][kSaMaxC + 1];   // updates
  __shared__ float s_h[kSaMaxK][kSaMaxC + 1];   // previous slots
  __shared__ float s_g[kSaMaxK][6 * kSaMaxC];   // gi (3C) | gh (3C)
  __shared__ float s_n[kSaMaxK][kSaMaxC + 1];   // new slot / LN
  __shared__ float s_m[kSaMaxK][kSaMaxHid];
  const int C = p.C, K = p.K, Hd = p.hidden;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < K * C; e += kSaThreads) {
    const int k = e / C, c = e % C;
    float num = 0.f, den = 0.f;
    for (int b = 0; b < nblk; ++b) {
      const float* part = p.partial + ((long long)b * K + k) * (C + 1);
      num += part[c];
      den += part[C];
    }
    s_u[k][c] = num / den;
    s_h[k][c] = p.slots[e];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * 6 * C; e += kSaThreads) {
    const int k = e / (6 * C), j = e % (6 * C);
    const bool hh = j >= 3 * C;
    const int o = hh ? j - 3 * C : j;
    const float* W = k == 0 ? (hh ? p.w.gru_bg_w_hh : p.w.gru_bg_w_ih) : (hh ? p.w.gru_w_hh : p.w.gru_w_ih);
    const float* B = k == 0 ? (hh ? p.w.gru_bg_b_hh : p.w.gru_bg_b_ih) : (hh ? p.w.gru_b_hh : p.w.gru_b_ih);
    const float* x = hh ? s_h[k] : s_u[k];
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(x[c], __ldg(W + o * C + c), acc);
    s_g[k][j] = acc + B[o];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * C; e += kSaThreads) {
    const int k = e / C, c = e % C;
    const float r = 1.f / (1.f + expf(-(s_g[k][c] + s_g[k][3 * C + c])));
    const float z = 1.f / (1.f + expf(-(s_g[k][C + c] + s_g[k][4 * C + c])));
    const float n = tanhf(s_g[k][2 * C + c] + r * s_g[k][5 * C + c]);
    s_u[k][c] = (1.f - z) * n + z * s_h[k][c];  // new hidden state
  }
  __syncthreads();
  for (int k = warp; k < K; k += kSaThreads / 32) {
    const float* lw = k == 0 ? p.w.to_res_bg_ln_w : p.w.to_res_ln_w;
    const float* lb = k == 0 ? p.w.to_res_bg_ln_b : p.w.to_res_ln_b;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += s_u[k][c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = s_u[k][c] - mean; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + 1e-5f);
    for (int c = lane; c < C; c += 32) s_n[k][c] = (s_u[k][c] - mean) * rstd * lw[c] + lb[c];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * Hd; e += kSaThreads) {
    const int k = e / Hd, j = e % Hd;
    const float* W = (k == 0 ? p.w.to_res_bg_w1 : p.w.to_res_w1) + j * C;
    float acc = (k == 0 ? p.w.to_res_bg_b1 : p.w.to_res_b1)[j];
    for (int c = 0; c < C; ++c) acc = fmaf(s_n[k][c], __ldg(W + c), acc);
    s_m[k][j] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * C; e += kSaThreads) {
    const int k = e / C, c = e % C;
    const float* W = (k == 0 ? p.w.to_res_bg_w2 : p.w.to_res_w2) + c * Hd;
    float acc = (k == 0 ? p.w.to_res_bg_b2 : p.w.to_res_b2)[c];
    for (int j = 0; j < Hd; ++j) acc = fmaf(s_m[k][j], __ldg(W + j), acc);
    p.slots[e] = s_u[k][c] + acc;
  }
}

size_t slot_attention_ws_floats(int ntok, int C, int K) {
  const size_t nblk = (ntok + kSaTok - 1) / kSaTok;
  return (size_t)2 * ntok * C + nblk * K * (C + 1);
}

int launch_slot_attention(const uorf_slot_attention_weights* w, const float* feat, long long feat_ts, long long feat_cs,
                          const float* noise_fg, const float* noise_bg, int ntok, int C, int K, int hidden, int iters, float eps,
                          float* workspace, float* slots, float* attn, cudaStream_t stream) {
  if (C > kSaMaxC || K > kSaMaxK || hidden > kSaMaxHid || K < 2) return UORF_ERR_UNSUPPORTED;
  SaParams p;
  p.w = *w;
  p.feat = feat; p.feat_ts = feat_ts; p.feat_cs = feat_cs;
  p.noise_fg = noise_fg; p.noise_bg = noise_bg;
  p.ntok = ntok; p.C = C; p.K = K; p.hidden = hidden; p.iters = iters; p.eps = eps;
  const int nblk = (ntok + kSaTok - 1) / kSaTok;
  p.kbuf = workspace;
  p.vbuf = workspace + (size_t)ntok * C;
  p.partial = workspace + (size_t)2 * ntok * C;
  p.slots = slots;
  p.attn = attn;
  sa_kv_kernel<<<nblk, kSaThreads, 0, stream>>>(p);
  sa_init_kernel<<<1, 256, 0, stream>>>(p);
  for (int it = 0; it < iters; ++it) {
    sa_attn_kernel<<<nblk, kSaThreads, 0, stream>>>(p);
    sa_update_kernel<<<1, kSaThreads, 0, stream>>>(p, nblk);
  }
  return cudaGetLastError() == cudaSuccess ? UORF_OK : UORF_ERR_DEVICE;
}

}  // namespace uorf
