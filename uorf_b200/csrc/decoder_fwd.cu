// Kernel 2 — fused K-slot decoder forward on tcgen05 / TMEM.
//
// Replaces Decoder.forward (reference models/model.py:106-161) and sin_emb (:261-277):
//   per-slot object-frame transform + locality box test, sin/cos positional encoding, the background
//   MLP (7 linears) and the shared foreground MLP (9 linears, f_after_latent folded into f_color.0),
//   ReLU density / tanh colour, density-weighted composition over the K slots.
// No activation ever leaves the SM: the only HBM traffic is coords in, API-mandated outputs out.
//
// Structure: one persistent CTA per SM, epilogue warps grouped into NCH "channels" + one MMA-issuer warp per channel.
//   A channel owns one 128-point tile at a time and walks it through bg MLP / the K-1 fg slots.  Thread <-> TMEM lane <-> point.
//     3-pass modes : NCH = 4 channels x 4 warps, one thread per point (PSM variant, the default).  TMEM per channel: D 64 | H hi 32 |
//                    H lo 32 = 128 columns, so FOUR tiles are in flight in the 512 columns.  The positional-encoding operand P
//                    (2 of a slot's 26 K-steps) lives in SHARED memory instead (one 16 KB strip per channel: 128 rows x [hi 32 | lo 32]
//                    16-bit elements, K-major, 128-byte swizzle, written by the epilogue threads with st.shared + fence.proxy.async) and
//                    is the A operand of SS MMAs (52 instead of 34.6 cycles, for those two K-steps only); the per-slot head results
//                    are stashed in a small L2-resident global buffer instead of shared memory.
//                    Only 32 of the 33 encoding elements go through the tensor core; cos(16 z) is an fp32 FMA in the epilogue.
//                    (Previous form, kept as -DUORF_FWD_X3_LEGACY: 3 channels x 8 warps, P in TMEM, 160 columns per channel.)
//     1-pass modes : NCH = 4 channels x 4 warps (TMEM per channel: D 64 | H 32 | P 24 | selector 8 columns)
//   warp EPI + c    MMA issuer of channel c (elected lane issues tcgen05.mma); the first issuer warp also owns TMEM alloc and the
//                   weight bulk copies.  One issuer per channel keeps the channels out of lock-step: while one channel is in
//                   its epilogue the tensor pipe works for another.
// Per layer:  A (activations, 16-bit hi [+lo]) lives in TMEM, written by the epilogue threads with tcgen05.st;
//   B (weights) is a 128B-swizzled K-major tile in shared memory; D (fp32) in TMEM.
//   epilogue --a_full mbarrier--> issuer --tcgen05.commit / d_full mbarrier--> epilogue (ReLU + 16-bit split).
//   1-pass modes: biases never touch the CUDA cores: the weight image carries a bias tile whose COLUMN j is bias vector j, the epilogue
//   writes a one-hot "selector" K-step (8 TMEM columns) naming the next layer's bias, and the issuer appends one MMA
//   D += selector . bias_tile^T to the layer - the accumulator the epilogue reads already contains the bias.
// Precision: 1 pass of 16-bit operands, or 3 passes a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (fp32 accumulate);
//   operands bf16 (range-safe, ~16 bits after the split) or fp16 (saturating, ~22 bits after the split).
// The kernel makes two sweeps over its tiles: background MLP first (raw bg output parked in a 16 B/point
// scratch), then the foreground MLP for slots 1..K-1 followed by the cross-slot composition.  The split
// lets the 3-pass weight image of ONE MLP (136 KB) stay resident in shared memory.
#include "decoder_fwd_common.cuh"

// -DUORF_TIMELINE (tuning variant only, scripts/timeline.py): CTA 0 / channel 0 records clock64 stamps of the hand-off
// points of every stage into a global buffer read back through uorf_debug_timeline()
#ifdef UORF_TIMELINE
__device__ long long g_tl_epi[8192];
__device__ long long g_tl_iss[8192];
#define TL_EPI(id) do { if (tl_on && tl_ne < 8192) g_tl_epi[tl_ne++] = (clock64() << 8) | (id); } while (0)
#define TL_ISS(id) do { if (tl_on && tl_ni < 8192) g_tl_iss[tl_ni++] = (clock64() << 8) | (id); } while (0)
#else
#define TL_EPI(id)
#define TL_ISS(id)
#endif

namespace uorf {

template <int PASSES, bool F16, int NCH, bool PSM = false>
__global__ void __launch_bounds__(Cfg<PASSES, NCH, PSM>::THREADS, 1) decoder_fwd_kernel(const DecFwdParams p) {
  using C = Cfg<PASSES, NCH, PSM>;
  constexpr int kEpiWarps = C::EPI_WARPS;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ __align__(8) ChanBars bars[C::NCH];
  __shared__ uint32_t tmem_base_s;

  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int K = p.K;
  const int w_region = C::NPL * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  const float* s_params = reinterpret_cast<const float*>(smem + C::NPL * kPlaneBytes);
  float4* s_stash = reinterpret_cast<float4*>(smem + ((w_region + 127) & ~127));  // [NCH][K][128]   (not PSM)
  unsigned char* s_strips = smem + ((w_region + 1023) & ~1023);                   // [NCH][128 rows x 128 B] (PSM)

  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int c = 0; c < C::NCH; ++c) {
      mbar_init(&bars[c].a_full, UORF_WARP_ARRIVE ? C::WPC : kTileM * C::CS);
      mbar_init(&bars[c].d_full, 1);
    }
    fence_mbar_init();
  }
  if (warp == kEpiWarps) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // rows of the fg transform: [r][0..2] rotation, [r][3] translation (0 unless fixed_locality)
  float T[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      T[r * 4 + c] = p.fixed_locality ? __ldg(p.fg_transform + r * 4 + c) : (c < 3 ? __ldg(p.fg_transform + r * 3 + c) : 0.f);
  }

  // completion counts of the barriers (parity = count & 1); the epilogue threads and the issuer keep their own copies
  uint32_t n_d = 0;   // epilogue: d_full completions consumed by this thread
  uint32_t n_a = 0;   // issuer: a_full completions consumed

  for (int phase = 0; phase < 2; ++phase) {  // 0: background MLP, 1: foreground MLP + composition
    const int nslots = phase == 0 ? 1 : K - 1;
    // ---- load this phase's weight image ----------------------------------------------------------
    if (warp == kEpiWarps && lane == 0) {
      fence_proxy_async_smem();
      const unsigned char* src = p.image + (phase ? blob_fg_offset(PASSES) : 0);
      const uint32_t bytes = (uint32_t)blob_bytes_one(PASSES, nslots);
      mbar_arrive_expect_tx(&bar_w, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
        bulk_g2s(smem + off, src + off, n, &bar_w);
      }
    }
    mbar_wait(&bar_w, phase & 1, p.err_flag, 100 + phase);

    if (warp >= kEpiWarps) {
      // =========================== MMA issuer of channel (warp - EPI_WARPS) ===========================
      // The loop is instantiated once per channel with the channel index as a compile-time constant: every MMA operand is
      // then derived from kernel parameters, constants and one shared-memory word, so the compiler keeps descriptors and
      // TMEM addresses on the uniform datapath (a runtime channel index costs ~10 vector instructions + R2URs per MMA, and
      // the issuer shares its scheduler with six epilogue warps: the hand-off timeline showed 80 cycles per MMA issued).
      auto issuer = [&](auto CH) {
      constexpr int c = decltype(CH)::value;
      const uint32_t idesc64 = make_idesc_f32acc(128, 64, F16);
      const uint32_t idesc_head = make_idesc_f32acc(128, phase == 0 ? 16 : 32, F16);
      const uint32_t w_smem = smem_u32(smem);
      const uint32_t tbd = tmem_base + C::d_col(c);
      const uint32_t tb = tmem_base + C::a_col(c);
#ifdef UORF_TIMELINE
      const bool tl_on = blockIdx.x == 0 && c == 0 && lane == 0 && phase == 1;
      int tl_ni = 0;
#endif
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * ((long long)it * C::NCH + c);
        if (tile >= p.num_tiles || (UORF_ACTIVE_CHANNELS < C::NCH && c >= UORF_ACTIVE_CHANNELS)) break;
        // the seven stages are instantiated separately (compile-time stage index): tile addresses, K-step counts and descriptors
        // are constants on the uniform datapath and the issuer's path between wake-up and commit is short straight-line code
        // (the loop + switch form spent ~500 cycles there per stage on branches and descriptor rebuilds - hand-off timeline)
        auto stage = [&](auto ST, int s) {
          constexpr int st = decltype(ST)::value;
          mbar_wait(&bars[c].a_full, n_a & 1, p.err_flag, 200 + st);
          ++n_a;
          tc_fence_after();
          TL_ISS(0x10 + st);
          bool first = !(PASSES == 3 && UORF_X3_BIAS_PRELOAD && st < 6);   // preloaded bias: accumulate from the first K-step
          const uint32_t t_d = tbd, t_p = tb + C::COL_PHI, t_pl = tb + C::COL_PLO, t_h = tb + C::COL_HHI, t_hl = tb + C::COL_HLO;
          if (C::BIAS_MMA && st < 6) {   // bias of this layer: D = selector . bias_tile^T  (K-step = bias_id / 16), hi [+ lo] plane
            const int bid = st == 0 ? bias_id_slot(s, 0) : st == 3 ? bias_id_slot(s, 1) : bias_id_fixed(st < 3 ? st - 1 : st - 2);
            const uint64_t db = make_sdesc_sw128(w_smem + kBiasTileOff) + 2 * (bid >> 4);
            mma_ts_elect(t_d, tb + C::COL_SEL, db, idesc64, 0u);
            if (PASSES == 3) mma_ts_elect(t_d, tb + C::COL_SEL, make_sdesc_sw128(w_smem + kPlaneBytes + kBiasTileOff) + 2 * (bid >> 4), idesc64, 1u);
            first = false;
          }
          if constexpr (st == 0 && PSM) {
            mma_group_p_ss(t_d, make_sdesc_sw128(smem_u32(s_strips) + c * C::kStripBytes), make_sdesc_sw128(w_smem + 0 * kTileBytes),
                           make_sdesc_sw128(w_smem + kPlaneBytes + 0 * kTileBytes), idesc64, first ? 0u : 1u);
            first = false;
          } else if constexpr (st == 0) {
            issue_group<C::NKP, PASSES>(t_d, t_p, t_pl, w_smem + 0 * kTileBytes, idesc64, first);
          } else if constexpr (st == 3 && PSM) {
            mma_group_p_ss(t_d, make_sdesc_sw128(smem_u32(s_strips) + c * C::kStripBytes), make_sdesc_sw128(w_smem + 3 * kTileBytes),
                           make_sdesc_sw128(w_smem + kPlaneBytes + 3 * kTileBytes), idesc64, first ? 0u : 1u);
            first = false;
            issue_group<4, PASSES>(t_d, t_h, t_hl, w_smem + 4 * kTileBytes, idesc64, first);
          } else if constexpr (st == 3) {
            issue_group<C::NKP, PASSES>(t_d, t_p, t_pl, w_smem + 3 * kTileBytes, idesc64, first);
            issue_group<4, PASSES>(t_d, t_h, t_hl, w_smem + 4 * kTileBytes, idesc64, first);
          } else if constexpr (st == 6) {
            issue_group<4, PASSES>(t_d, t_h, t_hl, w_smem + 7 * kTileBytes, idesc_head, first);
          } else {
            issue_group<4, PASSES>(t_d, t_h, t_hl, w_smem + (st < 3 ? st : st + 1) * kTileBytes, idesc64, first);
          }
          tc_commit_elect(&bars[c].d_full);
          TL_ISS(0x20 + st);
          __syncwarp();
        };
        for (int s = 0; s < nslots; ++s) {
          stage(std::integral_constant<int, 0>{}, s);
          stage(std::integral_constant<int, 1>{}, s);
          stage(std::integral_constant<int, 2>{}, s);
          stage(std::integral_constant<int, 3>{}, s);
          stage(std::integral_constant<int, 4>{}, s);
          stage(std::integral_constant<int, 5>{}, s);
          stage(std::integral_constant<int, 6>{}, s);
        }
      }
      };
      switch (warp - kEpiWarps) {
        case 0: issuer(std::integral_constant<int, 0>{}); break;
        case 1: issuer(std::integral_constant<int, 1>{}); break;
        case 2: if constexpr (C::NCH > 2) issuer(std::integral_constant<int, 2>{}); break;
        default: if constexpr (C::NCH > 3) issuer(std::integral_constant<int, 3>{}); break;
      }
    } else {
      // =========================== epilogue warps ===========================
      const int ch = warp / C::WPC;
      const int wic = warp % C::WPC;          // warp within channel
      const int half = wic >> 2;              // which half of the feature columns (3-pass modes); 0 otherwise
      const int t = (wic & 3) * 32 + lane;    // point within the tile == TMEM lane
      const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t tbd = tmem_base + C::d_col(ch) + lane_bits;             // accumulator
      const uint32_t tb = tmem_base + C::a_col(ch) + lane_bits;              // operand block
      ChanBars& cb = bars[ch];
      // per-slot head results of the tile in flight: shared memory, or (PSM) this channel's slice of a global buffer - each
      // thread only ever reads back what it wrote itself, so no fence is needed, and the 10-40 KB stay L2-resident
      float4* stash = PSM ? p.stash_g + ((size_t)blockIdx.x * C::NCH + ch) * K * kTileM : s_stash + (size_t)ch * K * kTileM;
      unsigned char* strip = s_strips + ch * C::kStripBytes;
      constexpr int NCOL = 64 / C::CS;        // accumulator columns this thread owns per hidden layer
      const int col0 = half * NCOL;
      // wait for the MMAs of the current stage
#ifdef UORF_TIMELINE
      const bool tl_on = blockIdx.x == 0 && ch == 0 && wic == 0 && lane == 0 && phase == 1;
      int tl_ne = 0;
#endif
      auto wait_d = [&](int site) {
        if (C::CS == 2) {
          // 3-pass modes (8 warps per channel): one polling warp, the others sleep in a hardware barrier and spend no issue
          // slots on probing (measured 3 % faster; no difference with 4 warps per channel)
          if (wic == 0) mbar_wait(&cb.d_full, n_d & 1, p.err_flag, site);
          named_bar_sync(13 + ch, C::WPC * 32);
        } else {
          mbar_wait(&cb.d_full, n_d & 1, p.err_flag, site);
        }
        ++n_d;
        tc_fence_after();
        TL_EPI(0x30 + (site & 15));
      };
      for (int it = 0;; ++it) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * ((long long)it * C::NCH + ch);
        if (tile >= p.num_tiles || (UORF_ACTIVE_CHANNELS < C::NCH && ch >= UORF_ACTIVE_CHANNELS)) break;
        const long long pt = tile * kTileM + t;
        const bool valid = pt < p.P;
        // coordinates are fetched one slot ahead (and once per tile when all fg slots share them, fg_slot_stride == 0), so
        // the global-load latency is off the per-slot critical path
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (valid) {
          const float* src = (phase == 0 ? p.coor_bg : p.coor_fg) + pt * 3;
          nx = __ldg(src); ny = __ldg(src + 1); nz = __ldg(src + 2);
        }
        // When every fg slot sees the same coordinates (the reference's expand() view, no moved object) the encoding operand
        // written for the first slot is still in TMEM, and the locality flag is the same: later slots skip straight to the MMA.
        const bool shared_enc = phase == 1 && p.fg_slot_stride == 0 && p.fg_slot_offset == nullptr;
        float x_last = 0.f;
        bool outside = false;
        for (int s = 0; s < nslots; ++s) {
          // ---- stage 0: coordinates -> object frame -> locality flag -> positional encoding ----
          if (s == 0 || !shared_enc) {
            float x = nx, y = ny, z = nz;
            outside = false;
            if (valid && phase == 1 && p.fg_slot_stride != 0 && s + 1 < nslots) {
              const float* src = p.coor_fg + (long long)(s + 1) * p.fg_slot_stride + pt * 3;
              nx = __ldg(src); ny = __ldg(src + 1); nz = __ldg(src + 2);
            }
            if (phase == 1) {
              if (p.fg_slot_offset != nullptr) {   // moved object: same fp32 subtraction the reference applies to its coordinate copy
                x = __fsub_rn(x, __ldg(p.fg_slot_offset + 3 * s));
                y = __fsub_rn(y, __ldg(p.fg_slot_offset + 3 * s + 1));
                z = __fsub_rn(z, __ldg(p.fg_slot_offset + 3 * s + 2));
              }
              if (p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
              // unfused multiply/add in the order ((r0*x + r1*y) + r2*z) [+ t]: bit-equal to the reference's
              // CPU matmul
              const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[3]);
              const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z)), T[7]);
              const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z)), T[11]);
              x = tx; y = ty; z = tz;
              if (!p.fixed_locality) outside = fabsf(x) > p.locality_ratio || fabsf(y) > p.locality_ratio || fabsf(z) > p.locality_ratio;
            }
            float v[kPosencK];
            posenc48<PASSES == 1>(x, y, z, v);
            if constexpr (PSM) {
              posenc_store_smem<F16>(v, strip, t);
              fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
            } else if (C::CS == 1) {
              posenc_store<PASSES, F16, 0, 24>(v, tb + C::COL_PHI, tb + C::COL_PLO);
            } else if (half == 0) {
              posenc_store<PASSES, F16, 0, 8>(v, tb + C::COL_PHI, tb + C::COL_PLO);
            } else {
              posenc_store<PASSES, F16, 8, 8>(v, tb + C::COL_PHI, tb + C::COL_PLO);
            }
            if (PASSES == 3) x_last = v[kPosencMma3];   // cos(16 z): added by FMA in the epilogues of layers before.0 / after.0
          }
          if (phase == 1 && p.outside && valid && half == 0) p.outside[(long long)s * p.P + pt] = outside ? 1 : 0;
          if (PASSES == 3 && UORF_X3_BIAS_PRELOAD) {
#pragma unroll
            for (int h32 = 0; h32 < NCOL / 32; ++h32)
              preload_bias32(tbd + col0 + 32 * h32, s_params + kParamFloats + s * kSlotBiasFloats + col0 + 32 * h32);
          }
          if (C::BIAS_MMA && half == 0) write_selector<F16>(tb + C::COL_SEL, bias_id_slot(s, 0));
          tc_wait_st();
          tc_fence_before();
          chan_arrive(&cb.a_full);
          TL_EPI(0x40);

          // ---- stages 1..6: hidden layers ----
          const float* sb = s_params + kParamFloats + s * kSlotBiasFloats;
#pragma unroll 1
          for (int st = 1; st <= 6; ++st) {   // kept rolled: unrolling the six epilogue stages was 10 % slower (instruction cache)
            const float* bias = (st == 1 ? sb : st == 2 ? s_params + kPB1 : st == 3 ? s_params + kPB2
                               : st == 4 ? sb + 64 : st == 5 ? s_params + kPA1 : s_params + kPA2) + col0;
            wait_d(300 + st);
#pragma unroll
            for (int h32 = 0; h32 < NCOL / 32; ++h32) {
              const int c32 = col0 + 32 * h32;   // first accumulator column of this 32-column block
              const float* wx = PASSES == 3 && (st == 1 || st == 4) ? s_params + (st == 1 ? kPW0x : kPW3x) + c32 : nullptr;
              epilogue32<PASSES, F16>(tbd + c32, bias + 32 * h32, wx, x_last, tb + C::COL_HHI + c32 / 2, tb + C::COL_HLO + c32 / 2);
            }
            if (PASSES == 3 && UORF_X3_BIAS_PRELOAD && st < 6) {   // bias of the layer MMA stage `st` computes (read back at st + 1)
#pragma unroll
              for (int h32 = 0; h32 < NCOL / 32; ++h32)
                preload_bias32(tbd + col0 + 32 * h32, (st == 1 ? s_params + kPB1 : st == 2 ? s_params + kPB2 : st == 3 ? sb + 64
                                                       : st == 4 ? s_params + kPA1 : s_params + kPA2) + col0 + 32 * h32);
            }
            // selector for the bias of the layer the next MMA stage (index st) computes; the head (st == 6) keeps its bias here
            if (C::BIAS_MMA && half == 0 && st < 6)
              write_selector<F16>(tb + C::COL_SEL, st == 3 ? bias_id_slot(s, 1) : bias_id_fixed(st < 3 ? st - 1 : st - 2));
            TL_EPI(0x50 + st);
            tc_wait_st();
            TL_EPI(0x60 + st);
            tc_fence_before();
            chan_arrive(&cb.a_full);
            TL_EPI(0x40 + st);
          }
          // ---- stage 7: heads (computed by the first column half; every thread consumes the barrier phase) ----
          wait_d(307);
          if (half == 0) {
            if (phase == 0) {
              uint32_t r[16];
              tmem_ld16(tbd, r);
              tc_wait_ld();
              const float* hb = s_params + kPHeadB;
              float4 o;
              o.x = __uint_as_float(r[0]) + hb[0];
              o.y = __uint_as_float(r[1]) + hb[1];
              o.z = __uint_as_float(r[2]) + hb[2];
              o.w = __uint_as_float(r[3]) + hb[3];
              if (valid) p.scratch_bg[pt] = o;
            } else {
              uint32_t r[32];
              tmem_ld32(tbd, r);
              tc_wait_ld();
              const float* hb = s_params + kPHeadB;
              const float* wc2 = s_params + kPWc2;
              float c0 = s_params[kPBc2 + 0], c1 = s_params[kPBc2 + 1], c2 = s_params[kPBc2 + 2];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float h = fmaxf(__uint_as_float(r[i]) + hb[i], 0.f);
                c0 = fmaf(wc2[i], h, c0);
                c1 = fmaf(wc2[16 + i], h, c1);
                c2 = fmaf(wc2[32 + i], h, c2);
              }
              float sig = __uint_as_float(r[kHeadShapeRow]) + hb[kHeadShapeRow];
              if (p.locality && outside) sig = 0.f;  // model.py:147-148
              stash[(s + 1) * kTileM + t] = make_float4(c0, c1, c2, sig);
            }
          }
        }
        if (phase == 1) {
          // ---- cross-slot composition (model.py:151-159); the two column halves take alternate slots ----
          if (half == 0 && valid) stash[t] = p.scratch_bg[pt];
          if (C::CS == 2) named_bar_sync(1 + ch * 4 + (wic & 3), 64);  // head results of the partner warp are visible
          float S = 0.f;
#pragma unroll 4
          for (int k = 0; k < K; ++k) S += fmaxf(stash[k * kTileM + t].w, 0.f);
          const float denom = S + 1e-5f;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool want_sum = half == 0 && p.raws != nullptr;
#pragma unroll 2
          for (int k = 0; k < K; ++k) {
            const bool mine = C::CS == 1 || (k & 1) == half;
            if (!mine && !want_sum) continue;
            const float4 rw = stash[k * kTileM + t];
            const float dens = fmaxf(rw.w, 0.f);
            const float m = dens / denom;
            float4 u;
            u.x = rgb_act<PASSES == 1>(rw.x);
            u.y = rgb_act<PASSES == 1>(rw.y);
            u.z = rgb_act<PASSES == 1>(rw.z);
            u.w = dens;
            if (p.noise != nullptr && valid) u.w = dens + p.dens_noise * __ldg(p.noise + (long long)k * p.P + pt);
            const float4 mk = make_float4(u.x * m, u.y * m, u.z * m, u.w * m);
            acc.x += mk.x; acc.y += mk.y; acc.z += mk.z; acc.w += mk.w;
            if (valid && mine) {
              if (p.unmasked) p.unmasked[(long long)k * p.P + pt] = u;
              if (p.masked) p.masked[(long long)k * p.P + pt] = mk;
              if (p.masks) p.masks[(long long)k * p.P + pt] = m;
            }
          }
          if (valid && want_sum) p.raws[pt] = acc;
          if (C::CS == 2) named_bar_sync(1 + ch * 4 + (wic & 3), 64);  // stash may be overwritten by the next tile
        }
      }
    }
    // all MMAs of this phase have been consumed; shared memory may be re-used for the next image
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == kEpiWarps) tmem_dealloc(tmem_base, kTmemCols);
}

static size_t decoder_smem_bytes(int passes, int K, int nch, bool psm) {
  const int npl = passes == 3 ? 2 : 1;
  const int w_region = npl * kPlaneBytes + kParamBytes + (K - 1 > 1 ? K - 1 : 1) * kSlotBiasBytes;
  if (psm) return 1024 + ((w_region + 1023) & ~1023) + (size_t)nch * kTileM * 128;
  return 1024 + ((w_region + 127) & ~127) + (size_t)nch * K * kTileM * sizeof(float4);
}

// 3-pass modes: four channels x 4 warps, P operand in shared memory (PSM).  -DUORF_FWD_X3_LEGACY (tuning / A-B builds) selects the
// round-1 form (three channels x 8 warps, P operand in TMEM).
#ifdef UORF_FWD_X3_LEGACY
constexpr bool kX3Psm = false;
#else
constexpr bool kX3Psm = true;
#endif
constexpr int kPsmChannels = 4;
// The half-tile kernel also runs the one-pass fast modes (-DUORF_FWD_1PASS_HALFTILE=1, parity-green), but there the whole-layer
// form of this file is faster (bf16: 1.51 ms against 1.75 ms): with a third of the MMAs and a light epilogue the doubled hand-offs
// cost more than the overlap gains.
#ifndef UORF_FWD_1PASS_HALFTILE
#define UORF_FWD_1PASS_HALFTILE 0
#endif
// The shipped 3-pass kernel is decoder_fwd_x3.cu (half-tile pipelining, one issuer code path); -DUORF_FWD_X3_V1 (A-B builds) keeps
// the whole-layer hand-off form of this file.
int launch_decoder_fwd_x3(const DecFwdParams& p, int passes, bool f16, int num_sms, cudaStream_t stream);

// floats of `workspace`: the background sweep's raw outputs [P][4] + (PSM) the per-channel head stash [SMs][4][K][128][4]
size_t decoder_fwd_workspace_floats(long long P, int K, int num_sms) {
  return (size_t)P * 4 + (size_t)num_sms * kPsmChannels * K * kTileM * 4;
}

int launch_decoder_fwd(const uorf_decoder_args* a, int num_sms, int* err_flag, cudaStream_t stream) {
  DecFwdParams p;
  p.coor_bg = a->coor_bg;
  p.coor_fg = a->coor_fg;
  p.fg_slot_stride = a->fg_slot_stride;
  p.noise = a->dens_noise != 0.f ? a->noise : nullptr;
  p.image = static_cast<const unsigned char*>(a->image);
  p.scratch_bg = reinterpret_cast<float4*>(a->workspace);
  p.stash_g = reinterpret_cast<float4*>(a->workspace) + a->P;
  p.raws = reinterpret_cast<float4*>(a->raws);
  p.masked = reinterpret_cast<float4*>(a->masked_raws);
  p.unmasked = reinterpret_cast<float4*>(a->unmasked_raws);
  p.masks = a->masks;
  p.outside = a->outside;
  p.fg_transform = a->fg_transform;
  p.fg_slot_offset = a->fg_slot_offset;
  p.locality_ratio = a->locality_ratio;
  p.dens_noise = a->dens_noise;
  p.P = a->P;
  p.K = a->K;
  p.fixed_locality = a->fixed_locality;
  p.locality = a->locality;
  p.num_tiles = (int)((a->P + kTileM - 1) / kTileM);
  p.err_flag = err_flag;
  p.zero = 0;
  p.discard_bg = (reinterpret_cast<uintptr_t>(a->workspace) & 127u) == 0;
  const int passes = a->precision >= 3 ? 3 : 1;
  const bool f16 = (a->precision % 2) == 0;
  const bool psm = passes == 3 && kX3Psm;
  // legacy 3-pass form: three tiles per SM when the per-slot stash of three channels fits beside the weight image (K <= 13)
  int nch = passes == 3 ? (psm ? kPsmChannels : 3) : 4;
  if (passes == 3 && !psm && decoder_smem_bytes(passes, a->K, 3, false) > 227 * 1024) nch = 2;
  const size_t smem = decoder_smem_bytes(passes, a->K, nch, psm);
  if (smem > 227 * 1024) return UORF_ERR_UNSUPPORTED;
  int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (grid < 1) return UORF_OK;
  auto launch = [&](auto kern, int threads) -> int {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return UORF_ERR_DEVICE;
    kern<<<grid, threads, smem, stream>>>(p);
    return UORF_OK;
  };
#ifndef UORF_FWD_X3_V1
  if (psm) return launch_decoder_fwd_x3(p, 3, f16, num_sms, stream);
#endif
#if UORF_FWD_1PASS_HALFTILE
  if (passes == 1) return launch_decoder_fwd_x3(p, 1, f16, num_sms, stream);   // one-pass fast modes on the half-tile kernel
#endif
  int rc;
  if (psm) rc = f16 ? launch(decoder_fwd_kernel<3, true, kPsmChannels, true>, Cfg<3, kPsmChannels, true>::THREADS)
                    : launch(decoder_fwd_kernel<3, false, kPsmChannels, true>, Cfg<3, kPsmChannels, true>::THREADS);
  else if (passes == 3 && nch == 3) rc = f16 ? launch(decoder_fwd_kernel<3, true, 3>, Cfg<3, 3>::THREADS) : launch(decoder_fwd_kernel<3, false, 3>, Cfg<3, 3>::THREADS);
  else if (passes == 3) rc = f16 ? launch(decoder_fwd_kernel<3, true, 2>, Cfg<3, 2>::THREADS) : launch(decoder_fwd_kernel<3, false, 2>, Cfg<3, 2>::THREADS);
  else rc = f16 ? launch(decoder_fwd_kernel<1, true, 4>, Cfg<1, 4>::THREADS) : launch(decoder_fwd_kernel<1, false, 4>, Cfg<1, 4>::THREADS);
  if (rc != UORF_OK) return rc;
  return finish_launch(1);
}

}  // namespace uorf

#ifdef UORF_TIMELINE
extern "C" int uorf_debug_timeline(long long* epi, long long* iss, int n) {
  if (n > 8192) n = 8192;
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(epi, g_tl_epi, n * sizeof(long long)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(iss, g_tl_iss, n * sizeof(long long)) != cudaSuccess) return -1;
  return 0;
}
#endif
