// Shared device-side helpers for the uORF B200 kernels (sm_100a only).
// Thin wrappers over the PTX this library is built on: mbarrier, bulk async copies (TMA engine),
// tcgen05 (TMEM allocation, MMA, TMEM load/store) and the fences between them.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/uorf_b200.h"  // UORF_OK / UORF_ERR_* codes

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "uorf_b200 kernels must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace uorf {

// End of every launch_* function (defined in abi.cu): adds `n_kernels` to the library's launch counter (uorf_launch_count) and
// turns the CUDA launch status into UORF_OK / UORF_ERR_DEVICE, keeping the CUDA error text for uorf_last_error().
int finish_launch(int n_kernels);

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (used by the polling MMA issuer)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// named barrier among `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Up to `iters` probes of the barrier phase in a 5-instruction loop (SYNCS.TRYWAIT, BRA, IADD, ISETP, BRA).  Waiting warps
// share issue slots with the warps doing the work they wait for, so the loop is kept as small as it can be; the optional
// suspend-time hint (UORF_WAIT_HINT_NS > 0) lets the hardware park the warp for longer per probe.
#ifndef UORF_WAIT_HINT_NS
#define UORF_WAIT_HINT_NS 0
#endif
#ifndef UORF_WAIT_SLEEP_NS
#define UORF_WAIT_SLEEP_NS 0
#endif
#define UORF_STR2(x) #x
#define UORF_STR(x) UORF_STR2(x)
#if UORF_WAIT_SLEEP_NS > 0
#define UORF_WAIT_BACKOFF "nanosleep.u32 " UORF_STR(UORF_WAIT_SLEEP_NS) ";\n\t"
#else
#define UORF_WAIT_BACKOFF ""
#endif
__device__ __forceinline__ bool mbar_spin(uint32_t addr, uint32_t parity, uint32_t iters) {
  uint32_t ok;
#if UORF_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\t"
      "mov.u32 n, %3;\n"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %4;\n\t"
      "@p bra LAB_DONE;\n\t" UORF_WAIT_BACKOFF
      "sub.u32 n, n, 1;\n\t"
      "setp.ne.u32 q, n, 0;\n\t"
      "@q bra LAB_WAIT;\n"
      "LAB_DONE:\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity), "r"(iters), "r"((uint32_t)UORF_WAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\t"
      "mov.u32 n, %3;\n"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "@p bra LAB_DONE;\n\t" UORF_WAIT_BACKOFF
      "sub.u32 n, n, 1;\n\t"
      "setp.ne.u32 q, n, 0;\n\t"
      "@q bra LAB_WAIT;\n"
      "LAB_DONE:\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity), "r"(iters)
      : "memory");
#endif
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU box.  The fast path never reads the clock; after 2^14 failed probes
// the wait continues in rounds of 2^14 with a wall-clock check between rounds, and after ~2 s the CTA records the failing
// site in *err_flag (if given) and traps; the host sees a launch failure.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag = nullptr, int site = 0) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_spin(addr, parity, 1u << 14)) return;
  const long long t0 = clock64();
  while (!mbar_spin(addr, parity, 1u << 14)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, site ? site : 1);
      __threadfence_system();
      __trap();
    }
  }
}

// Lean bounded wait on a shared-memory mbarrier address: ONE probe loop (5 instructions per probe); a protocol bug traps after
// 2^22 failed probes (~0.1 s) instead of hanging the box.  mbar_wait() above (two loop levels + clock reads) costs ~250 cycles
// of set-up, branches and instruction fetch per call - too much for a wait that sits on the dependent chain of every hand-off.
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar_addr, uint32_t parity, int* err_flag, int site) {
  if (!mbar_spin(bar_addr, parity, 1u << 22)) {
    if (err_flag) atomicExch(err_flag, site);
    __threadfence_system();
    __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// keeps a value in a register: the compiler may not rematerialise it from threadIdx / constants at every use
__device__ __forceinline__ uint32_t opaque(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// ----------------------------------------------------------------------------------------------
// bulk async copy global -> shared (TMA engine, 1-D; SASS UBLKCP), completion on an mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: the mbarrier receives one arrival once every MMA issued so far by this thread
// has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05.mma  kind::f16 (bf16 x bf16 -> fp32), M = 128, cta_group::1
// ----------------------------------------------------------------------------------------------
// Instruction descriptor (32 bit): c_format F32 @4, a/b_format (0 = F16, 1 = BF16) @7/@10, K-major A and B,
// N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc_f32acc(int M, int N, bool f16) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 64 bf16 (128 B)
// with the 128-byte swizzle (16-byte chunk index XOR (row & 7)); 8-row groups are 1024 B apart.
// version = 1 (sm_100) @46, layout SWIZZLE_128B (=2) @61, LBO field = 1, SBO = 1024 B.
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr) {
  return (static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4)) | (static_cast<uint64_t>(1) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
         (static_cast<uint64_t>(2) << 61);
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-collective forms: every lane of a converged warp executes the call, one elected lane issues.  Keeping the
// control flow warp-uniform lets the compiler hold the operands in uniform registers (no per-MMA waterfall loop).
__device__ __forceinline__ void mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// NK accumulating MMAs (both operands from shared memory) under ONE elect.sync, in a single asm block: K-step j uses
// a_desc + j * ASTEP and b_desc + j * BSTEP (descriptor units of 16 bytes).  The first MMA overwrites D when accumulate == 0.
#define UORF_SS_FIRST "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, p;\n\t"
#define UORF_SS_NEXT "add.u64 a, a, %5;\n\tadd.u64 b, b, %6;\n\t@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t"
#define UORF_SS_HEAD                                                                                      \
  "{\n\t.reg .pred e, p, t;\n\t.reg .b64 a, b;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t" \
  "setp.eq.b32 t, %4, %4;\n\tmov.b64 a, %1;\n\tmov.b64 b, %2;\n\t"
#define UORF_SS_ARGS ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "l"((uint64_t)ASTEP), "l"((uint64_t)BSTEP) : "memory"
template <int NK, int ASTEP, int BSTEP>
__device__ __forceinline__ void mma_group_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  static_assert(NK == 2 || NK == 3 || NK == 4 || NK == 8, "unsupported group");
  if constexpr (NK == 2) asm volatile(UORF_SS_HEAD UORF_SS_FIRST UORF_SS_NEXT "}" UORF_SS_ARGS);
  else if constexpr (NK == 3) asm volatile(UORF_SS_HEAD UORF_SS_FIRST UORF_SS_NEXT UORF_SS_NEXT "}" UORF_SS_ARGS);
  else if constexpr (NK == 4) asm volatile(UORF_SS_HEAD UORF_SS_FIRST UORF_SS_NEXT UORF_SS_NEXT UORF_SS_NEXT "}" UORF_SS_ARGS);
  else asm volatile(UORF_SS_HEAD UORF_SS_FIRST UORF_SS_NEXT UORF_SS_NEXT UORF_SS_NEXT UORF_SS_NEXT UORF_SS_NEXT UORF_SS_NEXT UORF_SS_NEXT "}"
                    UORF_SS_ARGS);
}
#undef UORF_SS_FIRST
#undef UORF_SS_NEXT
#undef UORF_SS_HEAD
#undef UORF_SS_ARGS
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// TMEM <-> registers, shape 32x32b: lane i of the warp touches TMEM lane (base_lane + i),
// N consecutive 32-bit columns.  A warp may only touch lanes [32*(warp%4), 32*(warp%4)+32).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// 16-bit operand packing.  pack2<F16>(e0, e1): e0 -> bits [0,16), e1 -> bits [16,32)  (element 2c in the
// low half of TMEM/SMEM word c); F16 = false: bf16, true: fp16 (saturating, so no inf is ever produced).
// split2 additionally returns the packed residual (x - hi(x)) -> "lo" word of the 3-term split
//   a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo      (bf16: ~16 mantissa bits, fp16: ~22 mantissa bits).
// ----------------------------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float e0, float e1) {
  uint32_t d;
  if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  else     asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
// residual of a 16-bit pair against its fp32 source: r = e - float(hi half), ONE mixed-precision FMA per element
// (fma.rn.f32.f16 / .bf16, SASS FHFMA, reads the half straight out of the packed register: no unpack, no separate subtract)
template <bool F16>
__device__ __forceinline__ void residual2(float e0, float e1, uint32_t hi, float& r0, float& r1) {
  if (F16)
    asm("{\n\t.reg .f16 x0, x1, m;\n\tmov.b32 {x0, x1}, %4;\n\tmov.b16 m, 0xBC00;\n\t"
        "fma.rn.f32.f16 %0, x0, m, %2;\n\tfma.rn.f32.f16 %1, x1, m, %3;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(e0), "f"(e1), "r"(hi));
  else
    asm("{\n\t.reg .b16 x0, x1, m;\n\tmov.b32 {x0, x1}, %4;\n\tmov.b16 m, 0xBF80;\n\t"
        "fma.rn.f32.bf16 %0, x0, m, %2;\n\tfma.rn.f32.bf16 %1, x1, m, %3;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(e0), "f"(e1), "r"(hi));
}
template <bool F16>
__device__ __forceinline__ void split2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  hi = pack2<F16>(e0, e1);
  float r0, r1;
  residual2<F16>(e0, e1, hi, r0, r1);
  lo = pack2<F16>(r0, r1);
}
// ReLU fused into the conversion (cvt.relu): pack2_relu(e0, e1) = pack2(max(e0,0), max(e1,0)) in one instruction.
// split2_relu rounds the hi part TOWARD ZERO, so for x >= 0 the residual x - hi is >= 0, and for x < 0 hi = 0 and the
// residual x is negative: a second cvt.relu then yields exactly the split of max(x, 0) without any explicit max.
template <bool F16>
__device__ __forceinline__ uint32_t pack2_relu(float e0, float e1) {
  uint32_t d;
  if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  else     asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
template <bool F16>
__device__ __forceinline__ void split2_relu(float e0, float e1, uint32_t& hi, uint32_t& lo) {
  if (F16) asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(e1), "f"(e0));
  else     asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(e1), "f"(e0));
  float r0, r1;
  residual2<F16>(e0, e1, hi, r0, r1);
  lo = pack2_relu<F16>(r0, r1);
}
// scalar versions for the weight image
template <bool F16>
__device__ __forceinline__ unsigned short to16(float v) {
  if (F16) {
    const float c = fminf(fmaxf(v, -65504.f), 65504.f);
    return __half_as_ushort(__float2half_rn(c));
  }
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
template <bool F16>
__device__ __forceinline__ float from16(unsigned short u) {
  return F16 ? __half2float(__ushort_as_half(u)) : __bfloat162float(__ushort_as_bfloat16(u));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

}  // namespace uorf
