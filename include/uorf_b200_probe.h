/* uorf_b200_probe.h - C ABI of libuorf_b200_probe.so: TEST / DIAGNOSTIC hooks, built from csrc/probe.cu alone.
 * Not part of the product library (libuorf_b200.so exports none of these); tests/ and scripts/ load it to pin the tcgen05
 * operand layouts the decoder kernels rely on and to time their hand-off building blocks.  Same conventions as uorf_b200.h. */
#ifndef UORF_B200_PROBE_H_
#define UORF_B200_PROBE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * One 128 x N x (16*ksteps) bf16 GEMM on the tcgen05 path the decoder uses
 * (A rows written to TMEM by threads, B a swizzled shared-memory tile).  a [128][64], b [N][64] fp32
 * (rounded to 16 bits inside), d [128][N].  Used by tests/ to pin the tensor-core data layouts.
 * ------------------------------------------------------------------------------------------- */
int uorf_probe_gemm(const float* a, const float* b, int32_t N, int32_t ksteps, int32_t precision, float* d,
                    void* stream);
/* The weight-gradient data path: d1 = a^T b1, d2 = a^T b2 with a, b [128][64] (points x features, rounded to
 * bf16), d [64][64]; both operands are read through MN-major shared-memory descriptors, M = 64, two accumulators interleaved. */
int uorf_probe_gemm_tn(const float* a, const float* b1, const float* b2, float* d1, float* d2, void* stream);
/* Diagnostic hook: SM-cycle counts of the decoder's handshake building blocks (out: 256 int64, device). */
int uorf_probe_timing(int64_t* out, int32_t reps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UORF_B200_PROBE_H_ */
