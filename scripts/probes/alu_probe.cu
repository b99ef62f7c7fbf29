// Issue-rate probe for the epilogue's instruction mix (one SM sub-partition view): cycles per warp-instruction with W warps per CTA
// (1 CTA on one SM).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_probe alu_probe.cu ; ./alu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define REP 256
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[16];
  uint32_t u[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; u[i] = threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) a[i] = a[i] + 1.0009765625f;
      if (OP == 8) a[i] = __uint_as_float(__float_as_uint(a[i]) | 0x3c003c00u);   // LOP3 alone (baseline for the F2FP rows)                                                    // FADD
      if (OP == 1) { asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[i])); a[i] = __uint_as_float(u[i] | 0x3c003c00u); }   // F2FP pack (+1 LOP3)
      if (OP == 2) { asm volatile("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[i])); a[i] = __uint_as_float(u[i] | 0x3c003c00u); }
      if (OP == 3) asm volatile("{ .reg .f16 x0, x1, m; mov.b32 {x0, x1}, %1; mov.b16 m, 0xBC00; fma.rn.f32.f16 %0, x0, m, %0; }" : "+f"(a[i]) : "r"(u[i]));  // FHFMA
      if (OP == 4) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i + 1) & 15]));                 // PRMT
      if (OP == 5) { asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[i])); a[i] = __uint_as_float(u[i] | 0x3c003c00u); }
      if (OP == 6) a[i] = fmaf(a[i], 1.0001f, 0.5f);                                               // FFMA
      if (OP == 7) u[i] = (u[i] & 0x7FFFu) == 0 ? u[(i + 1) & 15] : u[i];                            // LOP3 + SEL
    }
  }
  long long t1 = clock64();
  float s = 0; uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s += a[i]; x ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
  printf("%-28s", name);
  for (int warps : {1, 4, 8, 16, 24}) {
    k<OP><<<1, warps * 32>>>(out, cyc, 1.f);
    k<OP><<<1, warps * 32>>>(out, cyc, 1.f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // per sub-partition: warps/4 warps each issuing REP*8 instructions
    double per_smsp_instr = (warps < 4 ? 1.0 : warps / 4.0) * REP * 16;
    printf("  W=%2d: %.2f cyc/inst/SMSP", warps, c / per_smsp_instr);
  }
  printf("\n");
}
int main() {
  run<0>("FADD"); run<8>("LOP3"); run<6>("FFMA"); run<1>("F2FP f16x2 rn.relu"); run<2>("F2FP f16x2 rz.relu"); run<5>("F2FP bf16x2 rn.relu");
  run<3>("FHFMA (f32 += f16*f16)"); run<4>("PRMT"); run<7>("LOP3+SEL pair");
  return 0;
}
