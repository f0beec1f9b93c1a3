// Pipe-rate microbenchmark for the instruction mix of the ratio stage (sm_100a): cycles per warp-instruction per SMSP
// for independent instruction streams, at 1 / 2 / 4 warps per SM sub-partition.   nvcc -arch=sm_100a -O3 pipes.cu -o pipes
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#define REP 256
#define NCH 8     // independent chains per thread
template <int OP>
__global__ void k(float* out, long long* cyc, float a0, float b0, float c0) {
  float x[NCH], y[NCH];
  uint64_t X[NCH];
  uint32_t u[NCH];
  const float a = a0, b = b0, c = c0;
  for (int i = 0; i < NCH; ++i) { x[i] = a0 + i + threadIdx.x; y[i] = b0 + i; u[i] = threadIdx.x * 77 + i; asm("mov.b64 %0, {%1, %2};" : "=l"(X[i]) : "f"(x[i]), "f"(y[i])); }
  uint64_t A2, B2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(A2) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(B2) : "f"(b));
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));                 // FFMA reg,reg,reg
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3A000000;" : "+f"(x[i]));                  // FFMA imm
      if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[i]) : "l"(A2), "l"(B2));             // FFMA2
      if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(A2));                          // FMUL2
      if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(X[i]) : "l"(A2));                          // FADD2
      if (OP == 5) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(0x3c003c01u), "r"(0x00010001u));   // HFMA2 (imm-ish)
      if (OP == 6) asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(x[i]) : "r"(u[i] + r));   // HADD2.F32 (+IADD)
      if (OP == 7) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));                                  // MUFU.RCP
      if (OP == 8) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(y[i]));  // F2FP
      if (OP == 9) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));                             // FMUL
      if (OP == 10) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(u[(i + 1) % NCH]));       // PRMT
      if (OP == 11) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));                                         // SHF
      if (OP == 12) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % NCH]), "r"(u[(i + 2) % NCH]));   // HFMA2 3-reg
      if (OP == 13) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % NCH]), "r"(u[(i + 2) % NCH]));    // IMAD
      if (OP == 14) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y[i]), "f"(b));             // FFMA 3 distinct regs
      if (OP == 15) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X[i]) : "l"(X[(i + 1) % NCH]), "l"(B2));    // FFMA2 distinct
      if (OP == 16) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));                                 // MUFU.EX2
      if (OP == 17) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % NCH]), "r"(u[(i + 2) % NCH]));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < NCH; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(X[i])); s += x[i] + y[i] + lo + hi + __uint_as_float(u[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + c;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8 * 1024);
  printf("%-28s", name);
  for (int warps_per_smsp : {1, 2, 4}) {
    int threads = 128 * warps_per_smsp;
    k<OP><<<148, threads>>>(out, cyc, 1.0001f, 1e-4f, 0.f);
    cudaDeviceSynchronize();
    k<OP><<<148, threads>>>(out, cyc, 1.0001f, 1e-4f, 0.f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    // cycles per warp-instruction per SMSP = total cycles / (REP * NCH * warps_per_smsp)
    printf("  w/smsp=%d: %.2f cyc/inst", warps_per_smsp, c / (double)(REP * NCH * warps_per_smsp));
  }
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA r,r,r (2 shared srcs)"); run<14>("FFMA 3 distinct-ish"); run<1>("FFMA imm"); run<9>("FMUL");
  run<2>("FFMA2 (shared srcs)"); run<15>("FFMA2 distinct"); run<3>("FMUL2"); run<4>("FADD2");
  run<5>("HFMA2 imm"); run<12>("HFMA2 3-reg"); run<6>("HADD2.F32(+iadd)"); run<7>("MUFU.RCP"); run<16>("MUFU.EX2"); run<8>("F2FP pack");
  run<10>("PRMT"); run<11>("SHL"); run<13>("IMAD"); run<17>("LOP3");
  return 0;
}
