// Micro-benchmark: per-SM throughput of the instructions the ELU epilogue is built from, alone and
// interleaved 1:1 with FFMA (same pipe -> times add, different pipe -> they overlap).  B200 planning aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 512
#define UNR 8

template <int OP> __device__ __forceinline__ void op(uint32_t& r, uint32_t a, uint32_t b) {
  if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
  if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (OP == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r));
  if (OP == 4) { uint16_t h = (uint16_t)r; asm volatile("ex2.approx.f16 %0, %0;" : "+h"(h)); r = h; }
  if (OP == 5) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.b32 %0, %1, %2, p;}" : "+r"(r) : "r"(a), "r"(b));
  if (OP == 6) asm volatile("max.f32 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (OP == 7) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %0, %1;" : "+r"(r) : "r"(a));   // F2FP (float regs as b32)
  if (OP == 8) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
  if (OP == 9) asm volatile("set.lt.u32.f16x2 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (OP == 10) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (OP == 11) asm volatile("lop3.b32 %0, %0, %1, %2, 0xb8;" : "+r"(r) : "r"(a), "r"(b));
  if (OP == 12) asm volatile("add.s32 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (OP == 13) asm volatile("shl.b32 %0, %0, 23;" : "+r"(r));
  if (OP == 14) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(r) : "r"(a));
  if (OP == 15) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (OP == 16) asm volatile("{.reg .pred p; setp.ne.b32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(r) : "r"(a), "r"(b));
  if (OP == 17) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(r) : "r"(a));
  if (OP == 18) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b));
  if (OP == 19) asm volatile("cvt.rzi.s32.f32 %0, %0;" : "+r"(r));
  if (OP == 20) asm volatile("cvt.rmi.f32.f32 %0, %0;" : "+r"(r));   // floor (FRND)
}

template <int OPA, int OPB>
__global__ void __launch_bounds__(512, 1) bench(uint32_t* out, long long* cyc, uint32_t a, uint32_t b) {
  uint32_t r[UNR], q[UNR];
#pragma unroll
  for (int i = 0; i < UNR; ++i) { r[i] = threadIdx.x * 3 + i; q[i] = threadIdx.x + 7 * i; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < UNR; ++i) {
      op<OPA>(r[i], a, b);
      if (OPB >= 0) op<(OPB >= 0 ? OPB : 0)>(q[i], a, b);
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < UNR; ++i) s += r[i] + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int A, int B> double run(uint32_t* out, long long* cyc) {
  bench<A, B><<<148, 512>>>(out, cyc, 0x3f800001u, 0x3c003c00u);
  cudaDeviceSynchronize();
  bench<A, B><<<148, 512>>>(out, cyc, 0x3f800001u, 0x3c003c00u);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  return avg / (double)(ITER * UNR * 16);   // cycles per warp-instruction(-pair) per SM
}

#define ONE(A, name) do { double s = run<A, -1>(out, cyc); double m = run<A, 0>(out, cyc); double l = run<A, 11>(out, cyc); double x = run<A, 3>(out, cyc); \
  printf("%-28s solo %6.3f cyc/warp-instr/SM (%5.1f lanes/clk) | +FFMA %6.3f | +LOP3 %6.3f | +MUFU %6.3f\n", name, s, 32.0 / s, m, l, x); } while (0)

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  ONE(0, "FFMA"); ONE(1, "FADD"); ONE(2, "FMUL"); ONE(3, "MUFU.EX2 f32"); ONE(4, "MUFU.EX2.F16");
  ONE(5, "FSETP+SEL"); ONE(6, "FMNMX"); ONE(7, "F2FP f16x2"); ONE(17, "F2FP bf16x2"); ONE(8, "HFMA2"); ONE(9, "HSET2"); ONE(10, "HMNMX2");
  ONE(15, "HADD2"); ONE(11, "LOP3"); ONE(12, "IADD"); ONE(13, "SHL"); ONE(14, "PRMT"); ONE(16, "SETP.NE+SEL(reg)"); ONE(18, "IMAD"); ONE(19, "F2I"); ONE(20, "FRND.floor");
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
