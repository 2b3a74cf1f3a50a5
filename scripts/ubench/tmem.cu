// Micro-benchmark: TMEM <-> register bandwidth (tcgen05.ld / tcgen05.st, 32x32b shapes) per SM as a
// function of the number of warps issuing.  B200 planning aid for the epilogue design.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define R4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
#define W4(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3])

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12), R4(v, 16), R4(v, 20), R4(v, 24), R4(v, 28) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_x16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_16x256_x8(uint32_t taddr, uint32_t (&v)[32]) {   // 16 lanes x 256 bit, x8 -> 32 regs
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12), R4(v, 16), R4(v, 20), R4(v, 24), R4(v, 28) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st_x16(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), W4(v, 0), W4(v, 4), W4(v, 8), W4(v, 12) : "memory");
}
__device__ __forceinline__ void st_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), W4(v, 0), W4(v, 4), W4(v, 8), W4(v, 12), W4(v, 16), W4(v, 20), W4(v, 24), W4(v, 28) : "memory");
}

// MODE 0: ld x32 | 1: ld x16 | 2: st x16 | 3: st x32 | 4: ld 16x256b.x8 | 5: ld x32 + wait each | 6: ld x32 ; st x16 interleaved
template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* cyc, int nwarps) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
      const uint32_t col = (uint32_t)(((warp >> 2) * 64 + (it & 1) * 32) & 511);
      if (MODE == 0) { ld_x32(base + col, v); }
      if (MODE == 1) { ld_x16(base + col, v); }
      if (MODE == 2) { st_x16(base + col, v); }
      if (MODE == 3) { st_x32(base + col, v); }
      if (MODE == 4) { ld_16x256_x8(base + col, v); }
      if (MODE == 5) { ld_x32(base + col, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += v[0] + v[31]; }
      if (MODE == 6) { ld_x32(base + col, v); st_x16(base + ((col + 256) & 511), v); }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  __shared__ long long tmax;
  if (threadIdx.x == 0) tmax = 0;
  __syncthreads();
  if (warp < nwarps && (threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&tmax, (unsigned long long)(t1 - t0));
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = tmax;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr) : "memory");
}

template <int MODE> void run(const char* name, int bytes_per_instr, uint32_t* out, long long* cyc) {
  printf("%-34s", name);
  for (int nw : {1, 4, 8, 16, 32}) {
    bench<MODE><<<8, 1024>>>(out, cyc, nw);
    cudaDeviceSynchronize();
    bench<MODE><<<8, 1024>>>(out, cyc, nw);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf(" ERR %s", cudaGetErrorString(cudaGetLastError())); break; }
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 8; ++i) c += h[i]; c /= 8;
    printf(" | %2dw: %7.1f B/clk/SM (%5.1f cyc/instr)", nw, (double)bytes_per_instr * ITER * nw / c, c / ITER);
  }
  printf("\n");
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024 * 4); cudaMalloc(&cyc, 64);
  run<0>("tcgen05.ld 32x32b.x32 (4 KB)", 4096, out, cyc);
  run<1>("tcgen05.ld 32x32b.x16 (2 KB)", 2048, out, cyc);
  run<4>("tcgen05.ld 16x256b.x8 (4 KB)", 4096, out, cyc);
  run<5>("tcgen05.ld x32 + wait::ld each", 4096, out, cyc);
  run<2>("tcgen05.st 32x32b.x16 (2 KB)", 2048, out, cyc);
  run<3>("tcgen05.st 32x32b.x32 (4 KB)", 4096, out, cyc);
  run<6>("ld x32 + st x16 (6 KB)", 6144, out, cyc);
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
