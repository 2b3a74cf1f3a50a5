// Micro-benchmark 3: SS-form tcgen05.mma (both operands in shared memory, kind::f16, M=128, K=16 per instruction) as the
// wide GEMM kernel issues it: 4 K steps per 64-wide chunk, chunks back to back over resident operands (no loads).
// Compares the no-swizzle "interleaved" core-matrix layout (LBO 128 B, SBO 1024 B) with the 128-byte-swizzle K-major
// layout (row pitch 128 B, SBO 1024 B, K step +32 B).  Timing only: operand values are arbitrary.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../go2_onnx_controller_b200/csrc/ptx_sm100.cuh"
using namespace go2p;

// 128B swizzle: layout type 2 in bits [61,64), version 1 in [46,48)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) bench(long long* out, int N, int chunks, int stages, int swz, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = 16384 + N * 128;
  for (int i = threadIdx.x; i < stages * stage_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (warp == 0) {
    if (lane == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc<512>(&tmem_ptr);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc(ptx::FMT_F16, 128, (uint32_t)N);
    uint32_t par = 0;
    long long best_issue = 1ll << 40, best_total = 1ll << 40;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (ptx::elect_one_sync()) {
        for (int c = 0; c < chunks; ++c) {
          const uint32_t sa = ptx::smem_u32(smem) + (uint32_t)((c % stages) * stage_bytes);
          const uint64_t ad = swz ? desc_sw128(sa, 1024u) : ptx::make_smem_desc_nosw(sa, 128u, 1024u);
          const uint64_t bd = swz ? desc_sw128(sa + 16384u, 1024u) : ptx::make_smem_desc_nosw(sa + 16384u, 128u, 1024u);
          const uint64_t step = swz ? 2u : 16u;     // +32 B (swizzled rows) or +256 B (two core matrices) per K step
#pragma unroll
          for (int j = 0; j < 4; ++j) ptx::mma_f16_ss(tb, ad + step * j, bd + step * j, idesc, (c | j) ? 1u : 0u);
        }
        ptx::mma_commit(&bar);
      }
      __syncwarp();
      const long long t1 = clock64();
      ptx::mbar_wait(&bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      if (t1 - t0 < best_issue) best_issue = t1 - t0;
      if (t2 - t0 < best_total) best_total = t2 - t0;
    }
    if (lane == 0) { out[0] = best_issue; out[1] = best_total; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tb); }
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int swz : {0, 1})
    for (int N : {64, 128, 256}) {
      const int chunks = 32, stages = 4;
      bench<<<1, 128, 200 * 1024>>>(out, N, chunks, stages, swz, 10);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      const int n = chunks * 4;
      printf("SS %s N=%3d, %d MMAs: issue %6lld cyc (%5.1f/mma)  issue+complete %6lld cyc (%5.1f/mma; ideal %d) %s\n",
             swz ? "swizzle-128B" : "no-swizzle  ", N, n, h[0], (double)h[0] / n, h[1], (double)h[1] / n, N / 2,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
