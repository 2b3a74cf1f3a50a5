// Micro-benchmark: tcgen05.mma issue/completion cost per instruction for the forms the policy kernel uses
// (kind::f16, M=128, A from TMEM (TS) or smem (SS), B from smem in the no-swizzle K-major layout), as a
// function of N and of how many independent accumulators the K-steps rotate over.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../go2_onnx_controller_b200/csrc/ptx_sm100.cuh"
using namespace go2p;

// mode: 0 = TS form, 1 = SS form.  nacc: number of independent D tiles rotated over.  nmma: MMAs per commit.
__global__ void __launch_bounds__(128, 1) bench(long long* out, int mode, int N, int nacc, int nmma, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
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
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc(ptx::FMT_F16, 128, (uint32_t)N);
    const uint32_t sb = ptx::smem_u32(smem);
    uint32_t par = 0;
    long long best_issue = 1 << 30, best_total = 1 << 30;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int j = 0; j < nmma; ++j) {
        const uint32_t d = tb + (uint32_t)(j % nacc) * 128u * 0 + (uint32_t)((j % nacc) * N);   // independent column ranges
        const uint64_t bdesc = ptx::make_smem_desc_nosw(sb + (uint32_t)(j & 7) * 256u, 128u, 2304u);
        if (mode == 0) ptx::mma_f16_ts(d, tb + 448u + (uint32_t)(j & 7) * 8u, bdesc, idesc, j >= nacc ? 1u : 0u);
        else ptx::mma_f16_ss(d, ptx::make_smem_desc_nosw(sb + 24576u + (uint32_t)(j & 7) * 256u, 128u, 2304u), bdesc, idesc, j >= nacc ? 1u : 0u);
      }
      ptx::mma_commit(&bar);
      const long long t1 = clock64();
      ptx::mbar_wait(&bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      if (t1 - t0 < best_issue) best_issue = t1 - t0;
      if (t2 - t0 < best_total) best_total = t2 - t0;
    }
    out[0] = best_issue; out[1] = best_total;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tb); }
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 64, 128, 256})
      for (int nacc : {1, 2, 4}) {
        if (N * nacc > 448) continue;
        for (int nmma : {8, 32}) {
          bench<<<1, 128, 64 * 1024>>>(out, mode, N, nacc, nmma, 20);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
          printf("%s N=%3d accumulators=%d mmas=%2d: issue %6lld cyc (%6.1f/mma)  issue+complete %6lld cyc (%6.1f/mma) %s\n", mode ? "SS" : "TS", N, nacc, nmma,
                 h[0], (double)h[0] / nmma, h[1], (double)h[1] / nmma, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
  return 0;
}
