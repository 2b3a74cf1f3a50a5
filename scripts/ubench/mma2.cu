// Micro-benchmark 2: tcgen05.mma (TS form, kind::f16, M=128) issued the way the kernel does it (whole warp converged,
// elect.sync lane, precomputed descriptors): cycles per MMA for N = 16 / 64 / 128, chains of 9 K-steps, back-to-back
// chains into alternating accumulators (like two slots).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../go2_onnx_controller_b200/csrc/ptx_sm100.cuh"
using namespace go2p;

__global__ void __launch_bounds__(128, 1) bench(long long* out, int N, int ksteps, int chains, int reps) {
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
  if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc(ptx::FMT_F16, 128, (uint32_t)N);
    const uint64_t bdesc0 = ptx::make_smem_desc_nosw(ptx::smem_u32(smem), 128u, 2304u);
    uint32_t par = 0;
    long long best_issue = 1ll << 40, best_total = 1ll << 40;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (ptx::elect_one_sync()) {
        for (int c = 0; c < chains; ++c) {
          const uint32_t d = tb + (uint32_t)(c & 1) * 128u;
          const uint32_t a = tb + 256u + (uint32_t)(c & 1) * 128u;
          ptx::mma_f16_ts(d, a, bdesc0, idesc, 0u);
          for (int j = 1; j < ksteps; ++j) ptx::mma_f16_ts(d, a + (uint32_t)j * 8u, bdesc0 + (uint64_t)(j * 16), idesc, 1u);
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
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {16, 32, 64, 128})
    for (int chains : {1, 2, 8}) {
      const int ksteps = 9;
      bench<<<1, 128, 64 * 1024>>>(out, N, ksteps, chains, 20);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      const int n = chains * ksteps;
      printf("TS N=%3d chains=%d x %d K-steps: issue %6lld cyc (%5.1f/mma)  issue+complete %6lld cyc (%5.1f/mma) %s\n", N, chains, ksteps,
             h[0], (double)h[0] / n, h[1], (double)h[1] / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
