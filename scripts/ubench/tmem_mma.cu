// Micro-benchmark: tcgen05.ld / tcgen05.st latency seen by 16 epilogue warps while another warp keeps the tensor core busy
// with tcgen05.mma (TS or SS form) on OTHER TMEM columns -- does the MMA's TMEM traffic slow the epilogue's, and vice versa?
//   worker modes: 0 = ld x32 + wait::ld | 1 = 2 x ld x16 + wait::ld | 2 = st x16 + wait::st | 3 = ld x32 ; st x16 ; wait both
//   mma modes:    0 = none | 1 = TS chains (A in TMEM) | 2 = SS chains (A in shared memory) | 3 = 1 SS step then 8 TS | 4 = 8 TS then 1 SS
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../go2_onnx_controller_b200/csrc/ptx_sm100.cuh"
using namespace go2p;

constexpr int kIters = 200;

__global__ void __launch_bounds__(17 * 32, 1) bench(long long* out, int wmode, int mmode, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) stop = 0;
  if (warp == 16) {
    if (lane == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc<512>(&tmem_ptr);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 16) {
    // MMA warp: chains of 9 K steps (M128 x N128 x K16) into columns 256..383, A from columns 448..511 (TS) or smem (SS)
    const uint32_t idesc = ptx::make_idesc(ptx::FMT_F16, 128, 128);
    const uint64_t bdesc = ptx::make_smem_desc_nosw(ptx::smem_u32(smem), 128u, 2304u);
    const uint64_t adesc = ptx::make_smem_desc_nosw(ptx::smem_u32(smem) + 40960u, 128u, 256u);
    uint32_t par = 0;
    long long chains = 0;
    const long long t0 = clock64();
    if (mmode != 0) {
      while (!stop) {
        if (ptx::elect_one_sync()) {
          for (int j = 0; j < 9; ++j) {
            if (mmode == 1 || (mmode == 3 && j > 0) || (mmode == 4 && j < 8))
              ptx::mma_f16_ts(tb + 256u, tb + 448u + (uint32_t)(j & 7) * 8u, bdesc + (uint64_t)(j * 16), idesc, j > 0);
            else ptx::mma_f16_ss(tb + 256u, adesc, bdesc + (uint64_t)(j * 16), idesc, j > 0);
          }
          ptx::mma_commit(&bar);
        }
        __syncwarp();
        ptx::mbar_wait(&bar, par);
        par ^= 1u;
        ++chains;
      }
    }
    const long long t1 = clock64();
    if (lane == 0) { out[2] = chains; out[3] = t1 - t0; }
  } else {
    const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    uint32_t acc = 0;
    // let the MMA warp get going
    const long long ts = clock64();
    while (clock64() - ts < 20000) { }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
      if (wmode == 0) { ptx::tmem_ld_x32(base, v); ptx::tc_wait_ld(); acc += v[0] + v[31]; }
      if (wmode == 1) {
        ptx::tmem_ld_x16(base, reinterpret_cast<uint32_t(&)[16]>(v[0])); ptx::tmem_ld_x16(base + 16u, reinterpret_cast<uint32_t(&)[16]>(v[16]));
        ptx::tc_wait_ld(); acc += v[0] + v[31];
      }
      if (wmode == 2) { ptx::tmem_st_x16(base + 128u, reinterpret_cast<uint32_t(&)[16]>(v[0])); ptx::tc_wait_st(); }
      if (wmode == 3) {
        ptx::tmem_st_x16(base + 128u, reinterpret_cast<uint32_t(&)[16]>(v[0])); ptx::tmem_ld_x32(base, v);
        ptx::tc_wait_st(); ptx::tc_wait_ld(); acc += v[0] + v[31];
      }
      // some ALU work between accesses, like an epilogue (keeps the access rate realistic: ~1 per 300 cycles)
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = v[i] * 1664525u + 1013904223u;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; }
    sink[threadIdx.x] = acc + v[3];
    __syncwarp();
    if (warp == 0 && lane == 0) { __threadfence_block(); }
  }
  // workers done -> stop the MMA warp
  if (warp < 16) asm volatile("bar.sync 1, 512;" ::: "memory");   // the 16 worker warps
  if (threadIdx.x == 0) { stop = 1; __threadfence_block(); }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 16) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tb); }
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* wn[] = {"ld x32 + wait", "2 x ld x16 + wait", "st x16 + wait", "st x16 ; ld x32 ; wait both"};
  const char* mn[] = {"no MMA", "TS MMA chains", "SS MMA chains", "1 SS + 8 TS", "8 TS + 1 SS"};
  for (int w = 0; w < 1; ++w)
    for (int m = 0; m < 5; ++m) {
      cudaMemset(out, 0, 64);
      bench<<<1, 17 * 32, 64 * 1024>>>(out, w, m, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
      printf("%-28s | %-14s : %7.1f cyc/iter (16 warps)   mma: %lld chains, %.1f cyc/chain %s\n", wn[w], mn[m], (double)h[0] / kIters,
             h[2], h[2] ? (double)h[3] / h[2] : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
