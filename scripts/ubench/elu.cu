// Micro-benchmark: the ELU epilogue arithmetic of tc_mlp_kernel (elu_pack32) on register data, 16 warps per SM
// (4 per scheduler) -- cycles per 32-element job per warp, with and without the MUFU.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../go2_onnx_controller_b200/csrc/kernels_tc.cuh"
using namespace go2p;

template <int VARIANT>
__global__ void __launch_bounds__(576, 1) bench(uint32_t* out, long long* cyc, float c, int iters) {
  uint32_t v[32], p[16];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-0.01f * (threadIdx.x + i));
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x < 512) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      if (VARIANT == 0) elu_pack32<true>(v, true, c, p);
      if (VARIANT == 1) elu_pack32<true>(v, false, c, p);
#pragma unroll
      for (int j = 0; j < 16; ++j) { acc ^= p[j]; v[2 * j] += (p[j] & 1u); }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 576 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int var = 0; var < 2; ++var) {
    for (int rep = 0; rep < 2; ++rep) {
      if (var == 0) bench<0><<<148, 576>>>(out, cyc, 1.4427f, 1000); else bench<1><<<148, 576>>>(out, cyc, 1.4427f, 1000);
      cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double a = 0; for (int i = 0; i < 148; ++i) a += h[i]; a /= 148;
    printf("%s: %.1f cycles per 32-element job per warp (16 warps/SM) -> %.1f cycles per 128x128 layer-tile\n", var == 0 ? "ELU (MUFU+FFMA+pack+select)" : "pack only", a / 1000, a / 1000);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
