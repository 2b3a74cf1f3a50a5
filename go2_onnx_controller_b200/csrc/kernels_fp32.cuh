// Batched policy forward in plain fp32 on the CUDA cores -- the path that has to agree with the
// reference's ONNX Runtime CPU execution to 1e-5 (reference: onnx_actor.cpp:38-48), for any layer
// widths.  One launch per Gemm(+Elu) layer; activations ping-pong through a scratch buffer that stays
// L2 resident for the chunk sizes the host uses.
//   sgemm_bias_act_kernel : C[M,N] = act(A[M,K] * W^T + b), 128x128x8 tiles, 8x8 register tiles,
//                           register-prefetched double buffering
//   small_out_kernel      : the narrow output layer (N <= 32) with the fused A9/A11 epilogue
//                           (clamp + dead-man mask, q_des; reference: controller.cpp:217-223,244)
//   assemble_batch_kernel : A1-A6 for B robots (reference: controller.cpp:173-212)
#pragma once
#include <algorithm>
#include "kernels_b1.cuh"
#include "policy_dev.cuh"

namespace go2p {

constexpr int kSgBM = 128, kSgBN = 128, kSgBK = 8, kSgThreads = 256;

// A: [M,K] row-major (lda), Wkn: [Kp8][Np] k-major zero padded, bias: [Np] zero padded, C: [M,N] (ldc)
__global__ void __launch_bounds__(kSgThreads, 2)
sgemm_bias_act_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Wkn, int Np,
                      const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int K, int N,
                      int has_elu, float alpha) {
  __shared__ __align__(16) float As[2][kSgBK][kSgBM + 4];
  __shared__ __align__(16) float Bs[2][kSgBK][kSgBN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, each an 8x8 tile (2x2 blocks of 4x4)
  const int row0 = blockIdx.x * kSgBM, col0 = blockIdx.y * kSgBN;
  // A loader: thread -> (row = tid/2, 4 consecutive k)   B loader: thread -> (k = tid/32, 4 consecutive n)
  const int a_r = tid >> 1, a_k = (tid & 1) * 4;
  const int b_k = tid >> 5, b_n = (tid & 31) * 4;
  const int Kt = (K + kSgBK - 1) / kSgBK;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4];
  float4 rb;
  auto gload = [&](int kt) {
    const int k0 = kt * kSgBK;
    const int gr = row0 + a_r;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + a_k + j;
      ra[j] = (gr < M && k < K) ? A[(size_t)gr * lda + k] : 0.f;
    }
    rb = *reinterpret_cast<const float4*>(Wkn + (size_t)(k0 + b_k) * Np + col0 + b_n);   // padded: always in range
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[buf][a_k + j][a_r] = ra[j];
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb;
  };

  gload(0);
  sstore(0);
  block_sync();
  for (int kt = 0; kt < Kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < Kt) gload(kt + 1);
#pragma unroll
    for (int kk = 0; kk < kSgBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < Kt) {
      sstore(buf ^ 1);
      block_sync();
    }
  }
  // epilogue: rows {ty*4..+3, 64+ty*4..+3}, cols {tx*4..+3, 64+tx*4..+3}
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = row0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c = col0 + jh * 64 + tx * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][jh * 4 + j] + bias[c + j];
        if (has_elu) v = elu_exact(v, alpha);
        o[j] = v;
      }
      float* dst = C + (size_t)r * ldc + c;
      if (c + 3 < N && ((ldc & 3) == 0)) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (c + j < N) dst[j] = o[j];
      }
    }
  }
}

// Output layer with N <= 32: a CTA stages 128 rows of A in shared memory (odd pitch: conflict-free row
// walks), each thread owns one row and all N outputs; W (row-major [N][Kp]) is read as warp-uniform
// broadcasts.  Epilogue: bias (+ELU) (+clamp/mask) (+q_des).
constexpr int kSoRows = 128;
__global__ void __launch_bounds__(kSoRows, 1)
small_out_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Wrm, int Kp,
                 const float* __restrict__ bias, float* __restrict__ C, int ldc, long long M, int K, int N,
                 int has_elu, float alpha, uint32_t flags, const int32_t* __restrict__ button0,
                 double* __restrict__ qdes, CtrlConst cc) {
  extern __shared__ __align__(16) float so_sm[];
  const int pitch = K | 1;
  float* As = so_sm;                       // [128][pitch]
  float* Ws = so_sm + kSoRows * pitch;     // [N][Kp]
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * kSoRows;
  const int rows = (int)min((long long)kSoRows, M - row0);
  for (int i = tid; i < rows * K; i += kSoRows) {
    const int r = i / K, k = i - r * K;
    As[r * pitch + k] = A[(size_t)(row0 + r) * lda + k];
  }
  for (int i = tid; i < N * Kp; i += kSoRows) Ws[i] = Wrm[i];
  block_sync();
  if (tid >= rows) return;
  const float* x = As + tid * pitch;
  float acc[32];
#pragma unroll
  for (int n = 0; n < 32; ++n) acc[n] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float xv = x[k];
#pragma unroll
    for (int n = 0; n < 32; ++n)
      if (n < N) acc[n] = fmaf(xv, Ws[n * Kp + k], acc[n]);
  }
  const long long row = row0 + tid;
  const int b0 = (button0 && (flags & 1u)) ? button0[row] : 0;
#pragma unroll
  for (int n = 0; n < 32; ++n) {
    if (n >= N) break;
    float v = acc[n] + bias[n];
    if (has_elu) v = elu_exact(v, alpha);
    if (flags & 1u) v = clamp_mask(v, cc.action_limit, b0);
    C[(size_t)row * ldc + n] = v;
    if ((flags & 2u) && qdes && n < kDof) qdes[(size_t)row * kDof + n] = joint_target(v, cc.q0[n], cc.action_scale);
  }
}

// Elementwise A9/A11 for models whose output layer is too wide for small_out_kernel.
__global__ void post_kernel(float* __restrict__ act, long long total, int out_dim, uint32_t flags,
                            const int32_t* __restrict__ button0, double* __restrict__ qdes, CtrlConst cc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long row = i / out_dim;
  const int n = (int)(i - row * out_dim);
  float v = act[i];
  if (flags & 1u) { v = clamp_mask(v, cc.action_limit, button0 ? button0[row] : 0); act[i] = v; }
  if ((flags & 2u) && qdes && n < kDof) qdes[row * kDof + n] = joint_target(v, cc.q0[n], cc.action_scale);
}

// A1-A6 for B robots: one thread per observation element, history kept in the obs rows themselves
// (term-major, oldest frame first; reference: controller.cpp:200-212).  raw words as in kernels_b1.cuh.
struct RawStateDev {
  float quat[4]; float gyro[3]; float q[12]; float dq[12]; float axes[4];
  int16_t foot_force[4]; int32_t joy_valid; int32_t button0;
};
static_assert(sizeof(RawStateDev) == 4 * 35 + 8 + 8, "must match go2p_raw_state");

// button0_out (optional): the dead-man button of every robot as int32 [B], the form the batched epilogue reads.
// One WARP per robot row.  Lane i owns elements i, i+32, ... of the 49*H-float row; it reads all of them (old
// frames from the row itself, the newest frame from the raw state) into registers and only then writes, so the
// in-place shift needs no block barrier and rows move as contiguous segments.  The 156-byte raw state is loaded
// once per row as 39 coalesced words and handed out by warp shuffles; the gravity projection and the joystick
// command are evaluated by every lane (no divergence), with the same operations as the batch-1 kernel
// (kernels_b1.cuh: gravity_component, vel_cmd_component), so results are bit-identical to it.
constexpr int kAsmWarps = 8;

template <int kPerLane>
__global__ void __launch_bounds__(kAsmWarps * 32)
assemble_batch_kernel(const RawStateDev* __restrict__ raw, const float* __restrict__ prev_action,
                      float* __restrict__ vel_cmd, float* __restrict__ obs, long long B, CtrlConst cc,
                      int32_t* __restrict__ button0_out) {
  constexpr unsigned kFull = 0xffffffffu;
  const int H = cc.H, n_obs = kFrame * H;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kAsmWarps + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * kAsmWarps;
  // element classification is row-independent: hoisted out of the row loop.
  // meta = -1: no element; bit 8: old frame, low byte = distance to the value to copy; else newest frame:
  // term in bits 4..6, component in bits 0..3
  int meta[kPerLane];
  double q0e[kPerLane];                                    // default pose of the joint behind a (q - q0) element
#pragma unroll
  for (int e = 0; e < kPerLane; ++e) {
    const int tid = e * 32 + lane;
    meta[e] = -1;
    q0e[e] = 0.0;
    if (tid < n_obs) {
      int t, off, wdt;
      if (tid < 9 * H) { t = tid / (3 * H); off = t * 3 * H; wdt = 3; }
      else if (tid < 45 * H) { t = 3 + (tid - 9 * H) / (12 * H); off = 9 * H + (t - 3) * 12 * H; wdt = 12; }
      else { t = 6; off = 45 * H; wdt = 4; }
      const int local = tid - off; const int f = local / wdt; const int c = local - f * wdt;
      meta[e] = (f != H - 1) ? (0x100 | wdt) : ((t << 4) | c);
      if (f == H - 1 && t == 3) q0e[e] = cc.q0[c];
    }
  }
  // software pipeline over rows: the raw words and the old-frame values of the NEXT row are requested before the
  // current row is evaluated, so every warp always has one row's worth of DRAM latency in flight
  uint32_t nw0 = 0u, nw1 = 0u;
  float nv[kPerLane];
  auto request = [&](long long row) {
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw + row);
    const float* o = obs + row * n_obs;
    nw0 = rw[lane];                                        // words 0..31: quat, gyro, q, dq, axes[0]
    nw1 = lane < 7 ? rw[32 + lane] : 0u;                   // words 32..38: axes[1..3], foot_force (2), joy_valid, button0
#pragma unroll
    for (int e = 0; e < kPerLane; ++e)
      nv[e] = (meta[e] >= 0 && (meta[e] & 0x100)) ? o[e * 32 + lane + (meta[e] & 0xff)] : 0.f;
  };
  if (warp0 < B) request(warp0);
  for (long long row = warp0; row < B; row += n_warps) {
    float* o = obs + row * n_obs;
    const uint32_t w0 = nw0, w1 = nw1;
    float v[kPerLane];
#pragma unroll
    for (int e = 0; e < kPerLane; ++e) v[e] = nv[e];
    if (row + n_warps < B) request(row + n_warps);
    float quat[4], axes[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) quat[i] = __uint_as_float(__shfl_sync(kFull, w0, i));
    axes[0] = __uint_as_float(__shfl_sync(kFull, w0, 31));
#pragma unroll
    for (int i = 1; i < 4; ++i) axes[i] = __uint_as_float(__shfl_sync(kFull, w1, i - 1));
    const uint32_t ff01 = __shfl_sync(kFull, w1, 3), ff23 = __shfl_sync(kFull, w1, 4);
    const int joy_valid = (int)__shfl_sync(kFull, w1, 5);
    const float g0 = gravity_component(quat, 0), g1 = gravity_component(quat, 1), g2 = gravity_component(quat, 2);
    float c0, c1, c2;
    if (joy_valid) { c0 = vel_cmd_component(axes, 0); c1 = vel_cmd_component(axes, 1); c2 = vel_cmd_component(axes, 2); }
    else { c0 = vel_cmd[row * 3 + 0]; c1 = vel_cmd[row * 3 + 1]; c2 = vel_cmd[row * 3 + 2]; }   // warp-uniform branch
#pragma unroll
    for (int e = 0; e < kPerLane; ++e) {
      const int mt = meta[e];
      const bool fresh = mt >= 0 && !(mt & 0x100);
      const int t = (mt >> 4) & 7, c = mt & 15;
      // raw word behind a copied term: gyro 4+c, q 7+c, dq 19+c (all within the first 32 words)
      const int src = t == 1 ? 4 + c : (t == 3 ? 7 + c : 19 + c);
      const float x = __uint_as_float(__shfl_sync(kFull, w0, fresh ? src : 0));
      if (fresh) {
        float r;
        if (t == 0) r = c == 0 ? g0 : (c == 1 ? g1 : g2);
        else if (t == 2) r = c == 0 ? c0 : (c == 1 ? c1 : c2);
        else if (t == 3) r = __double2float_rn(__dsub_rn((double)x, q0e[e]));
        else if (t == 5) r = prev_action ? prev_action[row * kDof + c] : 0.f;
        else if (t == 6) {
          const int p = c ^ 1;                                  // [1,0,3,2]
          const uint32_t wpair = (p >> 1) ? ff23 : ff01;
          const int f = (int)(int16_t)(uint16_t)(wpair >> (16 * (p & 1)));
          r = (f >= cc.foot_threshold) ? 1.0f : 0.0f;
        } else r = x;                                           // gyro, dq
        v[e] = r;
      }
    }
    __syncwarp();      // every lane holds its old values before any lane overwrites them
#pragma unroll
    for (int e = 0; e < kPerLane; ++e)
      if (meta[e] >= 0) o[e * 32 + lane] = v[e];
    if (lane < 3 && joy_valid) vel_cmd[row * 3 + lane] = lane == 0 ? c0 : (lane == 1 ? c1 : c2);
    if (lane == 6 && button0_out) button0_out[row] = (int32_t)w1;
  }
}

// n_obs = 49*H floats per row: 4 elements per lane serve H <= 2, 13 serve H <= 8
inline void launch_assemble_batch(const RawStateDev* raw, const float* prev_action, float* vel_cmd, float* obs, long long B,
                                  const CtrlConst& cc, int32_t* button0_out, int sm_count, cudaStream_t st) {
  const unsigned grid = (unsigned)std::min<long long>((B + kAsmWarps - 1) / kAsmWarps, (long long)sm_count * 32);
  if (kFrame * cc.H <= 128)
    assemble_batch_kernel<4><<<grid, kAsmWarps * 32, 0, st>>>(raw, prev_action, vel_cmd, obs, B, cc, button0_out);
  else
    assemble_batch_kernel<(kFrame * kMaxHistory + 31) / 32><<<grid, kAsmWarps * 32, 0, st>>>(raw, prev_action, vel_cmd, obs, B, cc, button0_out);
}

}  // namespace go2p
