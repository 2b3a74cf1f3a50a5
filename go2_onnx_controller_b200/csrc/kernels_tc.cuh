// Batched policy forward on the 5th-gen tensor cores (tcgen05 + TMEM), one persistent CTA per SM.
//
//   A7  Gemm/Elu chain   reference: onnx_actor.cpp:38-48 (Ort::Session::Run), here for B rows at once
//   A9  clamp + mask     reference: controller.cpp:217-223   (optional fused epilogue)
//   A11 joint targets    reference: controller.cpp:244       (optional fused epilogue)
//
// Design (see DESIGN.md "tc_mlp_kernel"):
//   * all layer weights live in shared memory for the life of the CTA as 16-bit K-major UMMA
//     "interleaved" core matrices (packed once on the host at create time);
//   * observations stream in as 1-D bulk async copies (TMA engine) of whole 128-row slabs
//     ([128,in] fp32 is contiguous; its 392-byte row pitch rules out a 2-D tensor map) into a
//     2-stage shared-memory ring guarded by mbarriers;
//   * two 128-row tiles are in flight per CTA ("slots").  Each slot owns 128 TMEM columns of fp32
//     accumulator and 64 columns of 16-bit A operand.  The activation of layer l never leaves the SM:
//     epilogue warps read the accumulator with tcgen05.ld, add bias, apply ELU, pack to 16 bit and
//     write it back to TMEM with tcgen05.st, where the next layer's tcgen05.mma reads it as its A
//     operand (TS form) -- no shared-memory round trip and half the operand traffic of the SS form;
//   * one elected thread issues every tcgen05.mma and signals completion with tcgen05.commit; while the
//     tensor core works on one slot the epilogue warps of the other slot run on the CUDA cores.
#pragma once
#include "policy_dev.cuh"
#include "ptx_sm100.cuh"

namespace go2p {

constexpr int kTcTileM = 128;
constexpr int kTcHidden = 128;     // every hidden width handled by this kernel
constexpr int kTcOutPad = 16;      // last layer N padded to 16 (smallest UMMA N at M=128)
constexpr int kTcCtrlWarps = 4;    // warp 0 producer, warp 1 MMA issuer + TMEM owner, 2-3 spare

struct TcArgs {
  const float* obs;          // [B, in_dim]
  float* act;                // [B, out_dim]
  const int32_t* button0;    // [B] or null
  double* qdes;              // [B, 12] or null
  long long B;
  const uint16_t* wpack;     // per layer: N_l x Kp_l 16-bit, UMMA K-major interleaved layout
  const float* bias;         // per layer 128 floats (zero padded)
  int n_layers, in_dim, k0p, out_dim;
  int has_elu[kMaxLayers];
  float alpha[kMaxLayers];
  uint32_t flags;
  float action_limit;
  double action_scale;
  double q0[kDof];
};

__host__ __device__ inline int tc_layer_kp(const TcArgs& a, int l) { return l == 0 ? a.k0p : kTcHidden; }
__host__ __device__ inline int tc_layer_n(const TcArgs& a, int l) { return l == a.n_layers - 1 ? kTcOutPad : kTcHidden; }
__host__ __device__ inline size_t tc_weight_bytes(const TcArgs& a) {
  size_t s = 0;
  for (int l = 0; l < a.n_layers; ++l) s += (size_t)tc_layer_kp(a, l) * tc_layer_n(a, l) * 2;
  return s;
}
__host__ __device__ inline size_t tc_stage_bytes(const TcArgs& a) { return ((size_t)kTcTileM * a.in_dim * 4 + 127) & ~(size_t)127; }
__host__ __device__ inline size_t tc_smem_bytes(const TcArgs& a) {
  return tc_weight_bytes(a) + (size_t)a.n_layers * 128 * 4 + 2 * tc_stage_bytes(a) + 128;
}

// bias + ELU (exp via MUFU ex2) + pack 32 accumulator columns into 16 words of 2x16 bit
template <bool kFp16>
__device__ __forceinline__ void bias_elu_pack32(const uint32_t (&v)[32], const float* __restrict__ bias, bool has_elu,
                                                float alpha, uint32_t (&p)[16]) {
  const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = b4[q];
    float x0 = __uint_as_float(v[4 * q + 0]) + b.x;
    float x1 = __uint_as_float(v[4 * q + 1]) + b.y;
    float x2 = __uint_as_float(v[4 * q + 2]) + b.z;
    float x3 = __uint_as_float(v[4 * q + 3]) + b.w;
    if (has_elu) {
      const float kLog2e = 1.4426950408889634f;
      const float e0 = fmaf(alpha, ptx::ex2_approx(x0 * kLog2e), -alpha);
      const float e1 = fmaf(alpha, ptx::ex2_approx(x1 * kLog2e), -alpha);
      const float e2 = fmaf(alpha, ptx::ex2_approx(x2 * kLog2e), -alpha);
      const float e3 = fmaf(alpha, ptx::ex2_approx(x3 * kLog2e), -alpha);
      x0 = (x0 < 0.f) ? e0 : x0;
      x1 = (x1 < 0.f) ? e1 : x1;
      x2 = (x2 < 0.f) ? e2 : x2;
      x3 = (x3 < 0.f) ? e3 : x3;
    }
    if (kFp16) { p[2 * q] = ptx::pack_f16_sat(x0, x1); p[2 * q + 1] = ptx::pack_f16_sat(x2, x3); }
    else       { p[2 * q] = ptx::pack_bf16(x0, x1);    p[2 * q + 1] = ptx::pack_bf16(x2, x3); }
  }
}

template <bool kFp16, int kEpw>
__global__ void __launch_bounds__((kTcCtrlWarps + 2 * kEpw) * 32, 1) tc_mlp_kernel(const TcArgs a) {
  static_assert(kEpw == 4 || kEpw == 8, "epilogue warps per slot");
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthreads = (kTcCtrlWarps + 2 * kEpw) * 32;

  const size_t wbytes = tc_weight_bytes(a);
  const size_t stage_bytes = tc_stage_bytes(a);
  uint8_t* w_smem = smem;
  float* bias_smem = reinterpret_cast<float*>(smem + wbytes);
  uint8_t* stage0 = smem + wbytes + (size_t)a.n_layers * 128 * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + 2 * stage_bytes);
  uint64_t* obs_full = bars;        // [2]
  uint64_t* obs_empty = bars + 2;   // [2]
  uint64_t* a_ready = bars + 4;     // [2]
  uint64_t* acc_full = bars + 6;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  // ---- one-time setup: weights + biases -> smem, barriers, TMEM
  {
    const int4* src = reinterpret_cast<const int4*>(a.wpack);
    int4* dst = reinterpret_cast<int4*>(w_smem);
    const int n16 = (int)(wbytes >> 4);
    for (int i = tid; i < n16; i += nthreads) dst[i] = src[i];
    for (int i = tid; i < a.n_layers * 128; i += nthreads) bias_smem[i] = a.bias[i];
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&obs_full[s], 1);
        ptx::mbar_init(&obs_empty[s], kEpw);
        ptx::mbar_init(&a_ready[s], kEpw);
        ptx::mbar_init(&acc_full[s], 1);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::fence_proxy_async_smem();   // weights were written through the generic proxy; UMMA reads via the async proxy
  ptx::tc_fence_before();
  block_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const long long n_tiles = (a.B + kTcTileM - 1) / kTcTileM;
  const int n_local = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles b, b+G, ...
  const uint32_t tile_bytes = (uint32_t)(kTcTileM * a.in_dim * 4);

  if (warp == 0) {
    // ================= producer: bulk async copies of observation slabs =================
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int s = i & 1, n = i >> 1;
        ptx::mbar_wait(&obs_empty[s], (uint32_t)((n & 1) ^ 1));
        const long long tile = blockIdx.x + (long long)i * gridDim.x;
        const long long row0 = tile * kTcTileM;
        if (a.B - row0 >= kTcTileM) {
          ptx::mbar_arrive_expect_tx(&obs_full[s], tile_bytes);
          ptx::bulk_g2s(stage0 + s * stage_bytes, a.obs + row0 * a.in_dim, tile_bytes, &obs_full[s]);
        } else {
          ptx::mbar_arrive(&obs_full[s]);   // ragged last tile: consumers read global memory directly
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: one thread drives the tensor core =================
    if (lane == 0) {
      const uint32_t fmt = kFp16 ? ptx::FMT_F16 : ptx::FMT_BF16;
      uint32_t par[2] = {0u, 0u};
      const uint32_t w_base = ptx::smem_u32(w_smem);
      for (int pair = 0; pair * 2 < n_local; ++pair) {
        uint32_t w_off = 0;
        for (int l = 0; l < a.n_layers; ++l) {
          const int kp = tc_layer_kp(a, l), nl = tc_layer_n(a, l);
          const uint32_t idesc = ptx::make_idesc(fmt, kTcTileM, (uint32_t)nl);
          for (int s = 0; s < 2; ++s) {
            if (pair * 2 + s >= n_local) continue;
            ptx::mbar_wait(&a_ready[s], par[s]);
            par[s] ^= 1u;
            ptx::tc_fence_after();
            const uint32_t d_t = tmem_base + (uint32_t)s * 256u;
            const uint32_t a_t = d_t + 128u;
            for (int j = 0; j < kp / 16; ++j) {
              // K step of 16 elements = two 8x16B core matrices along K (LBO = 128 B apart);
              // 8-row groups are kp*16 B apart (SBO).
              const uint64_t bdesc = ptx::make_smem_desc_nosw(w_base + w_off + (uint32_t)j * 256u, 128u, (uint32_t)kp * 16u);
              ptx::mma_f16_ts(d_t, a_t + (uint32_t)j * 8u, bdesc, idesc, j > 0 ? 1u : 0u);
            }
            ptx::mma_commit(&acc_full[s]);
          }
          w_off += (uint32_t)(kp * nl * 2);
        }
      }
    }
  } else if (warp >= kTcCtrlWarps) {
    // ================= epilogue warps: 2 slots x kEpw warps =================
    const int ew = warp - kTcCtrlWarps;
    const int s = ew / kEpw;                 // slot
    const int wq = ew % kEpw;
    const int quarter = wq & 3;              // TMEM lane quarter this warp may touch (== warp % 4)
    const int half = wq >> 2;                // column half (only when kEpw == 8)
    constexpr int kHalves = kEpw / 4;
    const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
    const uint32_t acc_t = tmem_base + (uint32_t)s * 256u + lane_addr;
    const uint32_t a_t = acc_t + 128u;
    const int m = quarter * 32 + lane;       // row inside the tile
    uint32_t par_acc = 0u;
    int n = 0;
    for (int i = s; i < n_local; i += 2, ++n) {
      const long long tile = blockIdx.x + (long long)i * gridDim.x;
      const long long row0 = tile * kTcTileM;
      const int valid = (int)min((long long)kTcTileM, a.B - row0);
      const bool full = valid == kTcTileM;

      // ---- layer-0 A operand: fp32 observation row -> 16-bit -> TMEM
      ptx::mbar_wait(&obs_full[s], (uint32_t)(n & 1));
      {
        const float* rowp = full ? reinterpret_cast<const float*>(stage0 + s * stage_bytes) + (size_t)m * a.in_dim
                                 : a.obs + (row0 + m) * a.in_dim;
        const bool live = m < valid;
        const bool even = (a.in_dim & 1) == 0;
        const int n8 = a.k0p / 16;           // chunks of 8 packed columns (16 elements)
        const int c_lo = (n8 * half) / kHalves, c_hi = (n8 * (half + 1)) / kHalves;
        for (int c8 = c_lo; c8 < c_hi; ++c8) {
          uint32_t p[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = (c8 * 8 + j) * 2;
            float lo = 0.f, hi = 0.f;
            if (live) {
              if (even) {
                if (k < a.in_dim) { const float2 t = *reinterpret_cast<const float2*>(rowp + k); lo = t.x; hi = t.y; }
              } else {
                if (k < a.in_dim) lo = rowp[k];
                if (k + 1 < a.in_dim) hi = rowp[k + 1];
              }
            }
            p[j] = kFp16 ? ptx::pack_f16_sat(lo, hi) : ptx::pack_bf16(lo, hi);
          }
          ptx::tmem_st_x8(a_t + (uint32_t)c8 * 8u, p);
        }
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { ptx::mbar_arrive(&obs_empty[s]); ptx::mbar_arrive(&a_ready[s]); }

      // ---- hidden layers: accumulator -> bias + ELU -> 16-bit A operand of the next layer
      for (int l = 0; l < a.n_layers - 1; ++l) {
        ptx::mbar_wait(&acc_full[s], par_acc);
        par_acc ^= 1u;
        ptx::tc_fence_after();
        const float* bl = bias_smem + l * 128;
        const bool he = a.has_elu[l] != 0;
        const float al = a.alpha[l];
        constexpr int kChunks = 4 / kHalves;           // 32-column chunks this warp owns
        const int c0 = half * kChunks;
        uint32_t va[32], vb[32], p[16];
        ptx::tmem_ld_x32(acc_t + (uint32_t)c0 * 32u, va);
#pragma unroll
        for (int c = 0; c < kChunks; c += 2) {
          ptx::tc_wait_ld();
          if (c + 1 < kChunks) ptx::tmem_ld_x32(acc_t + (uint32_t)(c0 + c + 1) * 32u, vb);
          bias_elu_pack32<kFp16>(va, bl + (c0 + c) * 32, he, al, p);
          ptx::tmem_st_x16(a_t + (uint32_t)(c0 + c) * 16u, p);
          if (c + 1 < kChunks) {
            ptx::tc_wait_ld();
            if (c + 2 < kChunks) ptx::tmem_ld_x32(acc_t + (uint32_t)(c0 + c + 2) * 32u, va);
            bias_elu_pack32<kFp16>(vb, bl + (c0 + c + 1) * 32, he, al, p);
            ptx::tmem_st_x16(a_t + (uint32_t)(c0 + c + 1) * 16u, p);
          }
        }
        ptx::tc_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_ready[s]);
      }

      // ---- output layer: bias (+ELU) (+clamp/mask) (+q_des) -> global
      ptx::mbar_wait(&acc_full[s], par_acc);
      par_acc ^= 1u;
      ptx::tc_fence_after();
      if (half == 0) {
        uint32_t v[16];
        ptx::tmem_ld_x16(acc_t, v);
        ptx::tc_wait_ld();
        const int L = a.n_layers - 1;
        const float* bl = bias_smem + L * 128;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = __uint_as_float(v[j]) + bl[j];
          if (a.has_elu[L]) x = (x < 0.f) ? fmaf(a.alpha[L], ptx::ex2_approx(x * 1.4426950408889634f), -a.alpha[L]) : x;
          o[j] = x;
        }
        if (m < valid) {
          const long long row = row0 + m;
          if (a.flags & 1u) {
            const int b0 = a.button0 ? a.button0[row] : 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = clamp_mask(o[j], a.action_limit, b0);
          }
          float* dst = a.act + row * a.out_dim;
          if (a.out_dim == 12) {
            float4* d4 = reinterpret_cast<float4*>(dst);
            d4[0] = make_float4(o[0], o[1], o[2], o[3]);
            d4[1] = make_float4(o[4], o[5], o[6], o[7]);
            d4[2] = make_float4(o[8], o[9], o[10], o[11]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < a.out_dim) dst[j] = o[j];
          }
          if ((a.flags & 2u) && a.qdes) {
            double2* q2 = reinterpret_cast<double2*>(a.qdes + row * kDof);
#pragma unroll
            for (int j = 0; j < kDof; j += 2)
              q2[j >> 1] = make_double2(joint_target(o[j], a.q0[j], a.action_scale), joint_target(o[j + 1], a.q0[j + 1], a.action_scale));
          }
        }
      }
      ptx::tc_fence_before();
    }
  }

  // ---- teardown
  ptx::tc_fence_before();
  block_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace go2p
