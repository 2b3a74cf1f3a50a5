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
//     accumulator and 72 columns of 16-bit A operand.  The activation of layer l never leaves the SM:
//     epilogue warps read the accumulator with tcgen05.ld, apply ELU, pack to 16 bit and write it back
//     to TMEM with tcgen05.st, where the next layer's tcgen05.mma reads it as its A operand (TS form);
//   * the bias rides inside the MMA: every A operand carries two constant 1.0 columns and the weight
//     matrix two extra K rows holding hi/lo halves of the bias, so the epilogue has no bias add;
//   * the chain is evaluated in the base-2 exponent domain: layer l produces z' = log2(e)*z, the ELU is
//     h' = z' > 0 ? z' : c*(2^z' - 1) with c = alpha*log2(e) (one MUFU.EX2 + one FFMA), h' = log2(e)*h
//     feeds the next layer whose weights carry the inverse factor (folded on the host);
//   * the ELU select runs on packed 16-bit pairs (HSET2 + LOP3), so an element costs
//     1 MUFU + 1 FFMA + 2 ALU instructions; the MUFU pipe (16 lanes/clk/SM) is the per-SM floor;
//   * one elected thread issues every tcgen05.mma and signals completion with tcgen05.commit; while the
//     tensor core works on one slot the 8 epilogue warps of the other slot keep the MUFU pipe busy.
#pragma once
#include "policy_dev.cuh"
#include "ptx_sm100.cuh"

namespace go2p {

constexpr int kTcTileM = 128;
constexpr int kTcHidden = 128;     // every hidden width handled by this kernel
constexpr int kTcOutPad = 16;      // last layer N padded to 16 (smallest UMMA N at M=128)
constexpr int kTcBiasK = 16;       // extra K block carrying the two constant-one columns (bias hi / lo)
constexpr int kTcWorkers = 16;     // warps 0..7 serve slot 0, 8..15 slot 1: 4 TMEM lane quarters (warp % 4) x 2 column halves
constexpr int kTcPool = 8;         // worker warps per slot
constexpr int kTcProducerWarp = 16;  // bulk-copy producer
constexpr int kTcMmaWarp = 17;       // MMA issuer + TMEM owner
constexpr int kTcThreads = (kTcWorkers + 2) * 32;
constexpr int kTcSlotCols = 256;   // TMEM columns per slot: [0,128) accumulator, [128,200) A operand

struct TcArgs {
  const float* obs;          // [B, in_dim]
  float* act;                // [B, out_dim]
  const int32_t* button0;    // [B] or null
  double* qdes;              // [B, 12] or null
  long long B;
  const uint16_t* wpack;     // per layer: N_l x Kp_l 16-bit, UMMA K-major interleaved layout (bias rows included)
  int n_layers, in_dim, k0p, out_dim;
  int has_elu[kMaxLayers];
  float elu_c[kMaxLayers];   // alpha * log2(e)
  float out_scale;           // ln(2) if the last layer has an ELU (its output is in the scaled domain), else 1
  uint32_t flags;
  float action_limit;
  double action_scale;
  double q0[kDof];
  unsigned long long* trace;   // debug timeline (GO2P_TC_TRACE): [0] = count, then (event, clock64) pairs; CTA 0 only
};

#ifdef GO2P_TC_TRACE
// store-only tracer: every tracing warp owns a 1024-entry lane of the buffer and a local counter (no atomics,
// so a trace point costs a clock read and one fire-and-forget store)
struct TcTracer {
  unsigned long long* base; uint32_t n;
  __device__ void init(const TcArgs& a, int warp) { base = (a.trace && blockIdx.x == 0) ? a.trace + (size_t)warp * 2048 : nullptr; n = 0; }
  __device__ __forceinline__ void put(uint32_t ev) {
    if (base && n < 1023) { base[2 * n] = ev; base[2 * n + 1] = (unsigned long long)clock64(); ++n; base[2046] = n; }
  }
};
#define TC_TRACE_INIT() TcTracer tracer; tracer.init(a, warp)
#define TC_TRACE(ev) tracer.put(ev)
#else
#define TC_TRACE_INIT() do { } while (0)
#define TC_TRACE(ev) do { } while (0)
#endif

// K extent of layer l's MMA (including the bias block) and its N
__host__ __device__ inline int tc_layer_kp(const TcArgs& a, int l) { return l == 0 ? a.k0p : kTcHidden + kTcBiasK; }
__host__ __device__ inline int tc_layer_n(const TcArgs& a, int l) { return l == a.n_layers - 1 ? kTcOutPad : kTcHidden; }
__host__ __device__ inline size_t tc_weight_bytes(const TcArgs& a) {
  size_t s = 0;
  for (int l = 0; l < a.n_layers; ++l) s += (size_t)tc_layer_kp(a, l) * tc_layer_n(a, l) * 2;
  return s;
}
__host__ __device__ inline size_t tc_stage_bytes(const TcArgs& a) { return ((size_t)kTcTileM * a.in_dim * 4 + 127) & ~(size_t)127; }
__host__ __device__ inline size_t tc_smem_bytes(const TcArgs& a) { return tc_weight_bytes(a) + 2 * tc_stage_bytes(a) + 128; }

// ELU in the base-2 domain on 32 accumulator columns -> 16 words of packed 16-bit operands.
//   e = 2^z' (MUFU), f = c*e - c (FFMA), result = z' < 0 ? f : z' selected on the packed pair.
template <bool kFp16>
__device__ __forceinline__ void elu_pack32(const uint32_t (&v)[32], bool has_elu, float c, uint32_t (&p)[16]) {
  if (has_elu) {
    const float nc = -c;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float z0 = __uint_as_float(v[2 * j]), z1 = __uint_as_float(v[2 * j + 1]);
      const float f0 = fmaf(ptx::ex2_approx(z0), c, nc);
      const float f1 = fmaf(ptx::ex2_approx(z1), c, nc);
      if (kFp16) {
        const uint32_t zp = ptx::pack_f16_sat(z0, z1), fp = ptx::pack_f16_sat(f0, f1);
        p[j] = ptx::select_neg_f16x2(zp, fp);
      } else {
        const uint32_t zp = ptx::pack_bf16(z0, z1), fp = ptx::pack_bf16(f0, f1);
        p[j] = ptx::select_neg_bf16x2(zp, fp);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float z0 = __uint_as_float(v[2 * j]), z1 = __uint_as_float(v[2 * j + 1]);
      p[j] = kFp16 ? ptx::pack_f16_sat(z0, z1) : ptx::pack_bf16(z0, z1);
    }
  }
}

template <bool kFp16>
__global__ void __launch_bounds__(kTcThreads, 1) tc_mlp_kernel(const TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const size_t wbytes = tc_weight_bytes(a);
  const size_t stage_bytes = tc_stage_bytes(a);
  uint8_t* w_smem = smem;
  uint8_t* stage0 = smem + wbytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + 2 * stage_bytes);
  uint64_t* obs_full = bars;        // [2]
  uint64_t* obs_empty = bars + 2;   // [2]
  uint64_t* a_ready = bars + 4;     // [2]
  uint64_t* acc_full = bars + 6;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  // ---- one-time setup: weights -> smem, barriers, TMEM
  {
    const int4* src = reinterpret_cast<const int4*>(a.wpack);
    int4* dst = reinterpret_cast<int4*>(w_smem);
    const int n16 = (int)(wbytes >> 4);
    for (int i = tid; i < n16; i += kTcThreads) dst[i] = src[i];
  }
  if (warp == kTcMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&obs_full[s], 1);
        ptx::mbar_init(&obs_empty[s], kTcPool);
        ptx::mbar_init(&a_ready[s], kTcPool);
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
  TC_TRACE_INIT();

  if (warp == kTcProducerWarp) {
    // ================= producer: bulk async copies of observation slabs =================
    // (the whole warp walks the loop converged; one elected lane issues)
    for (int i = 0; i < n_local; ++i) {
      const int s = i & 1, n = i >> 1;
      ptx::mbar_wait(&obs_empty[s], (uint32_t)((n & 1) ^ 1));
      const long long tile = blockIdx.x + (long long)i * gridDim.x;
      const long long row0 = tile * kTcTileM;
      if (ptx::elect_one_sync()) {
        if (a.B - row0 >= kTcTileM) {
          TC_TRACE(0x100u | (uint32_t)s);
          ptx::mbar_arrive_expect_tx(&obs_full[s], tile_bytes);
          ptx::bulk_g2s(stage0 + s * stage_bytes, a.obs + row0 * a.in_dim, tile_bytes, &obs_full[s]);
        } else {
          ptx::mbar_arrive(&obs_full[s]);   // ragged last tile: consumers read global memory directly
        }
      }
      __syncwarp();
    }
  } else if (warp == kTcMmaWarp) {
    // ================= MMA issuer: one elected thread drives the tensor core =================
    // The two slots are served in whatever order their A operands become ready (non-blocking polls), so the
    // pools drift into anti-phase: one computes its ELU while the tensor core runs the other's layer.
    const uint32_t fmt = kFp16 ? ptx::FMT_F16 : ptx::FMT_BF16;
    const uint32_t w_base = ptx::smem_u32(w_smem);
    uint32_t par[2] = {0u, 0u};
    int lay[2] = {0, 0};
    uint32_t woff[2] = {0u, 0u};
    int left[2] = {(n_local + 1) / 2, n_local / 2};       // tiles still to issue per slot
    int s = 0;
    while (left[0] > 0 || left[1] > 0) {
      if (left[s] > 0) {
        const bool ready = __all_sync(0xffffffffu, ptx::mbar_try_wait(&a_ready[s], par[s]));
        if (ready) {
          par[s] ^= 1u;
          ptx::tc_fence_after();
          const int l = lay[s];
          const int kp = tc_layer_kp(a, l), nl = tc_layer_n(a, l);
          const uint32_t idesc = ptx::make_idesc(fmt, kTcTileM, (uint32_t)nl);
          // K step of 16 elements = two 8x16B core matrices along K (LBO = 128 B apart); 8-row groups are
          // kp*16 B apart (SBO); consecutive K steps are 256 B apart (+16 in the descriptor's address field)
          const uint64_t bdesc0 = ptx::make_smem_desc_nosw(w_base + woff[s], 128u, (uint32_t)kp * 16u);
          const int ksteps = kp / 16;
          const uint32_t d_t = tmem_base + (uint32_t)s * kTcSlotCols;
          const uint32_t a_t = d_t + 128u;
          if (ptx::elect_one_sync()) {
            TC_TRACE(0x200u | (uint32_t)(l << 4) | (uint32_t)s);
            ptx::mma_f16_ts(d_t, a_t, bdesc0, idesc, 0u);
            for (int j = 1; j < ksteps; ++j)
              ptx::mma_f16_ts(d_t, a_t + (uint32_t)j * 8u, bdesc0 + (uint64_t)(j * 16), idesc, 1u);
            ptx::mma_commit(&acc_full[s]);
            TC_TRACE(0x300u | (uint32_t)(l << 4) | (uint32_t)s);
          }
          __syncwarp();
          woff[s] += (uint32_t)(kp * nl * 2);
          if (++lay[s] == a.n_layers) { lay[s] = 0; woff[s] = 0u; --left[s]; }
        }
      }
      s ^= 1;
    }
  } else {
    // ================= worker warps: one pool of 8 warps per slot =================
    // pool = 4 TMEM lane quarters x 2 column halves; every warp owns 64 columns of each hidden layer (two
    // 32-column chunks, both loaded up front so the second load and the first store overlap the arithmetic).
    const int s = warp >> 3;                 // slot == pool
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch (== warp % 4)
    const int half = (warp >> 2) & 1;        // column half
    const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
    const uint32_t acc_t = tmem_base + (uint32_t)s * kTcSlotCols + lane_addr;
    const uint32_t a_t = acc_t + 128u;
    const int m = quarter * 32 + lane;       // row inside the tile
    const uint32_t one2 = kFp16 ? 0x3C003C00u : 0x3F803F80u;   // packed (1.0, 1.0)
    const int L = a.n_layers - 1;            // index of the output layer
    const bool out12 = a.out_dim == 12;
    const int n8 = a.k0p / 16;               // layer-0 A operand: chunks of 8 packed columns (16 elements)
    const bool even = (a.in_dim & 1) == 0;

    // constant-one columns of the hidden-layer A operand (K = 128,129), zeros up to K = 143: written once
    if (half == 0) {
      const uint32_t ones[8] = {one2, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      ptx::tmem_st_x8(a_t + (uint32_t)(kTcHidden / 2), ones);
      ptx::tc_wait_st();
    }

    uint32_t par_acc = 0u;
    int n = 0;
    for (int i = s; i < n_local; i += 2, ++n) {
      const long long row0 = (blockIdx.x + (long long)i * gridDim.x) * kTcTileM;
      const int valid = (int)min((long long)kTcTileM, a.B - row0);

      // ---- conv: fp32 observation rows -> 16-bit layer-0 A operand in TMEM, constant ones at K = in_dim, in_dim+1
      ptx::mbar_wait(&obs_full[s], (uint32_t)(n & 1));
      TC_TRACE(0x400u | (uint32_t)s);
      if (valid == kTcTileM && even) {
        // full tile in shared memory, rows 8-byte aligned: vector loads, uniform per-pair classification
        const float2* r2 = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(stage0 + s * stage_bytes) + (size_t)m * a.in_dim);
#pragma unroll
        for (int it = 0; it < 5; ++it) {
          const int c8 = half + 2 * it;
          if (c8 < n8) {
            uint32_t q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = c8 * 16 + 2 * j;
              if (k < a.in_dim) { const float2 t = r2[k >> 1]; q[j] = kFp16 ? ptx::pack_f16_sat(t.x, t.y) : ptx::pack_bf16(t.x, t.y); }
              else q[j] = (k == a.in_dim) ? one2 : 0u;
            }
            ptx::tmem_st_x8(a_t + (uint32_t)c8 * 8u, q);
          }
        }
      } else {
        // ragged last tile (read straight from global memory) or odd input width: scalar path
        const float* rowp = (valid == kTcTileM) ? reinterpret_cast<const float*>(stage0 + s * stage_bytes) + (size_t)m * a.in_dim
                                                : a.obs + (row0 + m) * a.in_dim;
        const bool live = m < valid;
        for (int c8 = half; c8 < n8; c8 += 2) {
          uint32_t q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = c8 * 16 + 2 * j;
            float lo = 0.f, hi = 0.f;
            if (k < a.in_dim) { if (live) lo = rowp[k]; } else if (k <= a.in_dim + 1) lo = 1.f;
            if (k + 1 < a.in_dim) { if (live) hi = rowp[k + 1]; } else if (k + 1 <= a.in_dim + 1) hi = 1.f;
            q[j] = kFp16 ? ptx::pack_f16_sat(lo, hi) : ptx::pack_bf16(lo, hi);
          }
          ptx::tmem_st_x8(a_t + (uint32_t)c8 * 8u, q);
        }
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { ptx::mbar_arrive(&obs_empty[s]); ptx::mbar_arrive(&a_ready[s]); }
      TC_TRACE(0x500u | (uint32_t)s);

      // ---- hidden layers: accumulator -> ELU -> 16-bit A operand of the next layer (64 columns per warp)
      for (int l = 0; l < L; ++l) {
        const bool he = a.has_elu[l] != 0;
        const float c = a.elu_c[l];
        uint32_t v0[32], v1[32], p0[16], p1[16];
        ptx::mbar_wait(&acc_full[s], par_acc);
        par_acc ^= 1u;
        ptx::tc_fence_after();
        ptx::tmem_ld_x32(acc_t + (uint32_t)(half * 64), v0);
        ptx::tmem_ld_x32(acc_t + (uint32_t)(half * 64 + 32), v1);
        ptx::tc_wait_ld();
        TC_TRACE(0x600u | (uint32_t)(l << 4) | (uint32_t)s);
        elu_pack32<kFp16>(v0, he, c, p0);
        ptx::tmem_st_x16(a_t + (uint32_t)(half * 32), p0);
        elu_pack32<kFp16>(v1, he, c, p1);
        ptx::tmem_st_x16(a_t + (uint32_t)(half * 32 + 16), p1);
        TC_TRACE(0xA00u | (uint32_t)(l << 4) | (uint32_t)s);
        ptx::tc_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_ready[s]);
        TC_TRACE(0x700u | (uint32_t)(l << 4) | (uint32_t)s);
      }

      // ---- output layer (bias already inside the accumulator): (+ELU) (+clamp/mask) (+q_des) -> global
      // 12 outputs: half 0 stores columns 0..7, half 1 columns 8..11; other widths: half 0 does all 16 columns
      ptx::mbar_wait(&acc_full[s], par_acc);
      par_acc ^= 1u;
      ptx::tc_fence_after();
      TC_TRACE(0x800u | (uint32_t)s);
      {
        uint32_t v[16];
        const int col0 = out12 ? half * 8 : 0;
        const int nv = out12 ? (half == 0 ? 8 : 4) : 16;
        const bool active = out12 || half == 0;
        if (active) {
          if (nv == 16) ptx::tmem_ld_x16(acc_t, v);
          else if (nv == 8) ptx::tmem_ld_x8(acc_t, reinterpret_cast<uint32_t(&)[8]>(v));
          else ptx::tmem_ld_x4(acc_t + 8u, reinterpret_cast<uint32_t(&)[4]>(v));
        }
        ptx::tc_wait_ld();
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = __uint_as_float(v[j]);
          if (a.has_elu[L]) x = ((x < 0.f) ? fmaf(ptx::ex2_approx(x), a.elu_c[L], -a.elu_c[L]) : x) * a.out_scale;
          o[j] = x;
        }
        if (active && m < valid) {
          const long long row = row0 + m;
          if (a.flags & 1u) {
            const int b0 = a.button0 ? a.button0[row] : 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < nv) o[j] = clamp_mask(o[j], a.action_limit, b0);
          }
          float* dst = a.act + row * a.out_dim + col0;
          if (out12) {
            reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
            if (half == 0) reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
            if ((a.flags & 2u) && a.qdes) {
              double2* q2 = reinterpret_cast<double2*>(a.qdes + row * kDof + col0);
#pragma unroll
              for (int j = 0; j < 8; j += 2)
                if (j < nv) q2[j >> 1] = make_double2(joint_target(o[j], a.q0[col0 + j], a.action_scale), joint_target(o[j + 1], a.q0[col0 + j + 1], a.action_scale));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < a.out_dim) dst[j] = o[j];
          }
        }
      }
      ptx::tc_fence_before();
      TC_TRACE(0x900u | (uint32_t)s);
    }
  }

  // ---- teardown
  ptx::tc_fence_before();
  block_sync();
  if (warp == kTcMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace go2p
