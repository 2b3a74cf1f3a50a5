// Batched policy forward on the 5th-gen tensor cores (tcgen05 + TMEM), one persistent CTA per SM.
//
//   A7  Gemm/Elu chain   reference: onnx_actor.cpp:38-48 (Ort::Session::Run), here for B rows at once
//   A9  clamp + mask     reference: controller.cpp:217-223   (optional fused epilogue)
//   A11 joint targets    reference: controller.cpp:235-251   (optional fused epilogue, Isaac and Unitree motor order)
//
// Design (see DESIGN.md "tc_mlp_kernel"):
//   * all layer weights live in shared memory for the life of the CTA as 16-bit K-major UMMA
//     "interleaved" core matrices (packed once on the host at create time);
//   * observations stream in as 1-D bulk async copies (TMA engine) of whole 128-row slabs
//     ([128,in] fp32 is contiguous; its 392-byte row pitch rules out a 2-D tensor map) into a
//     2-stage shared-memory ring guarded by mbarriers;
//   * THREE 128-row tiles ("slots") are in flight per CTA.  TMEM (512 columns) holds one fp32 accumulator per slot
//     (3 x 128 columns) and two 64-column A-operand buffers shared by all slots: an epilogue job reads a slot's
//     accumulator with tcgen05.ld, applies the ELU, packs to 16 bit and writes the next layer's A operand into the
//     A buffer whose turn it is (tcgen05.st); that layer's tcgen05.mma (TS form: A from TMEM, weights from shared
//     memory) accumulates back into the slot's own accumulator.  An accumulator is only live between the MMA and the
//     epilogue's load, an A buffer only between the epilogue's store and the MMA -- which is why three tiles fit
//     where ping-pong accumulators allowed two.  The activation never leaves the SM;
//   * ONE pool of 16 worker warps walks a STATIC job schedule built on the host (TcArgs::sched): round-robin over the
//     slots, every slot cycling through conv, E(layer 0..L-1), out, the slots staggered by a third of a tile so the
//     conversions (and the bulk copies behind them) are spread out.  A slot's MMA group is issued when its job
//     completes and has two whole jobs of cover before the pool returns to that slot: the MMA hand-off (~1,400
//     cycles, the term that bound the two-slot design) is off the critical path;
//   * the pool is software-pipelined across jobs: the next job's accumulator load (and the out job's button load)
//     is issued before the current job's tcgen05.st is drained and signalled;
//   * the bias rides inside the MMA: hidden layers get one extra SS-form K step whose A operand is a constant
//     128 x 16 tile (two columns of ones) in shared memory and whose B rows hold hi/lo halves of the bias; layer 0
//     carries the two ones columns inside its K padding.  The epilogue has no bias add;
//   * the chain is evaluated in the base-2 exponent domain: layer l produces z' = log2(e)*z, the ELU is
//     h' = z' > 0 ? z' : c*(2^z' - 1) with c = alpha*log2(e), h' = log2(e)*h feeds the next layer whose weights
//     carry the inverse factor (folded on the host);
//   * fp16 path: the ELU runs on packed pairs -- F2FP (pack z'), ex2.approx.f16x2 (2 MUFU + PRMT), one HFMA2
//     (c*e - c), HSET2 + LOP3 select: 7 instructions per pair; GO2P_TC_POLYMASK selects the pairs of a job that take a
//     packed-half polynomial 2^x on the FMA pipe instead (13 instructions, no MUFU) so the MUFU pipe (16 lanes/clk/SM,
//     the per-SM floor of this policy) and the FMA/ALU pipes share the load.  bf16 keeps the fp32 exponential;
//   * one control warp issues every bulk copy and every tcgen05.mma (elect.sync) in schedule order, in a fixed K
//     order, so the fp32 accumulation order (every output bit) does not depend on timing.
#pragma once
#include "policy_dev.cuh"
#include "ptx_sm100.cuh"

#ifndef GO2P_TC_POLYMASK
#define GO2P_TC_POLYMASK 0x8888u   // bit j: column pair j (of a warp's 16 per job) uses the FMA-pipe exponential
#endif

namespace go2p {

constexpr int kTcTileM = 128;
constexpr int kTcHidden = 128;     // every hidden width handled by this kernel
constexpr int kTcOutPad = 16;      // last layer N padded to 16 (smallest UMMA N at M=128)
constexpr int kTcBiasK = 16;       // extra K block of the weights carrying the bias rows (hi / lo)
constexpr int kTcWorkers = 16;     // warps 0..15: one pool, 4 TMEM lane quarters (warp % 4) x 4 column blocks (warp / 4)
constexpr int kTcCtrlWarp = 16;    // warp 16: bulk-copy producer + MMA issuer; owns the TMEM allocation
constexpr int kTcThreads = (kTcWorkers + 1) * 32;
constexpr int kTcSlots = 3;        // tiles in flight per CTA
constexpr uint32_t kTcColA = 384;  // TMEM: accumulator of slot s at columns 128*s, A buffer b at 384 + 64*b
constexpr int kTcOnesBytes = 4096; // constant A tile of the bias K step: 128 rows x 16 K, 16-bit
constexpr int kTcSchedMax = 3 * (kMaxLayers + 1);

// One entry of the static job schedule (a cycle = P rounds of 3 jobs, slot = position % 3):
//   bits [0,4) stage g: 0 = conv, 1..L = epilogue of hidden layer g-1, P-1 = out
//   bit  4     dc: the job belongs to the slot's tile of the PREVIOUS cycle (the slot lags: staggered start)
//   bit  5     b: A buffer an A-writing job uses (XOR the cycle's parity when a cycle holds an odd number of them)
//   bit  6     hz: the job two positions back is an A-writing job on the same buffer whose MMA group may still be
//              reading it: wait for that slot's commit before overwriting; bits [7,9) = that job's cycles-back count
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
  MotorCmdDev* cmd;          // [B] send_command arguments in Unitree motor order (flag 4) or null
  float kp, kd, kp_deadman;
  uint16_t sched[kTcSchedMax + 1];
  int sched_odd;             // a cycle holds an odd number of A-writing jobs: buffer parity alternates between cycles
  unsigned long long* trace;   // debug timeline (GO2P_TC_TRACE): [0] = count, then (event, clock64) pairs; CTA 0 only
};

// host: the schedule for this layer count (see the entry layout above)
inline void tc_build_schedule(TcArgs& a) {
  const int P = a.n_layers + 1, n = 3 * P;
  const int d[3] = {0, P / 3, (2 * P) / 3};
  int is_a[kTcSchedMax], dcv[kTcSchedMax], acount = 0;
  for (int r = 0; r < P; ++r)
    for (int s = 0; s < 3; ++s) {
      const int k = 3 * r + s;
      const int g = (r - d[s] + P) % P;
      dcv[k] = r < d[s] ? 1 : 0;
      is_a[k] = g < P - 1;
      a.sched[k] = (uint16_t)(g | (dcv[k] << 4) | ((is_a[k] ? (acount & 1) : 0) << 5));
      if (is_a[k]) ++acount;
    }
  a.sched_odd = acount & 1;
  for (int k = 0; k < n; ++k) {
    const int j2 = (k - 2 + n) % n, j1 = (k - 1 + n) % n;
    if (is_a[k] && is_a[j1] && is_a[j2]) a.sched[k] |= (uint16_t)((1 << 6) | ((dcv[j2] + (k < 2 ? 1 : 0)) << 7));
  }
  a.sched[n] = 0;
}

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


// K extent of layer l's weights (including the bias block) and its N
__host__ __device__ inline int tc_layer_kp(const TcArgs& a, int l) { return l == 0 ? a.k0p : kTcHidden + kTcBiasK; }
__host__ __device__ inline int tc_layer_n(const TcArgs& a, int l) { return l == a.n_layers - 1 ? kTcOutPad : kTcHidden; }
__host__ __device__ inline size_t tc_weight_bytes(const TcArgs& a) {
  size_t s = 0;
  for (int l = 0; l < a.n_layers; ++l) s += (size_t)tc_layer_kp(a, l) * tc_layer_n(a, l) * 2;
  return s;
}
__host__ __device__ inline size_t tc_stage_bytes(const TcArgs& a) { return ((size_t)kTcTileM * a.in_dim * 4 + 127) & ~(size_t)127; }
__host__ __device__ inline size_t tc_smem_bytes(const TcArgs& a) { return tc_weight_bytes(a) + 2 * tc_stage_bytes(a) + kTcOnesBytes + 256; }

// c*(2^z - 1) for z < 0 on a packed fp16 pair without the MUFU: clamp at -13 (2^-13 is below the resolution of the
// result), split z = -k + r with the 1536 = 1.5*2^10 rounding trick (k = 0..13 lands in the low mantissa bits of u),
// degree-3 polynomial for 2*2^r on [-0.5, 0.5] (the factor 2 keeps every pair's exponent field >= k, so the packed
// integer subtraction of k << 10 never borrows across the halves, whatever garbage a discarded z >= 0 lane holds),
// then (c/2)*p - c as one HFMA2: 10 instructions after the pack, none on the MUFU.  Max abs error 9.1e-4 over all
// negative fp16 inputs vs 8.4e-4 for a correctly rounded 2^z followed by the same HFMA2 (exhaustive CPU emulation,
// scripts/experiments/polyelu.py).
__device__ __forceinline__ uint32_t elu_neg_poly_f16x2(uint32_t z, uint32_t ch2, uint32_t nc2) {
  const uint32_t kM = 0x66006600u;                        // (1536, 1536)
  const uint32_t zc = ptx::max_f16x2(z, 0xCA80CA80u);     // max(z, -13)
  const uint32_t u = ptx::sub_f16x2(kM, zc);              // 1536 + k, k = round(-zc)
  const uint32_t r = ptx::add_f16x2(zc, ptx::sub_f16x2(u, kM));
  uint32_t p = ptx::fma_f16x2(0x2F102F10u, r, 0x37C337C3u);   // 0.11035 r + 0.48511
  p = ptx::fma_f16x2(p, r, 0x3D8C3D8Cu);                      // ... + 1.38672
  p = ptx::fma_f16x2(p, r, 0x40004000u);                      // ... + 2.0
  p -= (u & 0x000F000Fu) << 10;                               // * 2^-k
  return ptx::fma_f16x2(p, ch2, nc2);
}

// ELU in the base-2 domain on a warp's 32 accumulator columns -> 16 words of packed 16-bit operands.
//   e = 2^z', f = c*e - c, result = z' < 0 ? f : z' selected on the packed pair.
template <bool kFp16>
__device__ __forceinline__ void elu_pack32(const uint32_t (&v)[32], bool has_elu, float c, uint32_t (&p)[16]) {
  if (has_elu) {
    const float nc = -c;
    if constexpr (kFp16) {
      const uint32_t c2 = ptx::pack_f16_sat(c, c), nc2 = ptx::pack_f16_sat(nc, nc), ch2 = ptx::pack_f16_sat(0.5f * c, 0.5f * c);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t zp = ptx::pack_f16_sat(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        const uint32_t f = ((GO2P_TC_POLYMASK >> j) & 1u) ? elu_neg_poly_f16x2(zp, ch2, nc2)
                                                         : ptx::fma_f16x2(ptx::ex2_f16x2(zp), c2, nc2);
        p[j] = ptx::select_neg_f16x2(zp, f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float z0 = __uint_as_float(v[2 * j]), z1 = __uint_as_float(v[2 * j + 1]);
        const float f0 = fmaf(ptx::ex2_approx(z0), c, nc);
        const float f1 = fmaf(ptx::ex2_approx(z1), c, nc);
        p[j] = ptx::select_neg_bf16x2(ptx::pack_bf16(z0, z1), ptx::pack_bf16(f0, f1));
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

// ---- cold paths, kept out of line so the steady-state loop of the kernel stays small in the instruction cache

// layer-0 A operand of one row for a ragged last tile (read from global memory) or an odd input width
template <bool kFp16>
__device__ __noinline__ void tc_conv_slow(const TcArgs& a, const float* rowp, bool live, int c8_lo, int c8_hi, uint32_t a0_t) {
  for (int c8 = c8_lo; c8 < c8_hi; ++c8) {
    uint32_t q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c8 * 16 + 2 * j;
      float lo = 0.f, hi = 0.f;
      if (k < a.in_dim) { if (live) lo = rowp[k]; } else if (k <= a.in_dim + 1) lo = 1.f;
      if (k + 1 < a.in_dim) { if (live) hi = rowp[k + 1]; } else if (k + 1 <= a.in_dim + 1) hi = 1.f;
      q[j] = kFp16 ? ptx::pack_f16_sat(lo, hi) : ptx::pack_bf16(lo, hi);
    }
    ptx::tmem_st_x8(a0_t + (uint32_t)(8 * c8), q);
  }
}

// output epilogue for anything but the plain 12-output, no-activation case: one warp per lane quarter (column block 0)
__device__ __noinline__ void tc_out_generic(const TcArgs& a, uint32_t o_t, long long row, bool live) {
  uint32_t v[16];
  ptx::tmem_ld_x16(o_t, v);
  ptx::tc_wait_ld();
  if (!live) return;
  const int L = a.n_layers - 1;
  const int b0 = (a.flags & 5u) && a.button0 ? a.button0[row] : 0;
  float* dst = a.act + row * a.out_dim;
#pragma unroll 1
  for (int j = 0; j < a.out_dim; ++j) {
    float x = __uint_as_float(v[0]);
#pragma unroll
    for (int q = 1; q < 16; ++q) if (q == j) x = __uint_as_float(v[q]);
    if (a.has_elu[L]) x = ((x < 0.f) ? fmaf(ptx::ex2_approx(x), a.elu_c[L], -a.elu_c[L]) : x) * a.out_scale;
    if (a.flags & 1u) x = clamp_mask(x, a.action_limit, b0);
    dst[j] = x;
    if (j < kDof && (a.flags & 6u)) {
      const double qd = joint_target(x, a.q0[j], a.action_scale);
      if ((a.flags & 2u) && a.qdes) a.qdes[row * kDof + j] = qd;
      if ((a.flags & 4u) && a.cmd) a.cmd[row].q_des[motor_of_isaac(j)] = qd;
    }
  }
  if ((a.flags & 4u) && a.cmd) store_gains(a.cmd, row, b0, a.kp, a.kd, a.kp_deadman);
}

// Walks the static schedule: position k inside cycle c, slot = k % 3.  Both roles decode the same entries.
struct TcSched {
  int c, k, s;          // cycle, position in the cycle, slot
  __device__ __forceinline__ void init() { c = 0; k = 0; s = 0; }
  __device__ __forceinline__ void next(int n) {
    ++k;
    s = (s == kTcSlots - 1) ? 0 : s + 1;
    if (k == n) { k = 0; ++c; }       // n is a multiple of 3: s wraps together with k
  }
  // CTA-local tile of an entry of this position's slot (negative: none yet)
  __device__ __forceinline__ int tile(uint32_t e) const { return kTcSlots * (c - (int)((e >> 4) & 1u)) + s; }
};

template <bool kFp16>
__global__ void __launch_bounds__(kTcThreads, 1) tc_mlp_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t wbytes = (uint32_t)tc_weight_bytes(a);
  const uint32_t stage_bytes = (uint32_t)tc_stage_bytes(a);
  uint8_t* w_smem = smem;
  uint8_t* stage0 = smem + wbytes;
  uint32_t* ones = reinterpret_cast<uint32_t*>(stage0 + 2 * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ones) + kTcOnesBytes);
  // barrier words: [0,2) obs_full (observation slab of stage i landed) | [2,5) acc_full (slot's MMA group complete,
  // tcgen05.commit) | [5,8) a_ready (A operand of the slot's next layer written: one arrival per worker warp) |
  // [8,8+kMaxLayers) w_full (layer weights landed, completes once)
  uint64_t* w_full = bars + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8 + kMaxLayers);
  const uint32_t bar_obs = ptx::smem_u32(bars), bar_acc = bar_obs + 16u, bar_rdy = bar_obs + 40u;

  // ---- one-time setup: barriers, TMEM, the constant ones tile; the weights arrive as one bulk async copy per layer,
  // each with its own mbarrier
  if (warp == kTcCtrlWarp) {
    if (lane == 0) {
      for (int s = 0; s < 5; ++s) ptx::mbar_init(&bars[s], 1);
      for (int s = 0; s < kTcSlots; ++s) ptx::mbar_init(&bars[5 + s], kTcWorkers);
      for (int l = 0; l < a.n_layers; ++l) ptx::mbar_init(&w_full[l], 1);
      ptx::fence_mbar_init();
      uint32_t off = 0;
      for (int l = 0; l < a.n_layers; ++l) {
        const uint32_t bytes = (uint32_t)(tc_layer_kp(a, l) * tc_layer_n(a, l) * 2);
        ptx::mbar_arrive_expect_tx(&w_full[l], bytes);
        ptx::bulk_g2s(w_smem + off, reinterpret_cast<const uint8_t*>(a.wpack) + off, bytes, &w_full[l]);
        off += bytes;
      }
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  // ones tile, K-major core matrices like the weights (K-adjacent cores 128 B apart, 8-row groups 256 B apart):
  // word (r/8)*64 + (r%8)*4 of row r holds K = 0,1 -> (1.0, 1.0); everything else is zero
  {
    const uint32_t one2 = kFp16 ? 0x3C003C00u : 0x3F803F80u;
    for (int w = tid; w < kTcOnesBytes / 4; w += kTcThreads) ones[w] = ((w & 63) < 32 && (w & 3) == 0) ? one2 : 0u;
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  block_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Programmatic dependent launch: everything above (barriers, TMEM, the weights' bulk copies -- the weights never
  // change after go2p_create) may run while the previous kernel of the stream drains; observations are read and
  // actions written only after that kernel has completed.  The next launch may start its own prologue as soon as
  // this CTA's SM is free (a short batch per GPU, e.g. 1/8 of a sharded step, is otherwise launch-gap bound).
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();

  const long long n_tiles = (a.B + kTcTileM - 1) / kTcTileM;
  const int n_local = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles b, b+G, ...  (>= 1)
  const int L = a.n_layers - 1;            // hidden layers = epilogue stages; also the index of the output layer
  const int P = a.n_layers + 1;            // stages per tile
  const int n_sched = kTcSlots * P;
  const int n_cycles = (n_local + kTcSlots - 1) / kTcSlots + 1;   // one more for the jobs of the lagging slots
  TC_TRACE_INIT();
  TcSched it;
  it.init();

  if (warp == kTcCtrlWarp) {
    // ================= control warp: bulk-copy producer + MMA issuer, in schedule order =================
    const uint32_t fmt = kFp16 ? ptx::FMT_F16 : ptx::FMT_BF16;
    const uint32_t w_base = ptx::smem_u32(w_smem);
    const uint32_t tile_bytes = (uint32_t)(kTcTileM * a.in_dim * 4);
    const uint64_t ones_desc = ptx::make_smem_desc_nosw(ptx::smem_u32(ones), 128u, 256u);
    const uint32_t idesc_h = ptx::make_idesc(fmt, kTcTileM, kTcHidden), idesc_o = ptx::make_idesc(fmt, kTcTileM, kTcOutPad);
    // K step of 16 elements = two 8x16B core matrices along K (128 B apart); 8-row groups are kp*16 B apart;
    // consecutive K steps are 256 B apart (+16 in the descriptor's address field)
    const uint64_t bdesc_l0 = ptx::make_smem_desc_nosw(w_base, 128u, (uint32_t)a.k0p * 16u);
    const uint64_t bdesc_l1 = ptx::make_smem_desc_nosw(w_base + (uint32_t)(a.k0p * kTcHidden * 2), 128u, (uint32_t)(kTcHidden + kTcBiasK) * 16u);
    const uint32_t w_step16 = (uint32_t)((kTcHidden + kTcBiasK) * kTcHidden * 2) >> 4;   // hidden layer stride in descriptor units
    const int ksteps0 = a.k0p / 16;
    auto load_tile = [&](int i) {            // CTA-local tile i -> stage i & 1
      const long long row0 = (long long)(blockIdx.x + i * gridDim.x) * kTcTileM;
      if (a.B - row0 >= kTcTileM) {
        TC_TRACE(0x100u | (uint32_t)(i & 1));
        ptx::mbar_arrive_expect_tx(&bars[i & 1], tile_bytes);
        ptx::bulk_g2s(stage0 + (i & 1) * stage_bytes, a.obs + row0 * a.in_dim, tile_bytes, &bars[i & 1]);
      } else {
        ptx::mbar_arrive(&bars[i & 1]);      // ragged last tile: the pool reads global memory directly
      }
    };
    if (ptx::elect_one_sync()) {
      load_tile(0);
      if (n_local > 1) load_tile(1);
    }
    __syncwarp();
    for (int l = 0; l < a.n_layers; ++l) ptx::mbar_wait(&w_full[l], 0u);   // (the first conversion takes longer than this)
    uint32_t par_ready = 0u;                 // bit s: parity of the next wait on a_ready[s]
    while (it.c < n_cycles) {
      const uint32_t e = a.sched[it.k];
      const int g = (int)(e & 15u), s = it.s, i = it.tile(e);
      if (g < P - 1 && i >= 0 && i < n_local) {
        // conv feeds layer 0, the epilogue of hidden layer g-1 feeds layer g
        ptx::mbar_wait_u32(bar_rdy + 8u * (uint32_t)s, (par_ready >> s) & 1u);
        par_ready ^= 1u << s;
        ptx::tc_fence_after();
        if (ptx::elect_one_sync()) {
          TC_TRACE(0x200u | (uint32_t)(g << 4) | (uint32_t)s);
          const uint32_t b = ((e >> 5) ^ (uint32_t)(a.sched_odd & it.c)) & 1u;
          const uint32_t d_t = tmem_base + 128u * (uint32_t)s;
          const uint32_t a_t = tmem_base + kTcColA + 64u * b;
          if (g == 0) {
            for (int j = 0; j < ksteps0; ++j)        // ones inside the data
              ptx::mma_f16_ts(d_t, a_t + (uint32_t)(8 * j), bdesc_l0 + (uint64_t)(j * 16), idesc_h, j > 0 ? 1u : 0u);
          } else {
            const uint64_t bdesc = bdesc_l1 + (uint64_t)((uint32_t)(g - 1) * w_step16);
            const uint32_t idesc = (g == L) ? idesc_o : idesc_h;
            ptx::mma_f16_ss(d_t, ones_desc, bdesc + (uint64_t)(8 * 16), idesc, 0u);   // bias K step (SS form)
#pragma unroll
            for (int j = 0; j < kTcHidden / 16; ++j)
              ptx::mma_f16_ts(d_t, a_t + (uint32_t)(8 * j), bdesc + (uint64_t)(j * 16), idesc, 1u);
          }
          ptx::mma_commit(&bars[2 + s]);
          TC_TRACE(0x300u | (uint32_t)(g << 4) | (uint32_t)s);
          // the conversion that just completed has released its observation stage: refill it
          if (g == 0 && i + 2 < n_local) load_tile(i + 2);
        }
        __syncwarp();
      }
      it.next(n_sched);
    }
  } else {
    // ================= worker warps: one pool of 16 warps walks the schedule =================
    // pool = 4 TMEM lane quarters x 4 column blocks (warp = cb*4 + quarter): every warp owns a 32-column block of
    // its 32 rows
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch (== warp % 4)
    const int cb = warp >> 2;                // column block
    const int m = quarter * 32 + lane;       // row inside the tile
    const uint32_t t_lane = tmem_base + (((uint32_t)(quarter * 32)) << 16);
    const bool out12 = a.out_dim == 12 && !a.has_elu[L];

    uint32_t par_acc = 0u;                   // bit s: parity of the next wait on acc_full[s]
    uint32_t R[32];                          // accumulator columns of the job being / about to be processed
    int b0v = 0;                             // out job: the row's dead-man button, loaded a job ahead

    // ---- a job's inputs: wait for them and request them (accumulator columns -> R; button0)
    auto acquire = [&](int s, int g, int i) {
      TC_TRACE(0xD00u | (uint32_t)s);
      if (g == 0) {
        ptx::mbar_wait_u32(bar_obs + 8u * (uint32_t)(i & 1), (uint32_t)((i >> 1) & 1));
        TC_TRACE(0x400u | (uint32_t)s);
        return;
      }
      ptx::mbar_wait_u32(bar_acc + 8u * (uint32_t)s, (par_acc >> s) & 1u);
      par_acc ^= 1u << s;
      ptx::tc_fence_after();
      const uint32_t d_t = t_lane + 128u * (uint32_t)s;
      if (g < P - 1) {
        TC_TRACE(0x600u | (uint32_t)((g - 1) << 4) | (uint32_t)s);
        ptx::tmem_ld_x32(d_t + (uint32_t)(cb * 32), R);
      } else if (out12 && cb < 3) {
        TC_TRACE(0x800u | (uint32_t)s);
        ptx::tmem_ld_x4(d_t + (uint32_t)(cb * 4), reinterpret_cast<uint32_t(&)[4]>(R[0]));   // outputs 4cb .. 4cb+3
        if ((a.flags & 5u) && a.button0) {
          const long long row = (long long)(blockIdx.x + i * gridDim.x) * kTcTileM + m;
          b0v = row < a.B ? __ldg(a.button0 + row) : 0;
        }
      }
    };

    // software pipeline over the schedule: (e1, i1) is the entry after the current one, decoded one job ahead
    uint32_t e0 = a.sched[0];
    int i0 = it.tile(e0), s0 = 0, c0 = 0;
    bool pre = false;                        // the current job's inputs were already requested
    while (c0 < n_cycles) {
      it.next(n_sched);
      const uint32_t e1 = a.sched[it.k];
      const int i1 = it.tile(e1), s1 = it.s;
      const bool v1 = it.c < n_cycles && i1 >= 0 && i1 < n_local;
      if (i0 >= 0 && i0 < n_local) {
        const int s = s0, g = (int)(e0 & 15u), i = i0;
        const uint32_t b = ((e0 >> 5) ^ (uint32_t)(a.sched_odd & c0)) & 1u;
#ifndef GO2P_TC_NO_RESYNC
        // keep the pool in step: with three slots the MMA barriers are usually complete when a warp gets to them, so
        // nothing else re-aligns the warps; left alone they drift up to two jobs apart, every job then completes (and
        // its MMA group issues) only when the slowest warp gets there, and four warps per scheduler run four
        // different pieces of code
        asm volatile("bar.sync 1, 512;" ::: "memory");
#endif
        if (!pre) acquire(s, g, i);
        // the MMA group that read this A buffer two jobs ago (other slot) must be complete before it is overwritten
        auto a_buffer_free = [&]() {
          if (e0 & 64u) {
            const int sp = (s == kTcSlots - 1) ? 0 : s + 1;                       // slot of the job two positions back
            const int ip = kTcSlots * (c0 - (int)((e0 >> 7) & 3u)) + sp;          // its tile: did that job exist?
            if (ip >= 0 && ip < n_local) {
              ptx::mbar_wait_u32(bar_acc + 8u * (uint32_t)sp, (par_acc >> sp) & 1u);
              ptx::tc_fence_after();
            }
          }
        };
        if (g == 0) {
          // ---- conv: fp32 observation rows -> 16-bit layer-0 A operand, constant ones at K = in_dim, in_dim+1.
          //      chunk c8 (16 K elements) lands on columns 8*c8 .. 8*c8+7 of the A buffer; block cb = chunks 2cb, 2cb+1
          const long long row0 = (long long)(blockIdx.x + i * gridDim.x) * kTcTileM;
          const bool full = a.B - row0 >= kTcTileM;
          const float* srow = reinterpret_cast<const float*>(stage0 + (i & 1) * stage_bytes) + m * a.in_dim;
          const uint32_t a_t = t_lane + kTcColA + 64u * b;
          const int c8_hi = min(a.k0p / 16, 2 * cb + 2);
          a_buffer_free();
          if (full && (a.in_dim & 1) == 0) {
            // full tile in shared memory, rows 8-byte aligned: vector loads, uniform per-pair classification
            const float2* r2 = reinterpret_cast<const float2*>(srow);
#pragma unroll 1
            for (int c8 = 2 * cb; c8 < c8_hi; ++c8) {
              uint32_t q[8];
              if (c8 * 16 + 16 <= a.in_dim) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float2 t = r2[c8 * 8 + j]; q[j] = kFp16 ? ptx::pack_f16_sat(t.x, t.y) : ptx::pack_bf16(t.x, t.y); }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int k = c8 * 16 + 2 * j;
                  float2 t = make_float2(0.f, 0.f);
                  if (k < a.in_dim) t = r2[k >> 1]; else if (k == a.in_dim) t = make_float2(1.f, 1.f);
                  q[j] = kFp16 ? ptx::pack_f16_sat(t.x, t.y) : ptx::pack_bf16(t.x, t.y);
                }
              }
              ptx::tmem_st_x8(a_t + (uint32_t)(8 * c8), q);
            }
          } else {
            const long long left = a.B - row0;
            tc_conv_slow<kFp16>(a, full ? srow : a.obs + (row0 + m) * a.in_dim, m < left, 2 * cb, c8_hi, a_t);
          }
        } else if (g < P - 1) {
          // ---- E(l): accumulator block -> ELU -> 16-bit A operand of the next layer
          uint32_t pk[16];
          ptx::tc_wait_ld();
          TC_TRACE(0xF00u | (1u << 4) | (uint32_t)s);
          elu_pack32<kFp16>(R, a.has_elu[g - 1] != 0, a.elu_c[g - 1], pk);
          a_buffer_free();
          TC_TRACE(0xA00u | (uint32_t)s);
          ptx::tmem_st_x16(t_lane + kTcColA + 64u * b + (uint32_t)(cb * 16), pk);
          TC_TRACE(0xE00u | (uint32_t)s);
        } else {
          // ---- out (bias already inside the accumulator): (+ELU) (+clamp/mask) (+q_des, motor commands) -> global
          const long long row = (long long)(blockIdx.x + i * gridDim.x) * kTcTileM + m;
          if (out12) {
            // the policy's case: 12 outputs, no activation -- column block cb < 3 stores one float4 of every row
            if (cb < 3) {
              ptx::tc_wait_ld();
              if (row < a.B) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = __uint_as_float(R[j]);
                if (a.flags & 1u) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) o[j] = clamp_mask(o[j], a.action_limit, b0v);
                }
                reinterpret_cast<float4*>(a.act + row * 12)[cb] = make_float4(o[0], o[1], o[2], o[3]);
                if (a.flags & 6u) {
                  double qd[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) qd[j] = joint_target(o[j], a.q0[cb * 4 + j], a.action_scale);
                  if ((a.flags & 2u) && a.qdes) {
                    double2* q2 = reinterpret_cast<double2*>(a.qdes + row * kDof + cb * 4);
                    q2[0] = make_double2(qd[0], qd[1]);
                    q2[1] = make_double2(qd[2], qd[3]);
                  }
                  if ((a.flags & 4u) && a.cmd) {
                    // Isaac joint 4*cb + j = joint type cb of leg j -> Unitree motor (j ^ 1)*3 + cb
                    MotorCmdDev* c = a.cmd + row;
#pragma unroll
                    for (int j = 0; j < 4; ++j) c->q_des[(j ^ 1) * 3 + cb] = qd[j];
                    if (cb == 0) store_gains(a.cmd, row, b0v, a.kp, a.kd, a.kp_deadman);
                  }
                }
              }
            }
          } else if (cb == 0) {
            tc_out_generic(a, t_lane + 128u * (uint32_t)s, row, row < a.B);
          }
          // these accumulator loads are ordered before this warp's next arrive (the conversion of the slot's next
          // tile); the layer-0 MMA that overwrites the accumulator is issued after all such arrivals
          ptx::tc_fence_before();
          TC_TRACE(0x900u | (uint32_t)s);
        }
        // ---- the next job's inputs are requested before this job's stores are drained and signalled (software
        //      pipeline); the next position is always another slot, whose MMA group was issued two jobs ago
        pre = v1;
        if (pre) acquire(s1, (int)(e1 & 15u), i1);
        if (g < P - 1) {
          // publish the A operand: stores complete -> ordered before the arrive -> the control warp issues the MMAs
          ptx::tc_wait_st();
          TC_TRACE(0xB00u | (uint32_t)s);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_u32(bar_rdy + 8u * (uint32_t)s);
          TC_TRACE(0x700u | (uint32_t)s);
        }
      } else {
        pre = false;
      }
      e0 = e1; i0 = i1; s0 = s1; c0 = it.c;
    }
  }

  // ---- teardown
  ptx::tc_fence_before();
  block_sync();
  if (warp == kTcCtrlWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace go2p
