// Batched policy forward for the Go2 topology (4 layers, hidden width 128): THREE 128-row tiles in flight per CTA with a
// STATIC job schedule.  Same arithmetic and operand layouts as tc_mlp_kernel (kernels_tc.cuh: weights resident in shared
// memory as UMMA core matrices, base-2 exponent domain, packed-half ELU, fused clamp/mask/q_des epilogue); what differs
// is how TMEM is used and who waits for what:
//   * TMEM (512 columns) = one fp32 accumulator per slot (3 x 128 columns) + two 64-column A-operand buffers shared by
//     all slots.  An epilogue job reads its slot's accumulator, applies the ELU and writes the next layer's A operand
//     into the A buffer whose turn it is; that layer's MMA group accumulates back into the slot's own accumulator.  An
//     accumulator is live only between the MMA and the epilogue's load, an A buffer only between the epilogue's store
//     and the MMA -- so three tiles fit where the in-place ping-pong of tc_mlp_kernel allows two;
//   * the bias cannot ride in two spare A columns any more (64 columns = K 128 exactly): hidden layers get one SS-form
//     K step whose A operand is a constant 128 x 16 tile (two columns of ones) in shared memory and whose B rows hold
//     hi / lo halves of the bias (measured: a chain of 1 SS + 8 TS steps costs the same as 9 TS steps);
//   * the pool walks a FIXED cycle of 15 jobs over three tiles -- slot, layer, A buffer, barrier parities are
//     compile-time constants of every job, as in the two-slot kernel (three measured variants with run-time schedules
//     paid more in per-job bookkeeping than the hand-off they hid: scripts/experiments/README.md) --
//        A: E0(s0) E0(s1) out(s2) conv(s2)   B: E1(s0) E1(s1) E0(s2)   C: E2(s0) E2(s1) E1(s2)
//        D: out(s0) conv'(s0) out(s1) conv'(s1) E2(s2)
//     slot 2 runs one stage behind so that the three conversions (two observation stages) are spread out.  Every MMA
//     group has two whole jobs of cover before the pool returns to its slot: the ~1,400-cycle hand-off that bounds
//     tc_mlp_kernel (period max(2 D_E, D_E + T_h)) leaves the critical path (period 3 D_E per three tiles).
#pragma once
#include "kernels_tc.cuh"

namespace go2p {

constexpr int kT3Workers = 16;
constexpr int kT3CtrlWarp = 16;
constexpr int kT3Threads = (kT3Workers + 1) * 32;
constexpr uint32_t kT3ColA = 384;          // accumulator of slot s at columns 128*s, A buffer b at 384 + 64*b
constexpr int kT3OnesBytes = 4096;         // constant A tile of the bias K step: 128 rows x 16 K, 16-bit
constexpr int kT3Layers = 4;

__host__ __device__ inline size_t tc3_smem_bytes(const TcArgs& a) { return tc_weight_bytes(a) + 2 * tc_stage_bytes(a) + kT3OnesBytes + 256; }
// shapes this kernel serves: the Go2 topology (any input width the narrow family allows), 12 outputs, no activation on
// the output layer, even input width (vector loads in the conversion)
__host__ inline bool tc3_serves(const TcArgs& a) {
  return a.n_layers == kT3Layers && a.out_dim == 12 && !a.has_elu[kT3Layers - 1] && (a.in_dim & 1) == 0 && a.raw == nullptr &&
         a.sat_count == nullptr;
}

template <bool kFp16>
__global__ void __launch_bounds__(kT3Threads, 1) tc3_mlp_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t wbytes = (uint32_t)tc_weight_bytes(a);
  const uint32_t stage_bytes = (uint32_t)tc_stage_bytes(a);
  uint8_t* w_smem = smem;
  uint8_t* stage0 = smem + wbytes;
  uint32_t* ones = reinterpret_cast<uint32_t*>(stage0 + 2 * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ones) + kT3OnesBytes);
  uint64_t* obs_full = bars;        // [2]  observation slab of stage i landed
  uint64_t* acc_full = bars + 2;    // [3]  slot's MMA group complete (tcgen05.commit); 4 phases per tile: parity 0,1,0,1
  uint64_t* a_ready = bars + 5;     // [3]  A operand of the slot's next layer written (16 arrivals); 4 phases per tile
  uint64_t* w_full = bars + 8;      // [4]  layer weights landed (completes once)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8 + kT3Layers);

  if (warp == kT3CtrlWarp) {
    if (lane == 0) {
      for (int s = 0; s < 5; ++s) ptx::mbar_init(&bars[s], 1);
      for (int s = 0; s < 3; ++s) ptx::mbar_init(&a_ready[s], kT3Workers);
      for (int l = 0; l < kT3Layers; ++l) ptx::mbar_init(&w_full[l], 1);
      ptx::fence_mbar_init();
      uint32_t off = 0;
      for (int l = 0; l < kT3Layers; ++l) {
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
    for (int w = tid; w < kT3OnesBytes / 4; w += kT3Threads) ones[w] = ((w & 63) < 32 && (w & 3) == 0) ? one2 : 0u;
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  block_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::grid_dependency_wait();          // programmatic dependent launch: see tc_mlp_kernel
  ptx::grid_launch_dependents();

  const long long n_tiles = (a.B + kTcTileM - 1) / kTcTileM;
  const int n_local = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles b, b+G, ...  (>= 1)
  TC_TRACE_INIT();

  if (warp == kT3CtrlWarp) {
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
        ptx::mbar_arrive_expect_tx(&obs_full[i & 1], tile_bytes);
        ptx::bulk_g2s(stage0 + (i & 1) * stage_bytes, a.obs + row0 * a.in_dim, tile_bytes, &obs_full[i & 1]);
      } else {
        ptx::mbar_arrive(&obs_full[i & 1]);  // ragged last tile: the pool reads global memory directly
      }
    };
    // the MMA group of layer l for slot s from A buffer b, once the job that wrote the operand has signalled
    // (a_ready phase parity par); a conversion job also releases its observation stage
    auto issue = [&](int s, int l, uint32_t b, uint32_t par, int tile) {
      if (tile >= n_local) return;
      ptx::mbar_wait(&a_ready[s], par);
      ptx::tc_fence_after();
      if (ptx::elect_one_sync()) {
        TC_TRACE(0x200u | (uint32_t)(l << 4) | (uint32_t)s);
        const uint32_t d_t = tmem_base + 128u * (uint32_t)s;
        const uint32_t a_t = tmem_base + kT3ColA + 64u * b;
        if (l == 0) {
          for (int j = 0; j < ksteps0; ++j)          // ones inside the data
            ptx::mma_f16_ts(d_t, a_t + (uint32_t)(8 * j), bdesc_l0 + (uint64_t)(j * 16), idesc_h, j > 0 ? 1u : 0u);
        } else {
          const uint64_t bdesc = bdesc_l1 + (uint64_t)((uint32_t)(l - 1) * w_step16);
          const uint32_t idesc = (l == kT3Layers - 1) ? idesc_o : idesc_h;
          ptx::mma_f16_ss(d_t, ones_desc, bdesc + (uint64_t)(8 * 16), idesc, 0u);   // bias K step (SS form)
#pragma unroll
          for (int j = 0; j < kTcHidden / 16; ++j)
            ptx::mma_f16_ts(d_t, a_t + (uint32_t)(8 * j), bdesc + (uint64_t)(j * 16), idesc, 1u);
        }
        ptx::mma_commit(&acc_full[s]);
        TC_TRACE(0x300u | (uint32_t)(l << 4) | (uint32_t)s);
        if (l == 0 && tile + 2 < n_local) load_tile(tile + 2);   // the conversion released its observation stage
      }
      __syncwarp();
    };
    if (ptx::elect_one_sync()) {
      load_tile(0);
      if (n_local > 1) load_tile(1);
    }
    __syncwarp();
    for (int l = 0; l < kT3Layers; ++l) ptx::mbar_wait(&w_full[l], 0u);
    issue(0, 0, 1u, 0u, 0);                  // prologue: conv(s0), conv(s1)
    issue(1, 0, 0u, 0u, 1);
    for (int c = 0; 3 * c < n_local; ++c) {
      const int t0 = 3 * c, t1 = t0 + 1, t2 = t0 + 2;
      issue(0, 1, 0u, 1u, t0); issue(1, 1, 1u, 1u, t1); issue(2, 0, 0u, 0u, t2);             // A: E0 E0 conv
      issue(0, 2, 1u, 0u, t0); issue(1, 2, 0u, 0u, t1); issue(2, 1, 1u, 1u, t2);             // B: E1 E1 E0
      issue(0, 3, 0u, 1u, t0); issue(1, 3, 1u, 1u, t1); issue(2, 2, 0u, 0u, t2);             // C: E2 E2 E1
      issue(0, 0, 1u, 0u, t0 + 3); issue(1, 0, 0u, 0u, t1 + 3); issue(2, 3, 1u, 1u, t2);     // D: conv' conv' E2
    }
  } else {
    // ================= worker warps: one pool of 16 warps walks the static schedule =================
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch (== warp % 4)
    const int cb = warp >> 2;                // 32-column block
    const int m = quarter * 32 + lane;       // row inside the tile
    const uint32_t t_lane = tmem_base + (((uint32_t)(quarter * 32)) << 16);
    const int n8 = a.k0p / 16;
    const bool masked = (a.flags & 5u) && a.button0 != nullptr;

    uint32_t cur[16], nxt[16];               // accumulator columns of the epilogue job in flight (loaded one job ahead)
    uint32_t ov[4];                          // output job: this thread's four outputs
    int b0 = 0;                              // output job: the row's dead-man button

    // the MMA group that read A buffer b two A-jobs ago (slot hs, commit parity hp) must be complete before the buffer
    // is overwritten; `exists`: that job ran at all (start-up / tail of the tile list)
    auto a_buffer_free = [&](int hs, uint32_t hp, bool exists) {
      if (exists) {
        ptx::mbar_wait(&acc_full[hs], hp);
        ptx::tc_fence_after();
      }
    };
    // publish an A operand: stores complete -> ordered before the arrive -> the control warp issues the MMA group
    auto publish = [&](int s, int tile) {
      if (tile >= n_local) return;
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&a_ready[s]);
      TC_TRACE(0x700u | (uint32_t)s);
    };
    // Every job is split into "acquire" (wait for its inputs and request them) and "body"; the acquire of job k+1 sits
    // between the body of job k and the publication of job k's stores, so the accumulator load of the next job, its
    // barrier wait and the button load overlap the drain of the current job's tcgen05.st (software pipeline).
    auto acq_conv = [&](int tile) {
      if (tile >= n_local) return;
      ptx::mbar_wait(&obs_full[tile & 1], (uint32_t)((tile >> 1) & 1));
    };
    auto acq_e = [&](int s, uint32_t ap, int tile) {
      if (tile >= n_local) return;
      ptx::mbar_wait(&acc_full[s], ap);
      ptx::tc_fence_after();
      TC_TRACE(0x600u | (uint32_t)s);
      const uint32_t d_t = t_lane + 128u * (uint32_t)s + (uint32_t)(cb * 32);
      ptx::tmem_ld_x16(d_t, cur);
      ptx::tmem_ld_x16(d_t + 16u, nxt);
    };
    auto acq_out = [&](int s, int tile) {
      if (tile < 0 || tile >= n_local) return;
      const long long row = (long long)(blockIdx.x + tile * gridDim.x) * kTcTileM + m;
      b0 = (masked && cb < 3 && row < a.B) ? __ldg(a.button0 + row) : 0;
      ptx::mbar_wait(&acc_full[s], 1u);
      ptx::tc_fence_after();
      TC_TRACE(0x800u | (uint32_t)s);
      if (cb < 3) ptx::tmem_ld_x4(t_lane + 128u * (uint32_t)s + (uint32_t)(cb * 4), ov);
    };

    // ---- conv: fp32 observation rows -> 16-bit layer-0 A operand (ones at K = in_dim, in_dim+1) into A buffer b
    auto body_conv = [&](int s, uint32_t b, int tile, int hs, uint32_t hp, bool hz) {
      if (tile >= n_local) return;
      const long long row0 = (long long)(blockIdx.x + tile * gridDim.x) * kTcTileM;
      const bool full = a.B - row0 >= kTcTileM;
      const float* srow = reinterpret_cast<const float*>(stage0 + (tile & 1) * stage_bytes) + m * a.in_dim;
      const uint32_t a_t = t_lane + kT3ColA + 64u * b;
      const int c8_hi = min(n8, 2 * cb + 2);
      a_buffer_free(hs, hp, hz);
      if (full) {
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
        // ragged last tile: rows straight from global memory, chunk layout of the 3-slot A buffer (8 columns per chunk)
        const long long left = a.B - row0;
        const float* rowp = a.obs + (row0 + m) * a.in_dim;
        for (int c8 = 2 * cb; c8 < c8_hi; ++c8) {
          uint32_t q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = c8 * 16 + 2 * j;
            float lo = 0.f, hi = 0.f;
            if (k < a.in_dim) { if (m < left) lo = rowp[k]; } else if (k <= a.in_dim + 1) lo = 1.f;
            if (k + 1 < a.in_dim) { if (m < left) hi = rowp[k + 1]; } else if (k + 1 <= a.in_dim + 1) hi = 1.f;
            q[j] = kFp16 ? ptx::pack_f16_sat(lo, hi) : ptx::pack_bf16(lo, hi);
          }
          ptx::tmem_st_x8(a_t + (uint32_t)(8 * c8), q);
        }
      }
      TC_TRACE(0x500u | (uint32_t)s);
    };

    // ---- E(l): accumulator block (already requested: cur / nxt) -> ELU -> 16-bit A operand of layer l+1 into A buffer
    //      b.  Two halves of 16 columns: the first half's store is in flight while the second is evaluated.
    auto body_e = [&](int l, uint32_t b, int tile, int hs, uint32_t hp, bool hz) {
      if (tile >= n_local) return;
      const bool he = a.has_elu[l] != 0;
      const float c = a.elu_c[l];
      const uint32_t a_t = t_lane + kT3ColA + 64u * b + (uint32_t)(cb * 16);
      uint32_t pk[8];
      ptx::tc_wait_ld();
      elu_pack16<kFp16>(cur, he, c, pk);
      a_buffer_free(hs, hp, hz);
      ptx::tmem_st_x8(a_t, pk);
      elu_pack16<kFp16>(nxt, he, c, pk);
      ptx::tmem_st_x8(a_t + 8u, pk);
    };

    // ---- out: accumulator of the output layer -> (+clamp/mask) (+q_des, motor commands) -> global
    auto body_out = [&](int tile) {
      if (tile < 0 || tile >= n_local) return;
      const long long row = (long long)(blockIdx.x + tile * gridDim.x) * kTcTileM + m;
      if (cb < 3) {
        ptx::tc_wait_ld();
        if (row < a.B) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = __uint_as_float(ov[j]);
          if (a.flags & 1u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = clamp_mask(o[j], a.action_limit, b0);
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
              MotorCmdDev* cm = a.cmd + row;               // Isaac joint 4*cb + j -> Unitree motor (j ^ 1)*3 + cb
#pragma unroll
              for (int j = 0; j < 4; ++j) cm->q_des[(j ^ 1) * 3 + cb] = qd[j];
              if (cb == 0) store_gains(a.cmd, row, b0, a.kp, a.kd, a.kp_deadman);
            }
          }
        }
      }
      // these accumulator loads are ordered before this warp's next arrive (the conversion of the slot's next tile);
      // the layer-0 MMA group that overwrites the accumulator is issued after all such arrivals
      ptx::tc_fence_before();
      TC_TRACE(0x900u);
    };

    // A buffers alternate over the twelve operand-writing jobs of a cycle (E0 E0 conv | E1 E1 E0 | E2 E2 E1 | conv' conv'
    // E2); every job's buffer was last read by the MMA group of the job two places back, on slot (s+1)%3, which
    // completes the phase that slot's next job waits for -- waited for here first unless that job never ran.
    acq_conv(0);
    body_conv(0, 1u, 0, 0, 0u, false);  acq_conv(1);            publish(0, 0);
    body_conv(1, 0u, 1, 0, 0u, false);  acq_e(0, 0u, 0);        publish(1, 1);
    for (int c = 0; 3 * c < n_local + 3; ++c) {
      const int t0 = 3 * c, t1 = t0 + 1, t2 = t0 + 2;
      const bool x1 = t1 < n_local, x2 = t2 < n_local;
      // A: E0(s0) E0(s1) out(s2) conv(s2)
      body_e(0, 0u, t0, 1, 0u, x1);          acq_e(1, 0u, t1);       publish(0, t0);
      body_e(0, 1u, t1, 2, 1u, c > 0);       acq_out(2, t2 - 3);     publish(1, t1);
      body_out(t2 - 3);                      acq_conv(t2);
      body_conv(2, 0u, t2, 0, 1u, true);     acq_e(0, 1u, t0);       publish(2, t2);
      // B: E1(s0) E1(s1) E0(s2)
      body_e(1, 1u, t0, 1, 1u, x1);          acq_e(1, 1u, t1);       publish(0, t0);
      body_e(1, 0u, t1, 2, 0u, x2);          acq_e(2, 0u, t2);       publish(1, t1);
      body_e(0, 1u, t2, 0, 0u, true);        acq_e(0, 0u, t0);       publish(2, t2);
      // C: E2(s0) E2(s1) E1(s2)
      body_e(2, 0u, t0, 1, 0u, x1);          acq_e(1, 0u, t1);       publish(0, t0);
      body_e(2, 1u, t1, 2, 1u, x2);          acq_e(2, 1u, t2);       publish(1, t1);
      body_e(1, 0u, t2, 0, 1u, true);        acq_out(0, t0 < n_local ? t0 : -1);   publish(2, t2);
      // D: out(s0) conv'(s0) out(s1) conv'(s1) E2(s2)
      body_out(t0 < n_local ? t0 : -1);      acq_conv(t0 + 3);
      body_conv(0, 1u, t0 + 3, 1, 1u, x1);   acq_out(1, t1);         publish(0, t0 + 3);
      body_out(t1);                          acq_conv(t1 + 3);
      body_conv(1, 0u, t1 + 3, 2, 0u, x2);   acq_e(2, 0u, t2);       publish(1, t1 + 3);
      body_e(2, 1u, t2, 0, 0u, t0 + 3 < n_local);   acq_e(0, 0u, t0 + 3);   publish(2, t2);
    }
  }

  // ---- teardown
  ptx::tc_fence_before();
  block_sync();
  if (warp == kT3CtrlWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace go2p
