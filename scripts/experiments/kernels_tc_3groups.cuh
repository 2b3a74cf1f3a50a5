// Batched policy forward for the Go2 topology (4 layers, hidden width 128): THREE 128-row tiles in flight per CTA, each
// owned by its own group of four warps.  Same arithmetic and weight layout as tc_mlp_kernel (kernels_tc.cuh: weights
// resident in shared memory as UMMA core matrices, base-2 exponent domain, packed-half ELU, fused clamp/mask/q_des
// epilogue).  What differs is who works on what, and why:
//
//   * the ncu capture of tc_mlp_kernel (profiles/r02_tc_mlp_kernel_raw.csv) shows the special-function pipe at 44 % and
//     the issue slots at 68 % over the whole kernel -- but inside an epilogue job both are saturated (an ex2.f16x2 is
//     two MUFU.EX2.F16 at 8 cycles per warp each: 768 of the job's ~1,450 cycles), and in between all 16 warps wait
//     together for the same things (accumulator barrier, tcgen05.ld, tcgen05.st drain, the MMA hand-off): the pool moves
//     in lockstep, so nobody computes while anybody waits.  Static three-slot schedules of that one pool (scripts/
//     experiments/) removed the waits for the MMA but not the lockstep, and were no faster;
//   * here the pool is split into three groups of eight warps (TMEM lane quarter x column half).  A group owns one
//     slot and walks its own chain -- conv, E0, E1, E2, out, next tile -- waiting alone for its own MMA groups.  The
//     groups drift apart by construction, so while one waits (~1,400 cycles per layer for the MMA, plus the load/store
//     latencies) the other two keep the MUFU and FMA pipes busy; every scheduler holds two warps of each group
//     (measured with four warps per group: one warp per scheduler and group cannot hide its own dependent-issue
//     latency, an epilogue job took 2,700 cycles);
//   * TMEM (512 columns) = one fp32 accumulator per slot (3 x 128) + a 64-column A-operand buffer for slots 0 and 1;
//     slot 2's A operand lives in shared memory (32 KB, K-major core matrices) and its MMAs use the SS form.  The bias
//     of the hidden/output layers is one SS-form K step whose A operand is a constant 128 x 16 tile (two columns of
//     ones) in shared memory (measured: a chain of 1 SS + 8 TS steps costs the same as 9 TS steps);
//   * shared memory holds the weights (107 KB for the 98-wide policy), slot 2's A operand and three HALF-tile
//     observation stages (64 rows each): three whole-tile stages do not fit.  Half-tile units are loaded in one global
//     order (unit n = 2 * tile + half -> stage n % 3, reusable once the tile of unit n - 3 is converted), which also
//     keeps the three groups staggered: tiles are converted in tile order;
//   * one control warp loads the observation units and issues the MMA groups in the order the groups ask for them
//     (mbarrier.test_wait round-robin over the three a_ready barriers).
#pragma once
#include "kernels_tc.cuh"

namespace go2p {

constexpr int kT3Groups = 3;
constexpr int kT3GroupWarps = 8;           // warp = 8*g + 4*half + quarter
constexpr int kT3Workers = kT3GroupWarps * kT3Groups;
constexpr int kT3CtrlWarp = kT3Workers;
constexpr int kT3Threads = (kT3Workers + 1) * 32;
constexpr uint32_t kT3ColA = 384;          // accumulator of slot g at columns 128*g, A buffer of slot g < 2 at 384 + 64*g
constexpr int kT3OnesBytes = 4096;         // constant A tile of the bias K step: 128 rows x 16 K, 16-bit
constexpr int kT3ASmemBytes = 128 * 128 * 2;   // slot 2's A operand: 128 rows x K 128, 16-bit
constexpr int kT3Layers = 4;

__host__ __device__ inline size_t tc3_smem_bytes(const TcArgs& a) {
  return tc_weight_bytes(a) + 3 * (tc_stage_bytes(a) / 2) + kT3ASmemBytes + kT3OnesBytes + 256;
}
// shapes this kernel serves: the Go2 topology (any input width the narrow family allows), 12 outputs, no activation on
// the output layer, even input width (vector loads in the conversion)
__host__ inline bool tc3_serves(const TcArgs& a) {
  return a.n_layers == kT3Layers && a.out_dim == 12 && !a.has_elu[kT3Layers - 1] && (a.in_dim & 1) == 0 && a.raw == nullptr &&
         a.sat_count == nullptr;
}

// packed registers (4 per K core of 8 elements) of one row -> slot 2's shared-memory A operand: K-major core matrices of
// 8 rows x 16 B, K-adjacent cores 128 B apart, 8-row groups 2,048 B apart.  A quarter warp (8 rows) writes 128
// contiguous bytes per store: conflict-free.
__device__ __forceinline__ void t3_store_a_smem(uint8_t* a2_row, int kcore0, const uint32_t* p, int n_cores) {
  for (int j = 0; j < n_cores; ++j)
    *reinterpret_cast<uint4*>(a2_row + (kcore0 + j) * 128) = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
}

template <bool kFp16>
__global__ void __launch_bounds__(kT3Threads, 1) tc3_mlp_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t wbytes = (uint32_t)tc_weight_bytes(a);
  const uint32_t stage_bytes = (uint32_t)tc_stage_bytes(a);
  uint8_t* w_smem = smem;
  uint8_t* stage0 = smem + wbytes;
  const uint32_t half_bytes = stage_bytes / 2;               // 64 rows
  uint8_t* a2_smem = stage0 + 3 * half_bytes;
  uint32_t* ones = reinterpret_cast<uint32_t*>(a2_smem + kT3ASmemBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ones) + kT3OnesBytes);
  uint64_t* obs_full = bars;        // [3][2]  half-tile observation stage landed, even / odd use of the stage (see load_unit)
  uint64_t* acc_full = bars + 6;    // [3]  slot's MMA group complete (tcgen05.commit); 4 phases per tile: parity 0,1,0,1
  uint64_t* a_ready = bars + 9;     // [3]  A operand of the slot's next layer written (8 arrivals); 4 phases per tile
  uint64_t* w_full = bars + 12;     // [4]  layer weights landed (completes once)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12 + kT3Layers);

  if (warp == kT3CtrlWarp) {
    if (lane == 0) {
      for (int s = 0; s < 9; ++s) ptx::mbar_init(&bars[s], 1);
      for (int s = 0; s < kT3Groups; ++s) ptx::mbar_init(&a_ready[s], kT3GroupWarps);
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
    // ================= control warp: bulk-copy producer + MMA issuer, in the order the groups ask =================
    const uint32_t fmt = kFp16 ? ptx::FMT_F16 : ptx::FMT_BF16;
    const uint32_t w_base = ptx::smem_u32(w_smem);
    const uint64_t ones_desc = ptx::make_smem_desc_nosw(ptx::smem_u32(ones), 128u, 256u);
    const uint64_t a2_desc = ptx::make_smem_desc_nosw(ptx::smem_u32(a2_smem), 128u, 2048u);
    const uint32_t idesc_h = ptx::make_idesc(fmt, kTcTileM, kTcHidden), idesc_o = ptx::make_idesc(fmt, kTcTileM, kTcOutPad);
    // K step of 16 elements = two 8x16B core matrices along K (128 B apart); 8-row groups are kp*16 B apart;
    // consecutive K steps are 256 B apart (+16 in the descriptor's address field)
    const uint64_t bdesc_l0 = ptx::make_smem_desc_nosw(w_base, 128u, (uint32_t)a.k0p * 16u);
    const uint64_t bdesc_l1 = ptx::make_smem_desc_nosw(w_base + (uint32_t)(a.k0p * kTcHidden * 2), 128u, (uint32_t)(kTcHidden + kTcBiasK) * 16u);
    const uint32_t w_step16 = (uint32_t)((kTcHidden + kTcBiasK) * kTcHidden * 2) >> 4;   // hidden layer stride in descriptor units
    const int ksteps0 = a.k0p / 16;
    const uint32_t unit_bytes = (uint32_t)(kTcTileM / 2 * a.in_dim * 4);
    // half-tile unit n = 2 * tile + half -> stage n % 3, use k = n / 3 of that stage.  A stage has two barriers, for
    // its even and odd uses: a parity wait is only sound if the waiter has seen the barrier's previous phase complete,
    // and uses k and k - 2 of a stage belong to consecutive tiles of the SAME group (units n and n - 6), whereas use
    // k - 1 belongs to another group.  A ragged last tile is read from global memory by its group; its units only
    // complete the barrier phase.
    auto load_unit = [&](int n) {
      const long long row0 = (long long)(blockIdx.x + (n >> 1) * gridDim.x) * kTcTileM;
      const int st = n % 3;
      uint64_t* bar = &obs_full[2 * st + ((n / 3) & 1)];
      if (a.B - row0 >= kTcTileM) {
        ptx::mbar_arrive_expect_tx(bar, unit_bytes);
        ptx::bulk_g2s(stage0 + st * half_bytes, a.obs + (row0 + (n & 1) * (kTcTileM / 2)) * a.in_dim, unit_bytes, bar);
      } else {
        ptx::mbar_arrive(bar);
      }
    };
    int next_unit = min(3, 2 * n_local);     // units [0, next_unit) requested; unit n may go once tile (n - 3) / 2 is converted
    if (ptx::elect_one_sync()) {
      for (int n = 0; n < next_unit; ++n) load_unit(n);
    }
    __syncwarp();
    for (int l = 0; l < kT3Layers; ++l) ptx::mbar_wait(&w_full[l], 0u);

    // per group: the next operand it will publish is that of layer `lay` of its tile `til` (phase parity `par`)
    int lay[kT3Groups] = {0, 0, 0}, til[kT3Groups] = {0, 1, 2};
    uint32_t par[kT3Groups] = {0u, 0u, 0u};
    int live = min(n_local, kT3Groups);
    while (live > 0) {
#pragma unroll
      for (int g = 0; g < kT3Groups; ++g) {
        if (til[g] >= n_local) continue;
        if (!__any_sync(0xffffffffu, ptx::mbar_test_wait(&a_ready[g], par[g]))) continue;
        ptx::tc_fence_after();
        const int l = lay[g];
        if (ptx::elect_one_sync()) {
          TC_TRACE(0x200u | (uint32_t)(l << 4) | (uint32_t)g);
          const uint32_t d_t = tmem_base + 128u * (uint32_t)g;
          const uint32_t a_t = tmem_base + kT3ColA + 64u * (uint32_t)g;
          const uint64_t bdesc = (l == 0) ? bdesc_l0 : bdesc_l1 + (uint64_t)((uint32_t)(l - 1) * w_step16);
          const uint32_t idesc = (l == kT3Layers - 1) ? idesc_o : idesc_h;
          const int ksteps = (l == 0) ? ksteps0 : kTcHidden / 16;
          uint32_t acc = 0u;
          if (l > 0) {                              // bias K step (SS form); layer 0 carries its ones inside the data
            ptx::mma_f16_ss(d_t, ones_desc, bdesc + (uint64_t)(8 * 16), idesc, 0u);
            acc = 1u;
          }
          if (g < 2) {
            for (int j = 0; j < ksteps; ++j, acc = 1u)
              ptx::mma_f16_ts(d_t, a_t + (uint32_t)(8 * j), bdesc + (uint64_t)(j * 16), idesc, acc);
          } else {
            for (int j = 0; j < ksteps; ++j, acc = 1u)
              ptx::mma_f16_ss(d_t, a2_desc + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), idesc, acc);
          }
          ptx::mma_commit(&acc_full[g]);
          TC_TRACE(0x300u | (uint32_t)(l << 4) | (uint32_t)g);
          // a conversion released the stages of its tile's two units (tiles are converted in tile order)
          if (l == 0) {
            const int lim = min(2 * til[g] + 5, 2 * n_local);
            for (int n = next_unit; n < lim; ++n) load_unit(n);
          }
        }
        __syncwarp();
        if (l == 0) next_unit = max(next_unit, min(2 * til[g] + 5, 2 * n_local));
        par[g] ^= 1u;
        if (++lay[g] == kT3Layers) {
          lay[g] = 0;
          til[g] += kT3Groups;
          if (til[g] >= n_local) --live;
        }
      }
    }
  } else {
    // ================= worker groups: group g owns slot g and the CTA-local tiles g, g+3, g+6, ... =================
    const int g = warp >> 3;
    const int hh = (warp >> 2) & 1;          // column half of the accumulator / K half of the A operand
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch (== warp % 4)
    const int m = quarter * 32 + lane;       // row inside the tile
    const uint32_t t_lane = tmem_base + (((uint32_t)(quarter * 32)) << 16);
    const uint32_t d_t = t_lane + 128u * (uint32_t)g;            // this row's accumulator
    const uint32_t a_t = t_lane + kT3ColA + 64u * (uint32_t)g;   // this row's A operand (slots 0, 1)
    uint8_t* a2_row = a2_smem + (m >> 3) * 2048 + (m & 7) * 16;  // this row's A operand (slot 2)
    const bool a_in_smem = g == 2;
    const int n8 = a.k0p / 16;
    const int c8_lo = hh ? (n8 + 1) / 2 : 0, c8_hi = hh ? n8 : (n8 + 1) / 2;
    const bool masked = (a.flags & 5u) && a.button0 != nullptr;
    uint64_t* my_acc = &acc_full[g];
    uint64_t* my_ready = &a_ready[g];

    // publish an A operand: stores complete -> visible to the tensor core -> the control warp issues the MMA group
    auto publish = [&]() {
      if (a_in_smem) ptx::fence_proxy_async_smem(); else ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(my_ready);
    };
    auto store_a = [&](int kstep, const uint32_t (&p)[8]) {      // K = 16*kstep .. +15 of this row
      if (a_in_smem) t3_store_a_smem(a2_row, 2 * kstep, p, 2);
      else ptx::tmem_st_x8(a_t + (uint32_t)(8 * kstep), p);
    };

    for (int tile = g; tile < n_local; tile += kT3Groups) {
      const long long row0 = (long long)(blockIdx.x + tile * gridDim.x) * kTcTileM;
      const long long row = row0 + m;
      const int b0 = (masked && hh == 0 && row < a.B) ? __ldg(a.button0 + row) : 0;

      // ---- conv: fp32 observation row -> 16-bit layer-0 A operand (ones at K = in_dim, in_dim+1); this warp's K half
      {
        const int unit = 2 * tile + (quarter >> 1);
        const int st = unit % 3, use = unit / 3;
        ptx::mbar_wait(&obs_full[2 * st + (use & 1)], (uint32_t)((use >> 1) & 1));
        TC_TRACE(0x400u | (uint32_t)g);
        const bool full = a.B - row0 >= kTcTileM;
        const float* srow = full ? reinterpret_cast<const float*>(stage0 + st * half_bytes) + ((quarter & 1) * 32 + lane) * a.in_dim
                                 : a.obs + row * a.in_dim;
        const bool live_row = row < a.B;
#pragma unroll 1
        for (int c8 = c8_lo; c8 < c8_hi; ++c8) {
          uint32_t q[8];
          if (full && c8 * 16 + 16 <= a.in_dim) {
            const float2* r2 = reinterpret_cast<const float2*>(srow);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float2 t = r2[c8 * 8 + j]; q[j] = kFp16 ? ptx::pack_f16_sat(t.x, t.y) : ptx::pack_bf16(t.x, t.y); }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int kk = c8 * 16 + 2 * j;
              float lo = 0.f, hi = 0.f;
              if (kk < a.in_dim) { if (live_row) lo = srow[kk]; } else if (kk <= a.in_dim + 1) lo = 1.f;
              if (kk + 1 < a.in_dim) { if (live_row) hi = srow[kk + 1]; } else if (kk + 1 <= a.in_dim + 1) hi = 1.f;
              q[j] = kFp16 ? ptx::pack_f16_sat(lo, hi) : ptx::pack_bf16(lo, hi);
            }
          }
          store_a(c8, q);
        }
      }
      publish();
      TC_TRACE(0x500u | (uint32_t)g);

      // ---- E(l): this warp's 64 accumulator columns -> ELU -> 16-bit A operand of layer l+1.  16-column chunks, the
      //      load of the next two chunks in flight while a chunk is evaluated and stored.
#pragma unroll 1
      for (int l = 0; l < kT3Layers - 1; ++l) {
        const bool he = a.has_elu[l] != 0;
        const float c = a.elu_c[l];
        uint32_t cur[16], nxt[16], pk[8];
        const uint32_t d_h = d_t + (uint32_t)(64 * hh);
        ptx::mbar_wait(my_acc, (uint32_t)(l & 1));
        ptx::tc_fence_after();
        TC_TRACE(0x600u | (uint32_t)(l << 4) | (uint32_t)g);
        ptx::tmem_ld_x16(d_h, cur);
        ptx::tmem_ld_x16(d_h + 16u, nxt);
        ptx::tc_wait_ld();
        elu_pack16<kFp16>(cur, he, c, pk);
        ptx::tmem_ld_x16(d_h + 32u, cur);
        store_a(4 * hh, pk);
        elu_pack16<kFp16>(nxt, he, c, pk);
        ptx::tmem_ld_x16(d_h + 48u, nxt);
        store_a(4 * hh + 1, pk);
        ptx::tc_wait_ld();
        elu_pack16<kFp16>(cur, he, c, pk);
        store_a(4 * hh + 2, pk);
        elu_pack16<kFp16>(nxt, he, c, pk);
        store_a(4 * hh + 3, pk);
        publish();
        TC_TRACE(0x700u | (uint32_t)(l << 4) | (uint32_t)g);
      }

      // ---- out: accumulator of the output layer -> (+clamp/mask) (+q_des, motor commands) -> global; one row per
      //      thread of the group's first four warps.  Every warp waits: its next conversion store (A operand) must not
      //      pass the output layer's MMA group.
      {
        ptx::mbar_wait(my_acc, 1u);
        ptx::tc_fence_after();
        TC_TRACE(0x800u | (uint32_t)g);
        if (hh == 0) {
          uint32_t ov[16];
          ptx::tmem_ld_x16(d_t, ov);
          ptx::tc_wait_ld();
          // the loads are ordered before this warp's next arrive (the conversion of the group's next tile); the
          // layer-0 MMA group that overwrites the accumulator is issued after all eight such arrivals
          ptx::tc_fence_before();
          if (row < a.B) {
            float o[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) o[j] = __uint_as_float(ov[j]);
            if (a.flags & 1u) {
#pragma unroll
              for (int j = 0; j < 12; ++j) o[j] = clamp_mask(o[j], a.action_limit, b0);
            }
            float4* arow = reinterpret_cast<float4*>(a.act + row * 12);
            arow[0] = make_float4(o[0], o[1], o[2], o[3]);
            arow[1] = make_float4(o[4], o[5], o[6], o[7]);
            arow[2] = make_float4(o[8], o[9], o[10], o[11]);
            if (a.flags & 6u) {
              if ((a.flags & 2u) && a.qdes) {
                double2* q2 = reinterpret_cast<double2*>(a.qdes + row * kDof);
#pragma unroll
                for (int j = 0; j < 6; ++j)
                  q2[j] = make_double2(joint_target(o[2 * j], a.q0[2 * j], a.action_scale), joint_target(o[2 * j + 1], a.q0[2 * j + 1], a.action_scale));
              }
              if ((a.flags & 4u) && a.cmd) {
                MotorCmdDev* cm = a.cmd + row;             // Isaac joint i -> Unitree motor ((i & 3) ^ 1) * 3 + (i >> 2)
#pragma unroll
                for (int j = 0; j < 12; ++j) cm->q_des[((j & 3) ^ 1) * 3 + (j >> 2)] = joint_target(o[j], a.q0[j], a.action_scale);
                store_gains(a.cmd, row, b0, a.kp, a.kd, a.kp_deadman);
              }
            }
          }
        }
        TC_TRACE(0x900u | (uint32_t)g);
      }
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
