// EXPERIMENT, not the default (capi.cu: GO2P_TC32).  Measured on B200: 0.304 ms per 1,048,576 rows against 0.212 ms for
// the 18-warp kernel -- all parity tests pass, but halving the job size doubles the per-job fixed cost (accumulator
// wait, tcgen05.ld/st round trips, fences, arrival) that every warp pays, and that cost is issue work, not idle time.
// The lesson for the next round is the opposite direction: fewer, larger jobs per warp.
//
// Batched policy forward, 32-warp variant of tc_mlp_kernel (kernels_tc.cuh; same data path, same operand images,
// same MMA order -- bit-identical outputs), reorganised for thread-level parallelism:
//
//   * ALL 32 warps of the CTA are workers: 4 TMEM lane quarters (warp % 4) x 8 column groups of 16 accumulator
//     columns (warp / 4).  The timeline of the 18-warp kernel showed each scheduler running 4 worker warps whose jobs
//     are ~35 % latency (tcgen05.ld / st / fences / hand-off); with 8 half-size jobs per scheduler that latency is
//     covered by the other warps' arithmetic.
//   * There are no control warps.  A job ends with an acq_rel atomic increment of the slot's arrival counter; the warp
//     that completes the count of 32 ("last arriver") issues the slot's next-layer MMAs (and, at layer 0, the bulk copy
//     of the slot's next tile) before it goes on -- no wake-up of a parked control warp in the hand-off.
//   * A-operand layout: K chunk j (16 values = 8 packed columns) of the next layer's A lives in columns [16j, 16j+8)
//     of the buffer, i.e. inside the 16 accumulator columns its producer warp has already loaded, so the in-place
//     store never touches a column another warp still has to read; the constant-one (bias) chunk sits in columns
//     [8, 16), written by column group 0.
#pragma once
#include "kernels_tc.cuh"

namespace go2p {

constexpr int kTc32Warps = 32;
constexpr int kTc32Threads = kTc32Warps * 32;
constexpr int kTc32Groups = 8;       // column groups of 16 accumulator columns

// column of K chunk c8 (16 values = 8 packed columns) of a layer's A operand: chunk j < 8 at 16j; a ninth chunk
// (layer-0 inputs wider than 126) and the bias chunk of the deeper layers use columns [8, 16)
__device__ __forceinline__ uint32_t tc32_chunk_col(int c8) { return c8 < 8 ? (uint32_t)(16 * c8) : 8u; }

template <bool kFp16>
__device__ __noinline__ void tc32_conv_slow(const TcArgs& a, const float* rowp, bool live, int c8_first, int n8, uint32_t a0_t) {
  for (int c8 = c8_first; c8 < n8; c8 += kTc32Groups) {
    uint32_t q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c8 * 16 + 2 * j;
      float lo = 0.f, hi = 0.f;
      if (k < a.in_dim) { if (live) lo = rowp[k]; } else if (k <= a.in_dim + 1) lo = 1.f;
      if (k + 1 < a.in_dim) { if (live) hi = rowp[k + 1]; } else if (k + 1 <= a.in_dim + 1) hi = 1.f;
      q[j] = kFp16 ? ptx::pack_f16_sat(lo, hi) : ptx::pack_bf16(lo, hi);
    }
    ptx::tmem_st_x8(a0_t + tc32_chunk_col(c8), q);
  }
}

template <bool kFp16>
__global__ void __launch_bounds__(kTc32Threads, 1) tc_mlp32_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const size_t wbytes = tc_weight_bytes(a);
  const size_t stage_bytes = tc_stage_bytes(a);
  uint8_t* w_smem = smem;
  uint8_t* stage0 = smem + wbytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + 2 * stage_bytes);
  uint64_t* obs_full = bars;        // [2]
  uint64_t* acc_full = bars + 4;    // [2]
  uint64_t* w_full = bars + 14;     // [kMaxLayers]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14 + kMaxLayers);
  uint32_t* arrivals = tmem_ptr + 2;   // [2] monotonic per-slot job arrival counters (32 per job)

  const long long n_tiles = (a.B + kTcTileM - 1) / kTcTileM;
  const int n_local = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles b, b+G, ...
  const uint32_t tile_bytes = (uint32_t)(kTcTileM * a.in_dim * 4);

  auto load_tile = [&](int i) {            // one thread; i = CTA-local tile index, slot = i & 1
    const int s = i & 1;
    const long long row0 = (blockIdx.x + (long long)i * gridDim.x) * kTcTileM;
    if (a.B - row0 >= kTcTileM) {
      ptx::mbar_arrive_expect_tx(&obs_full[s], tile_bytes);
      ptx::bulk_g2s(stage0 + s * stage_bytes, a.obs + row0 * a.in_dim, tile_bytes, &obs_full[s]);
    } else {
      ptx::mbar_arrive(&obs_full[s]);      // ragged last tile: the workers read global memory directly
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&obs_full[s], 1); ptx::mbar_init(&acc_full[s], 1); arrivals[s] = 0u; }
      for (int l = 0; l < a.n_layers; ++l) ptx::mbar_init(&w_full[l], 1);
      ptx::fence_mbar_init();
      uint32_t off = 0;
      for (int l = 0; l < a.n_layers; ++l) {
        const uint32_t bytes = (uint32_t)(tc_layer_kp(a, l) * tc_layer_n(a, l) * 2);
        ptx::mbar_arrive_expect_tx(&w_full[l], bytes);
        ptx::bulk_g2s(w_smem + off, reinterpret_cast<const uint8_t*>(a.wpack) + off, bytes, &w_full[l]);
        off += bytes;
      }
      for (int i = 0; i < min(2, n_local); ++i) load_tile(i);
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::tc_fence_before();
  block_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  TC_TRACE_INIT();

  const int quarter = warp & 3;            // TMEM lane quarter this warp may touch (== warp % 4)
  const int cg = warp >> 2;                // column group: accumulator columns [16cg, 16cg+16)
  const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
  const int m = quarter * 32 + lane;       // row inside the tile
  const uint32_t one2 = kFp16 ? 0x3C003C00u : 0x3F803F80u;   // packed (1.0, 1.0)
  const int L = a.n_layers - 1;            // index of the output layer
  const bool out12 = a.out_dim == 12;
  const int n8 = a.k0p / 16;               // layer-0 A operand: chunks of 16 K elements
  const bool even = (a.in_dim & 1) == 0;
  const uint32_t fmt = kFp16 ? ptx::FMT_F16 : ptx::FMT_BF16;
  const uint32_t w_base = ptx::smem_u32(w_smem);

  // MMAs of layer l of the slot's tile `i` (CTA-local index), issued by one elected lane of the calling warp.
  // K step j reads A chunk j at column 16j of the source buffer; the bias chunk (layers >= 1) sits at column 8.
  // Order: chunk 0, chunk 1, bias, chunks 2..7 -- the order of the 18-warp kernel, so results are bit-identical.
  auto issue_layer = [&](int i, int l) {
    const int s = i & 1, phi = (i >> 1) & 1;
    uint32_t w_off = 0;
    for (int q = 0; q < l; ++q) w_off += (uint32_t)(tc_layer_kp(a, q) * tc_layer_n(a, q) * 2);
    const int kp = tc_layer_kp(a, l), nl = tc_layer_n(a, l);
    const uint32_t idesc = ptx::make_idesc(fmt, kTcTileM, (uint32_t)nl);
    const uint64_t bdesc0 = ptx::make_smem_desc_nosw(w_base + w_off, 128u, (uint32_t)kp * 16u);
    const uint32_t src = tmem_base + (uint32_t)s * kTcSlotCols + 128u * (uint32_t)((phi + l) & 1);
    const uint32_t dst = tmem_base + (uint32_t)s * kTcSlotCols + 128u * (uint32_t)((phi + l + 1) & 1);
    ptx::mbar_wait(&w_full[l], 0u);        // completes once; later waits return at the first probe
    ptx::tc_fence_after();
    if (ptx::elect_one_sync()) {
      TC_TRACE(0x200u | (uint32_t)(l << 4) | (uint32_t)s);
      if (l == 0) {
        if (i + 2 < n_local) load_tile(i + 2);      // every warp has converted this slot's stage: it is free
        const int ksteps = kp / 16;                 // constant ones inside the data (K = in_dim, in_dim + 1)
        for (int j = 0; j < ksteps; ++j)
          ptx::mma_f16_ts(dst, src + tc32_chunk_col(j), bdesc0 + (uint64_t)(j * 16), idesc, j > 0 ? 1u : 0u);
      } else {
        ptx::mma_f16_ts(dst, src, bdesc0, idesc, 0u);
        ptx::mma_f16_ts(dst, src + 16u, bdesc0 + 16u, idesc, 1u);
        ptx::mma_f16_ts(dst, src + 8u, bdesc0 + (uint64_t)(8 * 16), idesc, 1u);                    // bias K step
#pragma unroll
        for (int j = 2; j < 8; ++j) ptx::mma_f16_ts(dst, src + (uint32_t)(16 * j), bdesc0 + (uint64_t)(j * 16), idesc, 1u);
      }
      ptx::mma_commit(&acc_full[s]);
      TC_TRACE(0x300u | (uint32_t)(l << 4) | (uint32_t)s);
    }
    __syncwarp();
  };

  // end of a job on slot s of tile i: the stores are complete and ordered, the arrival is counted, and the warp that
  // completes the 32 arrivals issues layer `lnext` of that tile
  auto finish_job = [&](int i, int lnext) {
    const int s = i & 1;
    ptx::tc_wait_st();
    ptx::tc_fence_before();
    __syncwarp();
    uint32_t old = 0;
    if (lane == 0) asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(ptx::smem_u32(&arrivals[s])) : "memory");
    old = __shfl_sync(0xffffffffu, old, 0);
    if ((old & (kTc32Warps - 1)) == kTc32Warps - 1) issue_layer(i, lnext);
  };

  uint32_t par_acc[2] = {0u, 0u};

  // ---- conv(i): fp32 observation rows -> 16-bit layer-0 A operand (chunk c8 at column 16*c8 of buffer phi)
  auto conv_job = [&](int i) {
    const int s = i & 1, phi = (i >> 1) & 1;
    const long long row0 = (blockIdx.x + (long long)i * gridDim.x) * kTcTileM;
    const int valid = (int)min((long long)kTcTileM, a.B - row0);
    const uint32_t a0_t = tmem_base + (uint32_t)s * kTcSlotCols + lane_addr + 128u * (uint32_t)phi;
    ptx::mbar_wait(&obs_full[s], (uint32_t)((i >> 1) & 1));
    TC_TRACE(0x400u | (uint32_t)s);
    if (valid == kTcTileM && even) {
      const float2* r2 = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(stage0 + s * stage_bytes) + (size_t)m * a.in_dim);
#pragma unroll 1
      for (int c8 = cg; c8 < n8; c8 += kTc32Groups) {
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
        ptx::tmem_st_x8(a0_t + tc32_chunk_col(c8), q);
      }
    } else {
      const float* rowp = (valid == kTcTileM) ? reinterpret_cast<const float*>(stage0 + s * stage_bytes) + (size_t)m * a.in_dim
                                              : a.obs + (row0 + m) * a.in_dim;
      tc32_conv_slow<kFp16>(a, rowp, m < valid, cg, n8, a0_t);
    }
    finish_job(i, 0);
    TC_TRACE(0x500u | (uint32_t)s);
  };

  // ---- out(i) (bias already inside the accumulator): (+ELU) (+clamp/mask) (+q_des) -> global
  auto out_job = [&](int i) {
    const int s = i & 1, phi = (i >> 1) & 1;
    const long long row0 = (blockIdx.x + (long long)i * gridDim.x) * kTcTileM;
    const int valid = (int)min((long long)kTcTileM, a.B - row0);
    const uint32_t o_t = tmem_base + (uint32_t)s * kTcSlotCols + lane_addr + 128u * (uint32_t)((phi + 1 + L) & 1);
    ptx::mbar_wait(&acc_full[s], par_acc[s]);
    par_acc[s] ^= 1u;
    ptx::tc_fence_after();
    TC_TRACE(0x800u | (uint32_t)s);
    if (out12 && !a.has_elu[L]) {
      if (cg < 3) {      // 12 outputs, no activation: column group cg < 3 stores one float4 of every row
        uint32_t v[4];
        ptx::tmem_ld_x4(o_t + (uint32_t)(cg * 4), v);
        ptx::tc_wait_ld();
        if (m < valid) {
          const long long row = row0 + m;
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = __uint_as_float(v[j]);
          if (a.flags & 1u) {
            const int b0 = a.button0 ? a.button0[row] : 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = clamp_mask(o[j], a.action_limit, b0);
          }
          reinterpret_cast<float4*>(a.act + row * 12)[cg] = make_float4(o[0], o[1], o[2], o[3]);
          if ((a.flags & 2u) && a.qdes) {
            double2* q2 = reinterpret_cast<double2*>(a.qdes + row * kDof + cg * 4);
            q2[0] = make_double2(joint_target(o[0], a.q0[cg * 4 + 0], a.action_scale), joint_target(o[1], a.q0[cg * 4 + 1], a.action_scale));
            q2[1] = make_double2(joint_target(o[2], a.q0[cg * 4 + 2], a.action_scale), joint_target(o[3], a.q0[cg * 4 + 3], a.action_scale));
          }
        }
      }
    } else if (cg == 0) {
      tc_out_generic(a, o_t, row0 + m, m < valid);
    }
    ptx::tc_fence_before();
    TC_TRACE(0x900u | (uint32_t)s);
  };

  // Job order: conv of the first two tiles, then per pair  E(l,s0) E(l,s1) for every hidden layer |
  // out(s0) conv'(s0) out(s1) conv'(s1)  (conv' = the slot's next tile), as in the 18-warp kernel.
  for (int i = 0; i < min(2, n_local); ++i) conv_job(i);
  for (int pair = 0; pair * 2 < n_local; ++pair) {
    const int ns = min(2, n_local - pair * 2);
    const int phi = pair & 1;
    // ---- E(l,s): 16 accumulator columns -> ELU -> 8 packed columns of the next layer's A operand, in place
    for (int l = 0; l < L; ++l) {
      const bool he = a.has_elu[l] != 0;
      const float c = a.elu_c[l];
      for (int s = 0; s < ns; ++s) {
        const uint32_t d_t = tmem_base + (uint32_t)s * kTcSlotCols + lane_addr + 128u * (uint32_t)((phi + 1 + l) & 1) + (uint32_t)(cg * 16);
        ptx::mbar_wait(&acc_full[s], par_acc[s]);
        par_acc[s] ^= 1u;
        ptx::tc_fence_after();
        TC_TRACE(0x600u | (uint32_t)(l << 4) | (uint32_t)s);
        uint32_t cur[16], pk[8];
        ptx::tmem_ld_x16(d_t, cur);
        ptx::tc_wait_ld();
        elu_pack16<kFp16>(cur, he, c, pk);
        ptx::tmem_st_x8(d_t, pk);
        if (cg == 0) {   // constant-one chunk (K = 128,129; zeros up to 143) in columns [8,16) of the buffer
          const uint32_t ones[8] = {one2, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          ptx::tmem_st_x8(d_t + 8u, ones);
        }
        finish_job(pair * 2 + s, l + 1);
        TC_TRACE(0x700u | (uint32_t)(l << 4) | (uint32_t)s);
      }
    }
    for (int s = 0; s < ns; ++s) {
      out_job(pair * 2 + s);
      // even number of hidden layers: the output accumulator shares its buffer with the next tile's layer-0 operand;
      // the eight warps of a lane quarter meet before any of them converts the next tile (see kernels_tc.cuh)
      if ((L & 1) == 0) {
        __syncwarp();
        asm volatile("bar.sync %0, 256;" ::"r"(1 + quarter) : "memory");
      }
      if ((pair + 1) * 2 + s < n_local) conv_job((pair + 1) * 2 + s);
    }
  }

  // ---- teardown
  ptx::tc_fence_before();
  block_sync();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace go2p
