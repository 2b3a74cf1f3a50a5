// Batched forward for policies whose layers are too wide for the all-in-shared-memory kernel (kernels_tc.cuh), e.g.
// BASELINE.json configs[4]: 245 -> 1024 -> 512 -> 256 -> 12 with a 5-frame observation history.
//
//   A7  Gemm/Elu chain   reference: onnx_actor.cpp:38-48 (Ort::Session::Run), one tcgen05 GEMM launch per layer
//   A9/A11 epilogue      reference: controller.cpp:217-223,244 (fused into the last layer's epilogue)
//
// Design:
//   * activations travel between layers as 16-bit values in a TILE-BLOCKED layout in global memory (they stay L2
//     resident for the chunk sizes the host uses): tiles of 128 rows x 64 K-columns, each tile stored exactly as the
//     UMMA K-major no-swizzle core-matrix image the tensor core reads from shared memory (8x8-element cores of 128
//     contiguous bytes, K-adjacent cores contiguous, 8-row groups 1024 B apart).  A tile is therefore ONE contiguous
//     16 KB chunk: a single 1-D bulk async copy (TMA engine) brings it in, no tensor map and no relayout;
//   * weights are packed on the host the same way per (N-tile, K-chunk);
//   * wide_gemm_kernel: persistent CTAs over (row tile, N tile); warp 0 = producer (3-stage mbarrier ring of A/B
//     chunks), warp 1 = MMA issuer (SS-form tcgen05.mma, M=128, N=NT, four K steps per 64-wide chunk, accumulators
//     double-buffered in TMEM so the epilogue of one tile overlaps the MMAs of the next), warps 2-17 = epilogue
//     (tcgen05.ld -> +bias -> ELU -> 16-bit -> blocked store for the next layer, or the fused A9/A11 output epilogue);
//   * obs_to_blocked_kernel converts the fp32 [B,in] observations into the blocked 16-bit layout (zero padded).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "onnx_reader.hpp"
#include "policy_dev.cuh"
#include "ptx_sm100.cuh"

namespace go2p {

constexpr int kWdTileM = 128;
constexpr int kWdChunkK = 64;                       // K columns per pipeline stage (4 MMA K steps)
constexpr int kWdATileBytes = kWdTileM * kWdChunkK * 2;   // 16 KB
constexpr int kWdStages = 4;                        // 4 x 48 KB (N tile 256) = 192 KB of operands in flight per SM
// operand ring depth: as many stages as fit in ~192 KB (the L2 round trip under load is ~4,000 cycles, so the bytes in
// flight, not the stage count, set the delivered bandwidth), at most 8
__host__ __device__ constexpr int wd_stages(int nt, bool pair) {
  const int stage = 128 * 64 * 2 + (pair ? nt / 2 : nt) * 64 * 2;
  const int n = (192 * 1024) / stage;
  return n > 8 ? 8 : n;
}
#ifndef GO2P_WD_EPI_WARPS
#define GO2P_WD_EPI_WARPS 16
#endif
constexpr int kWdEpiWarps = GO2P_WD_EPI_WARPS;      // kWdEpiWarps/4 per TMEM lane quarter, each takes a share of the tile's columns
constexpr int kWdEpiParts = kWdEpiWarps / 4;
constexpr int kWdThreads = (2 + kWdEpiWarps) * 32;  // producer, MMA issuer, epilogue warps

// byte offset of element (r, c) inside a [rows x 64] blocked tile (rows multiple of 8)
__host__ __device__ inline uint32_t wd_tile_offset(int r, int c) {
  return (uint32_t)(((r >> 3) * 8 + (c >> 3)) * 128 + (r & 7) * 16 + (c & 7) * 2);
}

struct WideLayerDev {
  const uint16_t* w[2];    // [0] bf16, [1] fp16: per (n tile, k chunk) blocked image, NT rows x 64
  const float* bias;       // [Npad]
  int K, N, Kp, Np, NT;    // Kp multiple of 64, Np multiple of NT
  int has_elu;
  float alpha;
};

struct WideModel {
  int n_layers = 0;
  int in_dim = 0, out_dim = 0;
  WideLayerDev L[kMaxLayers];
  int max_kp = 0;          // widest blocked activation (in K columns)
  // activation ping-pong buffers (allocated lazily by wide_launch for a chunk of rows)
  // one pair per stream set (capi.cu: scratch_sel; the host-buffer pipeline runs several streams concurrently)
  static constexpr int kSets = 4;
  mutable uint16_t* act[kSets][2] = {};
  mutable long long act_rows[kSets] = {};
  mutable std::vector<void*> retired;     // outgrown buffers: freed with the model (cudaFree would synchronise the device)
};

struct WideGemmArgs {
  const uint16_t* a;        // blocked activations [Mp/128][Kp/64] tiles
  const uint16_t* w;        // blocked weights     [Np/NT][Kp/64] tiles
  const float* bias;
  uint16_t* out_blocked;    // next layer's blocked activations (Kp_next = Np), or null for the output layer
  float* out_rows;          // [M, out_dim] fp32 (output layer)
  const int32_t* button0;
  double* qdes;
  MotorCmdDev* cmd;         // send_command arguments in Unitree motor order (flag 4) or null
  float kp, kd, kp_deadman;
  long long M;              // valid rows
  int m_tiles, n_tiles, k_chunks;
  int N, out_dim, has_elu;
  float alpha;
  uint32_t flags;
  float action_limit;
  double action_scale;
  double q0[kDof];
};

// fp32 [B, in] row-major -> blocked 16-bit tiles, zero padded to (m_tiles*128) x Kp.
// A warp converts 8 rows x 32 columns: lane = (column group of 8) * 8 + (row & 7), so its 32 16-byte stores are the four
// K-adjacent cores of one 8-row group -- 512 contiguous bytes (the first version walked along a row: every lane's
// 16 bytes went to a different 128-byte line, and the conversion took 34 us of a 195-us pass).
template <bool kFp16>
__global__ void obs_to_blocked_kernel(const float* __restrict__ obs, uint16_t* __restrict__ out, long long B, int in_dim, int Kp,
                                      long long m_tiles) {
  const int kg4 = Kp / 32;                                           // groups of four 8-column cores per row
  const long long items = m_tiles * (kWdTileM / 8) * kg4;            // one warp per (8-row group, 32 columns)
  const int lane = threadIdx.x & 31;
  const int r8 = lane & 7, cgl = lane >> 3;
  // programmatic dependent launch: the previous pass may still be reading the buffer this kernel rewrites
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long it = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += warps) {
    const long long rg = it / kg4;                                   // global 8-row group
    const int kg = (int)(it % kg4) * 4 + cgl;                        // this lane's 8-column core
    const long long row = rg * 8 + r8;
    const float* src = obs + row * in_dim + kg * 8;
    uint32_t p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kg * 8 + 2 * j;
      const float lo = (row < B && k < in_dim) ? src[2 * j] : 0.f;
      const float hi = (row < B && k + 1 < in_dim) ? src[2 * j + 1] : 0.f;
      p[j] = kFp16 ? ptx::pack_f16_sat(lo, hi) : ptx::pack_bf16(lo, hi);
    }
    const long long mt = row / kWdTileM;
    const int r = (int)(row % kWdTileM);
    const int kc = (kg * 8) / kWdChunkK, c = (kg * 8) % kWdChunkK;
    uint8_t* dst = reinterpret_cast<uint8_t*>(out) + ((mt * (Kp / kWdChunkK) + kc) * (long long)kWdATileBytes) + wd_tile_offset(r, c);
    *reinterpret_cast<uint4*>(dst) = make_uint4(p[0], p[1], p[2], p[3]);
  }
}

// kCS = thread-block cluster size (1 or 2).  With kCS = 2 the two CTAs of a cluster work on neighbouring row tiles of
// the SAME N tile: each loads its own A tile and HALF of the weight tile, which the TMA engine multicasts into both
// CTAs' shared memory -- the weight tile crosses the L2 -> SM fabric once per cluster instead of once per CTA (32 KB
// instead of 48 KB of operand traffic per CTA and K chunk at NT = 256).  A stage is refilled only when the MMAs of
// BOTH CTAs have read it: tcgen05.commit arrives on the stage's empty barrier of every CTA in the cluster.
// kCS = 2 with kPair: the two CTAs form a CTA PAIR (tcgen05 cta_group::2).  One M256 x NT MMA spans both SMs: each CTA
// holds its own 128-row A tile and HALF of the weight tile (NT/2 rows), so an SM takes in 32 KB instead of 48 KB per K
// chunk at NT = 256 -- the remedy for the L2 -> SM intake limit of the single-CTA tile.  The leader (rank 0) issues the
// MMAs once its own and the peer's stage have landed (the peer relays its full-barrier to the leader), and its
// tcgen05.commit releases stages and publishes accumulators in both CTAs; both CTAs' epilogue warps report the drained
// accumulator to the leader.
template <int NT, bool kFp16, int kCS, bool kPair = false>
__global__ void __launch_bounds__(kWdThreads, 1) wide_gemm_kernel(const WideGemmArgs a) {
  static_assert(!kPair || kCS == 2, "a CTA pair is a cluster of two");
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kBTileBytes = NT * kWdChunkK * 2;
  constexpr int kStageBytes = kWdATileBytes + (kPair ? kBTileBytes / 2 : kBTileBytes);   // pair: half of the weight tile per CTA
  constexpr int kStages = wd_stages(NT, kPair);
  constexpr uint32_t kAccCols = NT < 32 ? 32 : NT;               // TMEM columns per accumulator buffer
  constexpr uint32_t kTmemCols = 2 * kAccCols;                   // 64 .. 512, a power of two
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                 // [kStages] chunk landed
  uint64_t* empty = bars + kStages;    // [kStages] chunk consumed by the tensor core
  uint64_t* acc_full = bars + 2 * kStages;       // [2]
  uint64_t* acc_empty = bars + 2 * kStages + 2;  // [2]
  uint64_t* peer_full = bars + 2 * kStages + 4;  // [kStages] leader only: the peer's chunk landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 4);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init(&full[s], 1);
        ptx::mbar_init(&empty[s], kPair ? 1 : kCS);      // pair: one multicast commit of the leader; multicast: one per CTA
        ptx::mbar_init(&peer_full[s], 1);
      }
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], (kPair ? 2 : 1) * kWdEpiWarps); }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    if (kPair) ptx::tmem_alloc_pair<kTmemCols>(tmem_ptr); else ptx::tmem_alloc<kTmemCols>(tmem_ptr);
  }
  ptx::tc_fence_before();
  block_sync();
  if (kCS > 1) ptx::cluster_sync_all();     // the peer's barriers exist before anything is multicast into them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Programmatic dependent launch: barriers and TMEM are set up while the previous layer's launch drains; its
  // activations are read (and the buffer two layers back is overwritten) only after that launch has completed.  The
  // next layer's launch may begin its own set-up as soon as this CTA's SM is free.  (5 launches per 18,944-row pass:
  // the gaps between them were a quarter of the wide policy's time.)
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();

  // job = (group of kCS row tiles, N tile), N tile fastest: the A tiles stay in L2 for the next N tile.
  // Every CTA of a cluster walks the same job list; CTA rank r takes row tile group*kCS + r.  A rank whose row tile
  // does not exist (odd tile count) still loads and multiplies a clamped tile -- its half of the weights is needed by
  // the peer -- and only skips the stores.
  const uint32_t crank = kCS > 1 ? ptx::cluster_ctarank() : 0u;
  const long long job0 = kCS > 1 ? (long long)ptx::cluster_id_x() : (long long)blockIdx.x;
  const long long job_stride = kCS > 1 ? (long long)ptx::cluster_nctaid_x() : (long long)gridDim.x;
  const long long m_groups = (a.m_tiles + kCS - 1) / kCS;
  const long long n_jobs = m_groups * a.n_tiles;
  auto job_mt = [&](long long job) { return (job / a.n_tiles) * kCS + crank; };
  if (warp == 0) {
    // ================= producer =================
    uint32_t it = 0;
    for (long long job = job0; job < n_jobs; job += job_stride) {
      const long long mt = min(job_mt(job), (long long)a.m_tiles - 1);
      const int nt = (int)(job % a.n_tiles);
      for (int kc = 0; kc < a.k_chunks; ++kc, ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&empty[s], ((it / kStages) & 1u) ^ 1u);
        if (ptx::elect_one_sync()) {
          uint8_t* st = smem + s * kStageBytes;
          ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(kPair ? kWdATileBytes + kBTileBytes / 2 : kStageBytes));
          ptx::bulk_g2s(st, reinterpret_cast<const uint8_t*>(a.a) + (mt * a.k_chunks + kc) * (long long)kWdATileBytes, kWdATileBytes, &full[s]);
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w) + ((long long)nt * a.k_chunks + kc) * (long long)kBTileBytes;
          if (kPair) {
            // this CTA's NT/2 rows of the weight tile, at the same shared-memory offset in both CTAs
            ptx::bulk_g2s(st + kWdATileBytes, wsrc + crank * (kBTileBytes / 2), kBTileBytes / 2, &full[s]);
          } else if (kCS > 1) {
            constexpr uint32_t kPart = kBTileBytes / kCS;     // whole 8-row groups: a contiguous slice of the tile
            ptx::bulk_g2s_multicast(st + kWdATileBytes + crank * kPart, wsrc + crank * kPart, kPart, &full[s], (uint16_t)((1u << kCS) - 1u));
          } else {
            ptx::bulk_g2s(st + kWdATileBytes, wsrc, kBTileBytes, &full[s]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (pair mode: leader only; the peer's warp relays its full barriers) =================
    const uint32_t idesc = ptx::make_idesc(kFp16 ? ptx::FMT_F16 : ptx::FMT_BF16, kPair ? 2 * kWdTileM : kWdTileM, (uint32_t)NT);
    uint32_t it = 0, tile_i = 0;
    if (kPair && crank != 0) {
      for (long long job = job0; job < n_jobs; job += job_stride)
        for (int kc = 0; kc < a.k_chunks; ++kc, ++it) {
          const int s = it % kStages;
          ptx::mbar_wait(&full[s], (it / kStages) & 1u);
          if (lane == 0) ptx::mbar_arrive_cluster(&peer_full[s], 0u);
          __syncwarp();
        }
    } else {
      for (long long job = job0; job < n_jobs; job += job_stride, ++tile_i) {
        const uint32_t buf = tile_i & 1u;
        ptx::mbar_wait(&acc_empty[buf], ((tile_i >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_t = tmem_base + buf * kAccCols;
        for (int kc = 0; kc < a.k_chunks; ++kc, ++it) {
          const int s = it % kStages;
          ptx::mbar_wait(&full[s], (it / kStages) & 1u);
          if (kPair) ptx::mbar_wait(&peer_full[s], (it / kStages) & 1u);
          ptx::tc_fence_after();
          if (ptx::elect_one_sync()) {
            const uint32_t sa = ptx::smem_u32(smem + s * kStageBytes);
            const uint64_t adesc = ptx::make_smem_desc_nosw(sa, 128u, 1024u);
            const uint64_t bdesc = ptx::make_smem_desc_nosw(sa + kWdATileBytes, 128u, 1024u);
#pragma unroll
            for (int j = 0; j < kWdChunkK / 16; ++j) {
              if (kPair) ptx::mma_f16_ss_pair(d_t, adesc + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), idesc, (kc | j) ? 1u : 0u);
              else ptx::mma_f16_ss(d_t, adesc + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), idesc, (kc | j) ? 1u : 0u);
            }
            // frees the stage (in every CTA of the cluster) when these MMAs have read it
            if (kPair) ptx::mma_commit_pair(&empty[s], 3);
            else if (kCS > 1) ptx::mma_commit_multicast(&empty[s], (uint16_t)((1u << kCS) - 1u));
            else ptx::mma_commit(&empty[s]);
            if (kc == a.k_chunks - 1) { if (kPair) ptx::mma_commit_pair(&acc_full[buf], 3); else ptx::mma_commit(&acc_full[buf]); }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================= epilogue warps (lane quarter = warp % 4) =================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                // which share of the N tile's columns this warp converts
    const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
    const int r = quarter * 32 + lane;               // row inside the tile
    uint32_t tile_i = 0;
    for (long long job = job0; job < n_jobs; job += job_stride, ++tile_i) {
      const long long mt = job_mt(job);
      const bool ghost = mt >= a.m_tiles;              // clamped duplicate of the last row tile: no stores
      const int nt = (int)(job % a.n_tiles);
      const uint32_t buf = tile_i & 1u;
      ptx::mbar_wait(&acc_full[buf], (tile_i >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t acc_t = tmem_base + buf * kAccCols + lane_addr;
      const long long row = mt * kWdTileM + r;
      if (ghost) {
        // nothing to store; the accumulator is simply released
      } else if (a.out_blocked) {
        // hidden layer: 32 columns at a time -> bias + ELU -> 16 bit -> four 16-byte stores into the next layer's tile
        const int kcn_per_tile = NT / kWdChunkK;       // 64-column chunks of the next layer covered by this N tile
        const long long next_chunks = (long long)a.n_tiles * kcn_per_tile;
        constexpr int kShare = (NT / kWdEpiParts) < 32 ? 32 : (NT / kWdEpiParts);   // columns per warp, whole 32-column groups
#pragma unroll 1
        for (int c0 = half * kShare; c0 < (half + 1) * kShare && c0 < NT; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld_x32(acc_t + (uint32_t)c0, v);
          ptx::tc_wait_ld();
          // the 32 biases of this column group as eight 16-byte loads (warp-uniform address, 128-byte aligned)
          float bl[32];
          {
            const float4* b4 = reinterpret_cast<const float4*>(a.bias + nt * NT + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 t = __ldg(b4 + j); bl[4 * j] = t.x; bl[4 * j + 1] = t.y; bl[4 * j + 2] = t.z; bl[4 * j + 3] = t.w; }
          }
          uint32_t p[16];
          // the activation test is hoisted out of the unrolled loop: inside it nvcc emitted a branch with its
          // reconvergence bookkeeping and a constant-bank reload of alpha per PAIR (15+ instead of 11 instructions)
          if (a.has_elu) {
            const float alpha = a.alpha, nalpha = -a.alpha;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x0 = __uint_as_float(v[2 * j]) + bl[2 * j];
              float x1 = __uint_as_float(v[2 * j + 1]) + bl[2 * j + 1];
              const float e0 = fmaf(alpha, ptx::ex2_approx(x0 * 1.4426950408889634f), nalpha);
              const float e1 = fmaf(alpha, ptx::ex2_approx(x1 * 1.4426950408889634f), nalpha);
              x0 = (x0 < 0.f) ? e0 : x0;
              x1 = (x1 < 0.f) ? e1 : x1;
              p[j] = kFp16 ? ptx::pack_f16_sat(x0, x1) : ptx::pack_bf16(x0, x1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x0 = __uint_as_float(v[2 * j]) + bl[2 * j];
              const float x1 = __uint_as_float(v[2 * j + 1]) + bl[2 * j + 1];
              p[j] = kFp16 ? ptx::pack_f16_sat(x0, x1) : ptx::pack_bf16(x0, x1);
            }
          }
          const int kn = nt * NT + c0;                 // K index of the next layer
          uint8_t* dst = reinterpret_cast<uint8_t*>(a.out_blocked) + ((mt * next_chunks + kn / kWdChunkK) * (long long)kWdATileBytes) +
                         wd_tile_offset(r, kn % kWdChunkK);
#pragma unroll
          for (int q = 0; q < 4; ++q)                  // consecutive 8-column cores are 128 B apart inside the tile
            *reinterpret_cast<uint4*>(dst + q * 128) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
        }
      } else {
        // output layer (NT == 16): bias (+ELU) (+clamp/mask) (+q_des) -> fp32 rows
        uint32_t v[16];
        ptx::tmem_ld_x16(acc_t, v);
        ptx::tc_wait_ld();
        if (row < a.M && half == 0) {
          const int b0 = ((a.flags & 5u) && a.button0) ? a.button0[row] : 0;
          float* dst = a.out_rows + row * a.out_dim;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < a.out_dim) {
              float x = __uint_as_float(v[j]) + a.bias[j];
              if (a.has_elu) x = (x < 0.f) ? fmaf(a.alpha, ptx::ex2_approx(x * 1.4426950408889634f), -a.alpha) : x;
              if (a.flags & 1u) x = clamp_mask(x, a.action_limit, b0);
              dst[j] = x;
              if (j < kDof && (a.flags & 6u)) {
                const double qd = joint_target(x, a.q0[j], a.action_scale);
                if ((a.flags & 2u) && a.qdes) a.qdes[row * kDof + j] = qd;
                if ((a.flags & 4u) && a.cmd) a.cmd[row].q_des[motor_of_isaac(j)] = qd;
              }
            }
          }
          if ((a.flags & 4u) && a.cmd) store_gains(a.cmd, row, b0, a.kp, a.kd, a.kp_deadman);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (kPair) ptx::mbar_arrive_cluster(&acc_empty[buf], 0u); else ptx::mbar_arrive(&acc_empty[buf]); }
    }
  }

  ptx::tc_fence_before();
  block_sync();
  if (kCS > 1) ptx::cluster_sync_all();     // no CTA leaves while the peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    if (kPair) ptx::tmem_dealloc_pair<kTmemCols>(tmem_base); else ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host side
inline int wd_round_up(int v, int m) { return (v + m - 1) / m * m; }

// blocked image of W [N][K] for N tiles of NT rows: tile (nt, kc) = NT x 64 elements
inline void wd_pack_weights(const MlpLayer& L, int NT, int Np, int Kp, bool fp16, std::vector<uint16_t>& out) {
  out.assign((size_t)Np * Kp, 0);
  const int k_chunks = Kp / kWdChunkK;
  for (int n = 0; n < L.out; ++n)
    for (int k = 0; k < L.in; ++k) {
      const float v = L.weight[(size_t)n * L.in + k];
      const size_t tile = ((size_t)(n / NT) * k_chunks + k / kWdChunkK) * ((size_t)NT * kWdChunkK);
      const size_t idx = tile + wd_tile_offset(n % NT, k % kWdChunkK) / 2;
      uint16_t b;
      if (fp16) { const float c = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v); b = __half_as_ushort(__float2half_rn(c)); }
      else b = __bfloat16_as_ushort(__float2bfloat16_rn(v));
      out[idx] = b;
    }
}

// returns 0 = ok, > 0 = go2p_status (CUDA failure), < 0 = this shape is not served by the wide path
inline int wide_prepare(const MlpModel& m, std::vector<void*>& dev_owned, WideModel* wm, std::string& err) {
  const int nl = (int)m.layers.size();
  if (nl < 2 || nl > kMaxLayers || m.layers.back().out > 16) return -1;
  for (int l = 0; l + 1 < nl; ++l) if (m.layers[l].out % 64 != 0) return -1;
  wm->n_layers = nl; wm->in_dim = m.layers.front().in; wm->out_dim = m.layers.back().out; wm->max_kp = 0;
  auto upload = [&](const void* src, size_t bytes, void** out) -> bool {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { err = "wide_prepare: cudaMalloc failed"; return false; }
    dev_owned.push_back(p);
    if (cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { err = "wide_prepare: cudaMemcpy failed"; return false; }
    *out = p; return true;
  };
  for (int l = 0; l < nl; ++l) {
    const MlpLayer& L = m.layers[l];
    WideLayerDev& D = wm->L[l];
    const bool last = l == nl - 1;
    D.K = L.in; D.N = L.out;
    D.Kp = wd_round_up(L.in, kWdChunkK);
    D.NT = last ? 16 : (L.out % 256 == 0 ? 256 : (L.out % 128 == 0 ? 128 : 64));
    D.Np = wd_round_up(L.out, D.NT);
    D.has_elu = L.has_elu ? 1 : 0; D.alpha = L.elu_alpha;
    if (l > 0 && D.Kp != wm->L[l - 1].Np) return -1;   // a hidden width that is not a multiple of 64
    wm->max_kp = std::max(wm->max_kp, D.Kp);
    std::vector<uint16_t> img;
    for (int f = 0; f < 2; ++f) {
      wd_pack_weights(L, D.NT, D.Np, D.Kp, f == 1, img);
      void* p;
      if (!upload(img.data(), img.size() * 2, &p)) return 4;
      D.w[f] = static_cast<const uint16_t*>(p);
    }
    std::vector<float> bias(D.Np, 0.f);
    for (int n = 0; n < L.out; ++n) bias[n] = L.bias[n];
    void* pb;
    if (!upload(bias.data(), bias.size() * 4, &pb)) return 4;
    D.bias = static_cast<const float*>(pb);
  }
  return 0;
}

// Launch mode of the GEMM kernels: 1 = one CTA per tile (default), 2 = cluster of two with TMA multicast of the weight
// tile, 3 = CTA pair (tcgen05 cta_group::2, M256 tiles).  All three pass the same tests.  Measured on B200
// (scripts/gpu_wide_variants.sh, 606,208 rows, round-1 pass size): 1: 2.25 ms, 2: 2.41 ms, 3: 3.19 ms.  ncu: mode 3
// does cut the L2 -> SM bytes by a third (233 -> 155 MB in layer 2) but its MMAs retire at a quarter of the single-CTA
// rate -- the leader/peer hand-offs (full-barrier relay, multicast commits) sit in the per-chunk critical path, and a
// 1-D bulk copy cannot complete on the peer CTA's barrier.  Modes 2 and 3 are kept as measured build options.
#ifndef GO2P_WD_CLUSTER
#define GO2P_WD_CLUSTER 1
#endif

template <int NT, bool kFp16>
inline cudaError_t wd_launch_gemm(const WideGemmArgs& a, int sm_count, cudaStream_t st) {
  constexpr int kCS = GO2P_WD_CLUSTER == 3 ? 2 : GO2P_WD_CLUSTER;      // 3 = CTA pair (cta_group::2)
  constexpr bool kPair = GO2P_WD_CLUSTER == 3;
  const size_t smem = (size_t)wd_stages(NT, kPair) * (kWdATileBytes + (kPair ? NT / 2 : NT) * kWdChunkK * 2) + 256;
  auto kernel = wide_gemm_kernel<NT, kFp16, kCS, kPair>;
  static thread_local int configured_dev = -1;           // the attribute is per device and per function
  static thread_local int max_clusters = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  cfg.blockDim = dim3(kWdThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see griddepcontrol.wait in the kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (kCS > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = kCS; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    max_clusters = sm_count / kCS;
    if (kCS > 1) {
      cfg.gridDim = dim3((unsigned)(sm_count / kCS * kCS));
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);   // clusters that can be co-resident (GPC boundaries)
      if (e != cudaSuccess) return e;
      if (n > 0 && n < max_clusters) max_clusters = n;
      if (std::getenv("GO2P_DEBUG")) fprintf(stderr, "wide_gemm<%d>: %d co-resident clusters of %d (of %d SMs)\n", NT, n, kCS, sm_count);
    }
    configured_dev = dev;
  }
  const long long jobs = ((long long)a.m_tiles + kCS - 1) / kCS * a.n_tiles;
  const long long groups = std::min<long long>(jobs, (long long)max_clusters);
  cfg.gridDim = dim3((unsigned)(groups * kCS));
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

inline void wide_release(WideModel* wm) {
  for (int s = 0; s < WideModel::kSets; ++s) {
    for (int i = 0; i < 2; ++i) { if (wm->act[s][i]) cudaFree(wm->act[s][i]); wm->act[s][i] = nullptr; }
    wm->act_rows[s] = 0;
  }
  for (void* p : wm->retired) cudaFree(p);
  wm->retired.clear();
}

// rows per pass: a multiple of 148 row tiles, so every layer's job count is a whole number of waves over the 148 SMs.
// Round 1 used ONE wave of row tiles (18,944 rows: the widest activation, 39 MB, stays in the 126 MB L2); measured in
// round 2 (scripts/gpu_wide_variants.sh, 606,208 rows, fp16 / bf16): 1 wave 1.86 / 1.85 ms, 2 waves 1.62 / 1.66 ms,
// 4 waves 1.57 / 1.51 ms, 32 waves (the whole batch) 1.54 / 1.48 ms -- the five launches per pass each end in a tail
// (layers 3 and 4 are ONE job per CTA at one wave: no overlap of epilogue and MMAs at all), and that costs more than
// reading the activations back from HBM instead of L2.  Four waves: 75,776 rows, 155 MB per activation buffer.
#ifndef GO2P_WD_PASS_WAVES
#define GO2P_WD_PASS_WAVES 4
#endif
constexpr long long kWdChunkRows = 148 * kWdTileM * GO2P_WD_PASS_WAVES;

inline int wide_launch(const WideModel& wm, const float* d_obs, const int32_t* d_button0, float* d_act, double* d_qdes, MotorCmdDev* d_cmd, long long B,
                       bool fp16, uint32_t flags, const CtrlConst& cc, int sm_count, cudaStream_t st, int* launches, std::string& err, int set = 0) {
  if (set < 0 || set >= WideModel::kSets) { err = "wide_launch: bad stream set"; return 2; }
  uint16_t** act = wm.act[set];
  const long long chunk = std::min<long long>(B, kWdChunkRows);
  const long long chunk_tiles = (chunk + kWdTileM - 1) / kWdTileM;
  if (wm.act_rows[set] < chunk_tiles * kWdTileM) {
    for (int i = 0; i < 2; ++i) { if (act[i]) wm.retired.push_back(act[i]); act[i] = nullptr; }
    const size_t bytes = (size_t)chunk_tiles * kWdTileM * wm.max_kp * 2;
    if (cudaMalloc((void**)&act[0], bytes) != cudaSuccess || cudaMalloc((void**)&act[1], bytes) != cudaSuccess) {
      err = "wide_launch: cudaMalloc of the activation buffers failed"; return 4;
    }
    wm.act_rows[set] = chunk_tiles * kWdTileM;
  }
  for (long long r0 = 0; r0 < B; r0 += chunk) {
    const long long rows = std::min(chunk, B - r0);
    const int m_tiles = (int)((rows + kWdTileM - 1) / kWdTileM);
    const WideLayerDev& L0 = wm.L[0];
    {
      const long long groups = (long long)m_tiles * kWdTileM * (L0.Kp / 8);
      const int blocks = (int)std::min<long long>((groups + 255) / 256, (long long)sm_count * 8);
      cudaLaunchConfig_t lc = {};
      cudaLaunchAttribute la;
      la.id = cudaLaunchAttributeProgrammaticStreamSerialization;
      la.val.programmaticStreamSerializationAllowed = 1;
      lc.gridDim = dim3((unsigned)blocks); lc.blockDim = dim3(256); lc.stream = st; lc.attrs = &la; lc.numAttrs = 1;
      const float* src = d_obs + r0 * wm.in_dim;
      const long long m_tiles_ll = m_tiles;
      cudaError_t ec = fp16 ? cudaLaunchKernelEx(&lc, obs_to_blocked_kernel<true>, src, act[0], rows, wm.in_dim, L0.Kp, m_tiles_ll)
                            : cudaLaunchKernelEx(&lc, obs_to_blocked_kernel<false>, src, act[0], rows, wm.in_dim, L0.Kp, m_tiles_ll);
      if (ec != cudaSuccess) { err = std::string("wide_launch: ") + cudaGetErrorString(ec); return 4; }
      ++*launches;
    }
    int cur = 0;
    for (int l = 0; l < wm.n_layers; ++l) {
      const WideLayerDev& L = wm.L[l];
      const bool last = l == wm.n_layers - 1;
      WideGemmArgs a{};
      a.a = act[cur]; a.w = L.w[fp16 ? 1 : 0]; a.bias = L.bias;
      a.out_blocked = last ? nullptr : act[cur ^ 1];
      a.out_rows = last ? d_act + r0 * wm.out_dim : nullptr;
      a.button0 = d_button0 ? d_button0 + r0 : nullptr;
      a.qdes = d_qdes ? d_qdes + r0 * kDof : nullptr;
      a.cmd = d_cmd ? d_cmd + r0 : nullptr;
      a.kp = cc.kp; a.kd = cc.kd; a.kp_deadman = cc.kp_deadman;
      a.M = rows; a.m_tiles = m_tiles; a.n_tiles = L.Np / L.NT; a.k_chunks = L.Kp / kWdChunkK;
      a.N = L.N; a.out_dim = wm.out_dim; a.has_elu = L.has_elu; a.alpha = L.alpha;
      a.flags = last ? flags : 0u; a.action_limit = cc.action_limit; a.action_scale = cc.action_scale;
      for (int i = 0; i < kDof; ++i) a.q0[i] = cc.q0[i];
      cudaError_t e;
      if (L.NT == 256) e = fp16 ? wd_launch_gemm<256, true>(a, sm_count, st) : wd_launch_gemm<256, false>(a, sm_count, st);
      else if (L.NT == 128) e = fp16 ? wd_launch_gemm<128, true>(a, sm_count, st) : wd_launch_gemm<128, false>(a, sm_count, st);
      else if (L.NT == 64) e = fp16 ? wd_launch_gemm<64, true>(a, sm_count, st) : wd_launch_gemm<64, false>(a, sm_count, st);
      else e = fp16 ? wd_launch_gemm<16, true>(a, sm_count, st) : wd_launch_gemm<16, false>(a, sm_count, st);
      if (e != cudaSuccess) { err = std::string("wide_launch: ") + cudaGetErrorString(e); return 4; }
      ++*launches;
      cur ^= 1;
    }
  }
  return 0;
}

}  // namespace go2p
