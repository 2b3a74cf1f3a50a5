// Measured experiment (round 2), not part of the product build: see README.md in this directory.
// Drop-in for kernels_wide.cuh (kernel before "host side", launcher before wide_release, dispatch in wide_launch:
//   if (L.NT == 256 && !last && m_tiles > 1) wd_launch_gemm2<fp16>(a, sm_count, st)).

// Hidden layers with 256-wide N tiles: TWO row tiles per job share one weight tile ("M256" as two M128 MMAs per K step).
// The single-tile kernel above asks every SM for 48 KB of operands per 512 tensor cycles and gets ~46 B/clk from L2
// (profiles/r01_summary.md: tensor pipe 51 % in the K = 1024 layer): the weight tile is two thirds of that.  Sharing it
// between two row tiles of the same CTA makes it 64 KB per 1,024 tensor cycles -- a third fewer bytes per flop, which is
// what the CTA-pair variant (cta_group::2) also bought, but without a second CTA in the per-chunk hand-off.  The price:
// two 256-column accumulators fill TMEM, so the epilogue of a job no longer overlaps the next job's MMAs; the operand
// ring (3 stages of 64 KB) keeps filling during the epilogue, and the operand intake, not the tensor pipe, is the
// bound, so most of that gap is hidden.
template <bool kFp16>
__global__ void __launch_bounds__(kWdThreads, 1) wide_gemm2_kernel(const WideGemmArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NT = 256;
  constexpr int kBTileBytes = NT * kWdChunkK * 2;
  constexpr int kStageBytes = 2 * kWdATileBytes + kBTileBytes;    // 64 KB: A tile 0, A tile 1, weight tile
  constexpr int kStages = 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                 // [kStages] chunk landed
  uint64_t* empty = bars + kStages;      // [kStages] chunk consumed by the tensor core
  uint64_t* acc_full = bars + 2 * kStages;       // both accumulators complete
  uint64_t* acc_empty = bars + 2 * kStages + 1;  // both accumulators drained (kWdEpiWarps arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
      ptx::mbar_init(acc_full, 1);
      ptx::mbar_init(acc_empty, kWdEpiWarps);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc<2 * NT>(tmem_ptr);
  }
  ptx::tc_fence_before();
  block_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::grid_dependency_wait();            // programmatic dependent launch, as in wide_gemm_kernel
  ptx::grid_launch_dependents();

  // job = (pair of row tiles, N tile), N tile fastest: the two A tiles stay in L2 for the next N tile.  With an odd
  // number of row tiles the last pair's second tile is a clamped duplicate (loaded and multiplied, never stored).
  const long long m_pairs = ((long long)a.m_tiles + 1) / 2;
  const long long n_jobs = m_pairs * a.n_tiles;
  if (warp == 0) {
    // ================= producer =================
    uint32_t it = 0;
    for (long long job = blockIdx.x; job < n_jobs; job += gridDim.x) {
      const long long mt0 = (job / a.n_tiles) * 2, mt1 = min(mt0 + 1, (long long)a.m_tiles - 1);
      const int nt = (int)(job % a.n_tiles);
      for (int kc = 0; kc < a.k_chunks; ++kc, ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&empty[s], ((it / kStages) & 1u) ^ 1u);
        if (ptx::elect_one_sync()) {
          uint8_t* st = smem + s * kStageBytes;
          ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)kStageBytes);
          const uint8_t* abase = reinterpret_cast<const uint8_t*>(a.a);
          ptx::bulk_g2s(st, abase + (mt0 * a.k_chunks + kc) * (long long)kWdATileBytes, kWdATileBytes, &full[s]);
          ptx::bulk_g2s(st + kWdATileBytes, abase + (mt1 * a.k_chunks + kc) * (long long)kWdATileBytes, kWdATileBytes, &full[s]);
          ptx::bulk_g2s(st + 2 * kWdATileBytes, reinterpret_cast<const uint8_t*>(a.w) + ((long long)nt * a.k_chunks + kc) * (long long)kBTileBytes,
                        kBTileBytes, &full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = ptx::make_idesc(kFp16 ? ptx::FMT_F16 : ptx::FMT_BF16, kWdTileM, (uint32_t)NT);
    uint32_t it = 0, job_i = 0;
    for (long long job = blockIdx.x; job < n_jobs; job += gridDim.x, ++job_i) {
      ptx::mbar_wait(acc_empty, (job_i & 1u) ^ 1u);
      ptx::tc_fence_after();
      for (int kc = 0; kc < a.k_chunks; ++kc, ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&full[s], (it / kStages) & 1u);
        ptx::tc_fence_after();
        if (ptx::elect_one_sync()) {
          const uint32_t sa = ptx::smem_u32(smem + s * kStageBytes);
          const uint64_t adesc0 = ptx::make_smem_desc_nosw(sa, 128u, 1024u);
          const uint64_t adesc1 = ptx::make_smem_desc_nosw(sa + kWdATileBytes, 128u, 1024u);
          const uint64_t bdesc = ptx::make_smem_desc_nosw(sa + 2 * kWdATileBytes, 128u, 1024u);
#pragma unroll
          for (int j = 0; j < kWdChunkK / 16; ++j) {
            ptx::mma_f16_ss(tmem_base, adesc0 + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), idesc, (kc | j) ? 1u : 0u);
            ptx::mma_f16_ss(tmem_base + (uint32_t)NT, adesc1 + (uint64_t)(j * 16), bdesc + (uint64_t)(j * 16), idesc, (kc | j) ? 1u : 0u);
          }
          ptx::mma_commit(&empty[s]);
          if (kc == a.k_chunks - 1) ptx::mma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps (lane quarter = warp % 4, a quarter of the columns each) =================
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    const uint32_t lane_addr = ((uint32_t)(quarter * 32)) << 16;
    const int r = quarter * 32 + lane;
    constexpr int kShare = NT / kWdEpiParts;
    const long long next_chunks = (long long)a.n_tiles * (NT / kWdChunkK);
    uint32_t job_i = 0;
    for (long long job = blockIdx.x; job < n_jobs; job += gridDim.x, ++job_i) {
      const long long mt0 = (job / a.n_tiles) * 2;
      const int nt = (int)(job % a.n_tiles);
      ptx::mbar_wait(acc_full, job_i & 1u);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const long long mt = mt0 + t;
        if (mt >= a.m_tiles) break;                       // clamped duplicate: nothing to store
        const uint32_t acc_t = tmem_base + (uint32_t)(t * NT) + lane_addr;
#pragma unroll 1
        for (int c0 = part * kShare; c0 < (part + 1) * kShare; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld_x32(acc_t + (uint32_t)c0, v);
          ptx::tc_wait_ld();
          const float* bl = a.bias + nt * NT + c0;
          uint32_t p[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x0 = __uint_as_float(v[2 * j]) + bl[2 * j];
            float x1 = __uint_as_float(v[2 * j + 1]) + bl[2 * j + 1];
            if (a.has_elu) {
              const float e0 = fmaf(a.alpha, ptx::ex2_approx(x0 * 1.4426950408889634f), -a.alpha);
              const float e1 = fmaf(a.alpha, ptx::ex2_approx(x1 * 1.4426950408889634f), -a.alpha);
              x0 = (x0 < 0.f) ? e0 : x0;
              x1 = (x1 < 0.f) ? e1 : x1;
            }
            p[j] = kFp16 ? ptx::pack_f16_sat(x0, x1) : ptx::pack_bf16(x0, x1);
          }
          const int kn = nt * NT + c0;                   // K index of the next layer
          uint8_t* dst = reinterpret_cast<uint8_t*>(a.out_blocked) + ((mt * next_chunks + kn / kWdChunkK) * (long long)kWdATileBytes) +
                         wd_tile_offset(r, kn % kWdChunkK);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(dst + q * 128) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
    }
  }

  ptx::tc_fence_before();
  block_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<2 * NT>(tmem_base);
  }
}

// launch of wide_gemm2_kernel (hidden layers with NT = 256)
#ifndef GO2P_WD_TWO_TILES
#define GO2P_WD_TWO_TILES 1
#endif
template <bool kFp16>
inline cudaError_t wd_launch_gemm2(const WideGemmArgs& a, int sm_count, cudaStream_t st) {
  const size_t smem = (size_t)3 * (2 * kWdATileBytes + 256 * kWdChunkK * 2) + 256;
  auto kernel = wide_gemm2_kernel<kFp16>;
  static thread_local int configured_dev = -1;           // the attribute is per device and per function
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kWdThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const long long jobs = ((long long)a.m_tiles + 1) / 2 * a.n_tiles;
  cfg.gridDim = dim3((unsigned)std::min<long long>(jobs, (long long)sm_count));
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

