// Persistent split-plane GEMM (1-CTA and 2-CTA/cta_group::2 in one template): one CTA (pair) per SM (pair) walks a
// static list of output tiles, so the TMA ring, the MMA stream and the epilogue of consecutive tiles overlap:
//
//   producer  : keeps filling the smem ring straight across tile boundaries
//   MMA issuer: after the last chunk of tile i it starts tile i+1 at once -- "main" chunk accumulators are
//               double-buffered per chunk, the "corr" accumulator per tile parity (TMEM: 2 x main + 2 x corr = 512)
//   epilogue  : folds chunks as they complete; after the last chunk of a tile it folds corr, releases both TMEM
//               regions and only then runs the fused output stores, while the tensor pipe is already on tile i+1
#pragma once
#include "gemm_sm100_2cta.cuh"

namespace oob {

template <int BN, bool TWO_CTA, int BK>
__host__ __device__ constexpr int gemmp_stage_bytes(int nsplit) {
  return nsplit * (GEMM_BM + (TWO_CTA ? BN / 2 : BN)) * BK * 2;   // per CTA
}

// grid.x = number of CTAs (1-CTA) or 2 x number of pairs (2-CTA, cluster dims (2,1,1)); tiles are distributed
// round-robin: unit u (CTA or pair) takes tiles u, u + units, u + 2*units, ...; tile t -> (t % tiles_m, t / tiles_m)
// so that the units running concurrently share the same B panel (L2 reuse) and stream different A row blocks.
template <int BN, bool A_MN, bool B_MN, bool TWO_CTA, int BK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16x3_persistent_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                              const __grid_constant__ GemmParams p, const int num_stages) {
  static_assert(BN == 128, "epilogue holds BN fp32 running sums per thread; TMEM = 4 x BN columns");
  constexpr int BROWS = TWO_CTA ? BN / 2 : BN;        // B rows staged per CTA
  constexpr int TILE_M = TWO_CTA ? 2 * GEMM_BM : GEMM_BM;
  using C = KCfg<BK>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int nsplit = p.nsplit;
  const int a_bytes = nsplit * GEMM_BM * BK * 2;
  const int b_bytes = nsplit * BROWS * BK * 2;
  const int stage_bytes = a_bytes + b_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // after the ring: one transposition tile per epilogue warp (32 rows x p.slab fp32), then the barriers
  float* stage_out = reinterpret_cast<float*>(smem + (size_t)num_stages * stage_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_out + 4 * GEMM_EPI_GROUPS * 32 * p.slab);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tmem_full_bar = empty_bar + num_stages;   // [2] chunk accumulator ready
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2] chunk accumulator drained
  uint64_t* corr_empty_bar = tmem_empty_bar + 2;       // [2] corr accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(corr_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = TWO_CTA ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = TWO_CTA ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_units = TWO_CTA ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tiles_m = (p.M + TILE_M - 1) / TILE_M;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int chunk_kb = p.chunk_kb > 0 ? p.chunk_kb : GEMM_CHUNK_ELEMS / BK;
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t kEpiArrivals = (TWO_CTA ? 8 : 4) * GEMM_EPI_GROUPS;
  constexpr int CG = BN / GEMM_EPI_GROUPS;            // columns per epilogue warp

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], kEpiArrivals);
      mbar_init(&corr_empty_bar[b], kEpiArrivals);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if constexpr (TWO_CTA) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (TWO_CTA) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {
      int g = 0;   // k-blocks issued so far (ring position); whole warp loops, one elected lane issues
      for (int t = unit; t < num_tiles; t += num_units) {
        const int m0 = (t % tiles_m) * TILE_M + (int)rank * GEMM_BM;
        const int nb0 = (t / tiles_m) * BN + (int)rank * BROWS;
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % num_stages;
          if (g >= num_stages) mbar_wait(&empty_bar[s], ((g / num_stages) - 1) & 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          const int k0 = kb * BK;
          if (!elect_one()) continue;
          if (p.debug & 2) {   // diagnostics: no loads, the consumer runs on whatever is in smem
            if (leader) mbar_arrive(&full_bar[s]);
            continue;
          }
          if constexpr (TWO_CTA) {
            if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * stage_bytes);
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
            if constexpr (!A_MN) tma_load_3d_2sm(sa, &tma_a, bar, k0, m0, 0);
            else
#pragma unroll
              for (int i = 0; i < GEMM_BM / C::ATOM; ++i)
                tma_load_3d_2sm(sa + (size_t)i * nsplit * BK * C::ROWB, &tma_a, bar, m0 + i * C::ATOM, k0, 0);
            if constexpr (!B_MN) tma_load_3d_2sm(sb, &tma_b, bar, k0, nb0, 0);
            else
#pragma unroll
              for (int i = 0; i < BROWS / C::ATOM; ++i)
                tma_load_3d_2sm(sb + (size_t)i * nsplit * BK * C::ROWB, &tma_b, bar, nb0 + i * C::ATOM, k0, 0);
          } else {
            mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
            if constexpr (!A_MN) tma_load_3d(sa, &tma_a, &full_bar[s], k0, m0, 0);
            else
#pragma unroll
              for (int i = 0; i < GEMM_BM / C::ATOM; ++i)
                tma_load_3d(sa + (size_t)i * nsplit * BK * C::ROWB, &tma_a, &full_bar[s], m0 + i * C::ATOM, k0, 0);
            if constexpr (!B_MN) tma_load_3d(sb, &tma_b, &full_bar[s], k0, nb0, 0);
            else
#pragma unroll
              for (int i = 0; i < BROWS / C::ATOM; ++i)
                tma_load_3d(sb + (size_t)i * nsplit * BK * C::ROWB, &tma_b, &full_bar[s], nb0 + i * C::ATOM, k0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    // The whole warp walks the loop (convergent control flow) and ONE elected lane issues: with a divergent
    // `if (lane == 0)` around it the compiler wraps every uniform-datapath instruction (UTCHMMA, UTCBAR) in an
    // ELECT/PLOP3/BRA.U.ANY retry loop, ~40 clk of issue per 64-clk MMA.
    if (leader) {
      const uint32_t idesc = make_idesc_f16kind(TILE_M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0, 0, 0) |
                             ((uint32_t)p.a_bf16 << 7) | ((uint32_t)p.b_bf16 << 10);
      int g = 0;    // k-blocks consumed so far
      int gc = 0;   // chunks issued so far (main buffer = gc & 1)
      int ti = 0;   // tiles started so far (corr buffer = ti & 1)
      for (int t = unit; t < num_tiles; t += num_units, ++ti) {
        const int cb = ti & 1;
        const uint32_t t_corr = tmem_base + (uint32_t)((2 + cb) * BN);
        if (ti >= 2) {   // corr[cb] was last used by tile ti-2: its fold must be over
          mbar_wait(&corr_empty_bar[cb], ((ti >> 1) - 1) & 1);
          tc_fence_after();
        }
        uint32_t acc_corr = 0u;
        int kb = 0;
        for (int c = 0; c < num_chunks; ++c, ++gc) {
          const int buf = gc & 1;
          if (gc >= 2) {
            mbar_wait(&tmem_empty_bar[buf], ((gc >> 1) - 1) & 1);
            tc_fence_after();
          }
          const uint32_t t_main = tmem_base + (uint32_t)(buf * BN);
          const int kb_end = min(kb + chunk_kb, num_kb);
          uint32_t acc_main = 0u;
          for (; kb < kb_end; ++kb, ++g) {
            const int s = g % num_stages;
            mbar_wait(&full_bar[s], (g / num_stages) & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
            const uint32_t sb = sa + a_bytes;
            if (elect_one()) {
              if (p.debug & 4) { /* diagnostics: no MMAs */ }
              else if (nsplit == 3) issue_kblock<3, BROWS, A_MN, B_MN, TWO_CTA, BK>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
              else if (nsplit == 2) issue_kblock<2, BROWS, A_MN, B_MN, TWO_CTA, BK>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
              else issue_kblock<1, BROWS, A_MN, B_MN, TWO_CTA, BK>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
              if constexpr (TWO_CTA) umma_commit_2sm(&empty_bar[s]);
              else umma_commit(&empty_bar[s]);
            }
            __syncwarp();
          }
          if (elect_one()) {
            if constexpr (TWO_CTA) umma_commit_2sm(&tmem_full_bar[buf]);
            else umma_commit(&tmem_full_bar[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2 ..) =====================
    const int quarter = warp & 3;
    const int cgrp = (warp - 2) >> 2;                    // which CG-column slice of the tile this warp owns
    const bool has_corr = nsplit > 1;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cgrp * CG);
    int gc = 0, ti = 0;
    for (int t = unit; t < num_tiles; t += num_units, ++ti) {
      const int m0 = (t % tiles_m) * TILE_M + (int)rank * GEMM_BM;
      const int n0 = (t / tiles_m) * BN + cgrp * CG;
      const int cb = ti & 1;
      float racc[CG];
#pragma unroll
      for (int j = 0; j < CG; ++j) racc[j] = 0.f;
      for (int c = 0; c < num_chunks; ++c, ++gc) {
        const int buf = gc & 1;
        mbar_wait(&tmem_full_bar[buf], (gc >> 1) & 1);
        tc_fence_after();
        const bool last_chunk = (c == num_chunks - 1);
#pragma unroll
        for (int g = 0; g < CG / 32; ++g) {
          uint32_t v[32];
          tmem_ld_32x32(t_lane + (uint32_t)(buf * BN + g * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[g * 32 + j] += __uint_as_float(v[j]);
          if (has_corr && last_chunk) {
            tmem_ld_32x32(t_lane + (uint32_t)((2 + cb) * BN + g * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) racc[g * 32 + j] = fmaf(__uint_as_float(v[j]), p.corr_scale, racc[g * 32 + j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) {
            mbar_arrive(&tmem_empty_bar[buf]);
            if (last_chunk) mbar_arrive(&corr_empty_bar[cb]);
          } else {
            mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[buf]), 0));
            if (last_chunk) mbar_arrive_cluster(mapa_u32(smem_u32(&corr_empty_bar[cb]), 0));
          }
        }
      }
      // fused output for this tile; the tensor pipe is already working on the next one
      if (!(p.debug & 1)) {
        float* my_stage = stage_out + (warp - 2) * 32 * p.slab;
        const int row0 = m0 + quarter * 32;
        if (p.slab == 32) {
#pragma unroll
          for (int g = 0; g < CG / 32; ++g)
            if (n0 + g * 32 < p.N) epilogue_slab<32>(&racc[g * 32], my_stage, p.epi, row0, n0 + g * 32, p.M, p.N, lane);
        } else {
#pragma unroll
          for (int g = 0; g < CG / 16; ++g)
            if (n0 + g * 16 < p.N) epilogue_slab<16>(&racc[g * 16], my_stage, p.epi, row0, n0 + g * 16, p.M, p.N, lane);
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (TWO_CTA) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (TWO_CTA) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace oob
