// 2-CTA (cta_group::2) variant of the split-bf16 GEMM: a CTA pair on one TPC computes a 256 x BN output tile with
// M=256 tcgen05.mma instructions.  Each CTA stages its own 128 rows of A and only HALF of the B tile, so the
// shared-memory fill per MMA (the measured limiter of the 1-CTA kernel: ~19 B/clk/SM of TMA ingest against a
// 1536-clk tensor budget per k-block, profiles/README.md) drops from 96 KB to 72 KB per k-block and a third
// pipeline stage fits.
//
// Pair protocol (rank 0 = leader):
//   * both CTAs' producers TMA-load into their OWN smem but complete_tx on the LEADER's full barrier
//     (cp.async.bulk.tensor ... .cta_group::2 with a mapa-translated barrier address);
//   * only the leader issues tcgen05.mma.cta_group::2; tcgen05.commit ... .multicast::cluster releases the ring slot /
//     publishes the accumulator in BOTH CTAs;
//   * each CTA's epilogue warps fold and store their own 128 rows; "buffer drained" arrives go to the leader.
#pragma once
#include "gemm_sm100.cuh"

namespace oob {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_any<true>(uint32_t td, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  umma_bf16_2sm(td, da, db, idesc, acc);
}
// arrive on the barrier at this smem offset in both CTAs of the pair when the issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 0x3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <int BN, int BK>
__host__ __device__ constexpr int gemm2_stage_bytes(int nsplit) {
  return nsplit * (GEMM_BM + BN / 2) * BK * 2;   // per CTA
}

// grid: x = 2 * ceil(M / 256) (cluster dims (2,1,1): the two CTAs of a pair are adjacent in x), y = ceil(N / BN)
template <int BN, bool A_MN, bool B_MN, int BK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16x3_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                        const GemmParams p, const int num_stages) {
  static_assert(BN == 128, "epilogue holds BN fp32 running sums per thread");
  constexpr int BNH = BN / 2;   // B rows staged by each CTA
  using C = KCfg<BK>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int nsplit = p.nsplit;
  const int a_bytes = nsplit * GEMM_BM * BK * 2;
  const int b_bytes = nsplit * BNH * BK * 2;
  const int stage_bytes = a_bytes + b_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)num_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tmem_full_bar = empty_bar + num_stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]  (the leader's copies are the live ones)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m0 = (blockIdx.x >> 1) * (2 * GEMM_BM) + (int)rank * GEMM_BM;   // this CTA's 128 rows of the pair tile
  const int n0 = blockIdx.y * BN;                                           // pair tile columns
  const int nb0 = n0 + (int)rank * BNH;                                     // the half of B this CTA stages
  const int num_kb = (p.K + BK - 1) / BK;
  const int chunk_kb = p.chunk_kb > 0 ? p.chunk_kb : GEMM_CHUNK_ELEMS / BK;
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;
  constexpr uint32_t TMEM_COLS = 512;   // main[2] + corr = 3 x BN columns, rounded to a power of two

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);    // leader: its own arrive.expect_tx; bytes from both CTAs' TMA
      mbar_init(&empty_bar[s], 1);   // one multicast commit per use
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 8);  // 4 epilogue warps x 2 CTAs arrive at the leader
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % num_stages;
        if (kb >= num_stages) mbar_wait(&empty_bar[s], ((kb / num_stages) - 1) & 1);
        if (!elect_one()) continue;
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * stage_bytes);
        const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
        const int k0 = kb * BK;
        if constexpr (!A_MN) {
          tma_load_3d_2sm(sa, &tma_a, bar, k0, m0, 0);  // box {BK, 128, nsplit}
        } else {
#pragma unroll
          for (int i = 0; i < GEMM_BM / C::ATOM; ++i)
            tma_load_3d_2sm(sa + (size_t)i * nsplit * BK * C::ROWB, &tma_a, bar, m0 + i * C::ATOM, k0, 0);
        }
        if constexpr (!B_MN) {
          tma_load_3d_2sm(sb, &tma_b, bar, k0, nb0, 0);  // box {BK, BN/2, nsplit}
        } else {
#pragma unroll
          for (int i = 0; i < BNH / C::ATOM; ++i)
            tma_load_3d_2sm(sb + (size_t)i * nsplit * BK * C::ROWB, &tma_b, bar, nb0 + i * C::ATOM, k0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {   // convergent warp, one elected lane issues (see gemm_sm100_persistent.cuh)
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int kb = 0;
      uint32_t acc_corr = 0u;
      for (int c = 0; c < num_chunks; ++c) {
        const int buf = c & 1;
        if (c >= 2) {
          mbar_wait(&tmem_empty_bar[buf], ((c >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t t_main = tmem_base + (uint32_t)(buf * BN);
        const uint32_t t_corr = tmem_base + (uint32_t)(2 * BN);
        const int kb_end = min(kb + chunk_kb, num_kb);
        uint32_t acc_main = 0u;
        for (; kb < kb_end; ++kb) {
          const int s = kb % num_stages;
          mbar_wait(&full_bar[s], (kb / num_stages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          if (elect_one()) {
            if (nsplit == 3) issue_kblock<3, BNH, A_MN, B_MN, true, BK>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
            else if (nsplit == 2) issue_kblock<2, BNH, A_MN, B_MN, true, BK>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
            else issue_kblock<1, BNH, A_MN, B_MN, true, BK>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
            umma_commit_2sm(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(&tmem_full_bar[buf]);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..5, both CTAs) =====================
    const int quarter = warp & 3;
    const int row = m0 + quarter * 32 + lane;
    const bool has_corr = nsplit > 1;
    float racc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) racc[j] = 0.f;
    for (int c = 0; c < num_chunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&tmem_full_bar[buf], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const bool last_chunk = (c == num_chunks - 1);
#pragma unroll
      for (int g = 0; g < BN / 32; ++g) {
        uint32_t v[32];
        tmem_ld_32x32(t_lane + (uint32_t)(buf * BN + g * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) racc[g * 32 + j] += __uint_as_float(v[j]);
        if (has_corr && last_chunk) {
          tmem_ld_32x32(t_lane + (uint32_t)(2 * BN + g * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[g * 32 + j] += __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty_bar[buf]);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[buf]), 0));
      }
    }
    if (row < p.M) {
#pragma unroll
      for (int g = 0; g < BN / 32; ++g) {
        const int col0 = n0 + g * 32;
        if (col0 < p.N) {
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = racc[g * 32 + j];
          epilogue_store32(x, p.epi, row, col0, p.N);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody may exit (or free TMEM) while the peer can still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace oob
