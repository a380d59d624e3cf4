// tcgen05 / TMA / TMEM GEMM for the per-stage transformer layers (sm_100a).
//
//   D[M,N] (fp32) = sum over split products of A_p[M,K] * B_q[K,N]      (+ fused epilogue)
//
// Operands are split-bf16 planes (common.cuh) so that the tensor cores reproduce the reference's fp32
// Conv1D / Linear arithmetic (oobleck/execution/layer.py:104-105 pins fp32; HF GPT-2 addmm).
// Either operand may be K-major (contraction index contiguous) or MN-major (output index contiguous), which is
// what lets forward (X.W), dgrad (dY.W^T) and wgrad (X^T.dY) all read the SAME natural row-major buffers
// without a transpose pass:
//     forward  Y  = X  . W      A = X  [M,K] K-major    B = W  [K,N] N-major
//     dgrad    dX = dY . W^T    A = dY [M,N] K-major    B = W  [K,N] as [N_out=K][contract=N] K-major
//     wgrad    dW = X^T. dY     A = X  [tok,K] M-major  B = dY [tok,N] N-major
//
// Structure (one CTA per 128 x BN output tile, 192 threads):
//   warp 0    TMA producer  : cp.async.bulk.tensor.3d of all planes of the A and B k-block into a
//                             multi-stage SWIZZLE_128B shared-memory ring, mbarrier complete_tx
//   warp 1    MMA issuer    : one elected thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 in TMEM),
//                             tcgen05.commit releases ring slots / signals the epilogue; owns TMEM alloc
//   warps 2-5 epilogue      : tcgen05.ld 32x32b from TMEM, fused bias / residual / GELU / dGELU /
//                             accumulate, fp32 and/or split-bf16 stores
#pragma once
#include "common.cuh"

namespace oob {

enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_DGELU = 2 };

struct GemmEpilogue {
  float* d;             // fp32 output [M, ldd] (may be null when only planes are written)
  long ldd;
  const float* bias;    // [N] or null
  const float* resid;   // [M, ldr] or null: d = acc + bias + resid
  long ldr;
  int accumulate;       // d += value (wgrad accumulation across micro-batches)
  int act;              // GemmAct: GELU writes d = pre-activation, planes = gelu(pre);
                        //          DGELU reads aux = pre-activation, value = acc * gelu'(aux)
  const float* aux;     // [M, ldaux] for ACT_DGELU
  long ldaux;
  bf16* planes;         // split output [nplanes_out][M][ldp] or null
  long ldp;
  long plane_stride;
  int nplanes_out;
  float alpha;          // value = alpha * acc (before bias etc.)
};

struct GemmParams {
  int M, N, K;
  int nsplit;   // planes used from each operand (1..3)
  GemmEpilogue epi;
  int chunk_kb; // promotion chunk in k-blocks (0 = default GEMM_CHUNK_KB)
  int debug;    // diagnostics only (tools/gemm_sweep.py): 1 = skip the output stores
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;           // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int GEMM_THREADS = 192;

template <int BN>
__host__ __device__ constexpr int gemm_stage_bytes(int nsplit) {
  return nsplit * (GEMM_BM + BN) * GEMM_BK * 2;
}

// product list per split level (plane index of A / B for product q), see issue_kblock:
//   q:      0 1 2 3 4 5
//   A plane 0 0 1 1 0 2      B plane 0 1 0 1 2 0


// ---------------------------------------------------------------------------------------------------------------
// MMA issue for one k-block (64 contraction elements), fully unrolled with compile-time descriptor offsets.
// The issuing thread is a single lane: every extra dependent instruction per MMA is ~4-6 clk of issue latency, and a
// 128x128x16 MMA only lasts 64 clk.  (First version computed descriptors in a runtime triple loop: ~220 clk of issue
// per MMA, the GEMM ran at 26% tensor-pipe utilisation -- profiles/README.md.)
template <bool TWO_CTA>
__device__ __forceinline__ void umma_any(uint32_t td, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc);

template <int NS, int BROWS, bool A_MN, bool B_MN, bool TWO_CTA>
__device__ __forceinline__ void issue_kblock(uint32_t sa, uint32_t sb, uint32_t t_main, uint32_t t_corr,
                                             uint32_t idesc, uint32_t& acc_main, uint32_t& acc_corr) {
  constexpr uint32_t A_PLANE = (A_MN ? GEMM_BK : GEMM_BM) * 128;   // bytes between planes inside a stage
  constexpr uint32_t B_PLANE = (B_MN ? GEMM_BK : BROWS) * 128;
  constexpr uint32_t A_KSTEP = A_MN ? 2048 : 32;                   // bytes per 16 contraction elements
  constexpr uint32_t B_KSTEP = B_MN ? 2048 : 32;
  constexpr uint32_t A_LBO = A_MN ? NS * GEMM_BK * 128 : 0;        // MN-major: distance between 64-wide atoms
  constexpr uint32_t B_LBO = B_MN ? NS * GEMM_BK * 128 : 0;
  constexpr int NPROD = NS == 1 ? 1 : (NS == 2 ? 3 : 6);
  constexpr int PA[6] = {0, 0, 1, 1, 0, 2};
  constexpr int PB[6] = {0, 1, 0, 1, 2, 0};
  const uint64_t a_base = make_smem_desc(sa, A_LBO, 1024, SWZ_128B);
  const uint64_t b_base = make_smem_desc(sb, B_LBO, 1024, SWZ_128B);
#pragma unroll
  for (int k = 0; k < GEMM_BK / 16; ++k) {   // leading product p0q0 -> "main"
    umma_any<TWO_CTA>(t_main, a_base + ((k * A_KSTEP) >> 4), b_base + ((k * B_KSTEP) >> 4), idesc, acc_main);
    acc_main = 1u;
  }
#pragma unroll
  for (int k = 0; k < GEMM_BK / 16; ++k) {   // corrections -> "corr"
#pragma unroll
    for (int q = 1; q < NPROD; ++q) {
      umma_any<TWO_CTA>(t_corr, a_base + ((PA[q] * A_PLANE + k * A_KSTEP) >> 4),
                        b_base + ((PB[q] * B_PLANE + k * B_KSTEP) >> 4), idesc, acc_corr);
      acc_corr = 1u;
    }
  }
}

template <>
__device__ __forceinline__ void umma_any<false>(uint32_t td, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  umma_bf16(td, da, db, idesc, acc);
}

// Fused epilogue for 32 consecutive columns of one output row held in registers.
__device__ __forceinline__ void epilogue_store32(float (&x)[32], const GemmEpilogue& e, int row, int col0, int N) {
  const bool full = (col0 + 32 <= N);
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] *= e.alpha;
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (full || col0 + j < N) x[j] += __ldg(e.bias + col0 + j);
  }
  if (e.resid) {
    const float* r = e.resid + (long)row * e.ldr + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (full || col0 + j < N) x[j] += r[j];
  }
  if (e.accumulate) {
    const float* dprev = e.d + (long)row * e.ldd + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (full || col0 + j < N) x[j] += dprev[j];
  }
  if (e.act == ACT_DGELU) {
    const float* a = e.aux + (long)row * e.ldaux + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (full || col0 + j < N) x[j] *= gelu_new_grad_f(a[j]);
  }
  if (e.d) {
    float* dp = e.d + (long)row * e.ldd + col0;
    if (full && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) stg_f4(dp + j, make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) dp[j] = x[j];
    }
  }
  if (e.planes) {
    if (e.act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = gelu_new_f(x[j]);
    }
    // pack the split planes two bf16 per 32-bit word so the stores stay in registers
    uint32_t w[3][16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      bf16 a0, a1, a2, b0, b1, b2;
      split3(x[j], a0, a1, a2);
      split3(x[j + 1], b0, b1, b2);
      w[0][j >> 1] = (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(b0) << 16);
      w[1][j >> 1] = (uint32_t)__bfloat16_as_ushort(a1) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
      w[2][j >> 1] = (uint32_t)__bfloat16_as_ushort(a2) | ((uint32_t)__bfloat16_as_ushort(b2) << 16);
    }
    bf16* pp = e.planes + (long)row * e.ldp + col0;
    const bool vec = full && ((reinterpret_cast<uintptr_t>(pp) & 15) == 0) && ((e.plane_stride & 7) == 0);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      if (pl < e.nplanes_out) {
        bf16* dst = pp + (long)pl * e.plane_stride;
        if (vec) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(dst + 2 * j) = make_uint4(w[pl][j], w[pl][j + 1], w[pl][j + 2], w[pl][j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < N)
              dst[j] = __ushort_as_bfloat16((unsigned short)((w[pl][j >> 1] >> ((j & 1) * 16)) & 0xFFFF));
        }
      }
    }
  }
}

// Accumulation scheme.  The tensor core's fp32 accumulator TRUNCATES on every accumulate (measured on B200: the
// error of a K=1600 six-product GEMM grew linearly with the number of MMAs, ~2^-25.7 per MMA, and did not depend
// on the split level).  To keep fp32-grade results for any K:
//   * the leading product p0q0 ("main") is accumulated in chunks of GEMM_CHUNK_KB k-blocks; after every chunk the
//     epilogue warps fold the chunk into fp32 registers with round-to-nearest adds (promotion).  Two TMEM buffers
//     let the MMA warp run one chunk ahead of the fold;
//   * the correction products ("corr", 2^-8 of main and smaller) accumulate over the whole contraction in a third
//     TMEM region -- their truncation error is 2^-8 smaller still -- and are folded once at the end.
// TMEM: main[2] + corr = 3 x BN columns.  (First version folded main+corr every 4 k-blocks: the extra TMEM reads
// cost 35% of the GEMM's throughput -- profiles/README.md.)
constexpr int GEMM_CHUNK_KB = 8;   // 8 x 64 = 512 contraction elements = 32 main MMAs per chunk

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const GemmParams p, const int num_stages) {
  static_assert(BN == 128, "TMEM budget: (2 main buffers + corr) x BN columns must be <= 512");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stage ring][barriers]
  const int nsplit = p.nsplit;
  const int a_bytes = nsplit * GEMM_BM * GEMM_BK * 2;
  const int b_bytes = nsplit * BN * GEMM_BK * 2;
  const int stage_bytes = a_bytes + b_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)num_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + num_stages;
  uint64_t* tmem_full_bar = empty_bar + num_stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GEMM_BM;
  const int n0 = blockIdx.x * BN;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int chunk_kb = p.chunk_kb > 0 ? p.chunk_kb : GEMM_CHUNK_KB;
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;
  constexpr uint32_t TMEM_COLS = 4 * BN;  // 512

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % num_stages;
        if (kb >= num_stages) mbar_wait(&empty_bar[s], ((kb / num_stages) - 1) & 1);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        const int k0 = kb * GEMM_BK;
        if constexpr (!A_MN) {
          tma_load_3d(sa, &tma_a, &full_bar[s], k0, m0, 0);  // box {64, BM, nsplit}
        } else {
#pragma unroll
          for (int i = 0; i < GEMM_BM / 64; ++i)                // box {64 (M), BK rows, nsplit}
            tma_load_3d(sa + (size_t)i * nsplit * GEMM_BK * 128, &tma_a, &full_bar[s], m0 + i * 64, k0, 0);
        }
        if constexpr (!B_MN) {
          tma_load_3d(sb, &tma_b, &full_bar[s], k0, n0, 0);  // box {64, BN, nsplit}
        } else {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tma_load_3d(sb + (size_t)i * nsplit * GEMM_BK * 128, &tma_b, &full_bar[s], n0 + i * 64, k0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int kb = 0;
      for (int c = 0; c < num_chunks; ++c) {
        const int buf = c & 1;
        if (c >= 2) {  // the fold of the chunk that used this buffer two chunks ago must be done
          mbar_wait(&tmem_empty_bar[buf], ((c >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t t_main = tmem_base + (uint32_t)(buf * BN);
        const uint32_t t_corr = tmem_base + (uint32_t)(2 * BN);
        const int kb_end = min(kb + chunk_kb, num_kb);
        uint32_t acc_main = 0u;                 // first main MMA of a chunk overwrites its buffer
        uint32_t acc_corr = (c == 0) ? 0u : 1u; // corrections accumulate across the whole contraction
        for (; kb < kb_end; ++kb) {
          const int s = kb % num_stages;
          mbar_wait(&full_bar[s], (kb / num_stages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          if (nsplit == 3) issue_kblock<3, BN, A_MN, B_MN, false>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
          else if (nsplit == 2) issue_kblock<2, BN, A_MN, B_MN, false>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
          else issue_kblock<1, BN, A_MN, B_MN, false>(sa, sb, t_main, t_corr, idesc, acc_main, acc_corr);
          umma_commit(&empty_bar[s]);  // frees the ring slot once these MMAs have read it
        }
        umma_commit(&tmem_full_bar[buf]);  // chunk accumulators complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5): promotion + fused output =====================
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
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
        if (has_corr && last_chunk) {   // the last chunk's commit also covers every correction MMA
          tmem_ld_32x32(t_lane + (uint32_t)(2 * BN + g * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[g * 32 + j] += __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    if (row < p.M && !(p.debug & 1)) {
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
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Host side -----------------------------------------------------------------------------------------------------

// A split-bf16 matrix as stored in HBM: [nplanes][rows][ld] row-major.
struct PlaneMat {
  const bf16* base;
  long rows;          // number of rows of the stored matrix
  long cols;          // number of (valid) columns
  long ld;            // row stride in elements (multiple of 8)
  long plane_stride;  // elements between planes (multiple of 8)
  int nplanes;
};

// D = A.B with A given as [M,K] (a_mn_major=0) or [K,M] (a_mn_major=1); B as [N,K] (0) or [K,N] (1).
int gemm_launch(const PlaneMat& A, int a_mn_major, const PlaneMat& B, int b_mn_major, const GemmParams& p,
                cudaStream_t stream);

// CUDA-event instrumentation of GEMM launches (bench.py roofline): see oob_gemm_timing_begin/end
int gemm_timing_begin();
int gemm_timing_end(double* total_ms, double* total_flops, long* launches);

}  // namespace oob
