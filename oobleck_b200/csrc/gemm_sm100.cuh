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
  int debug;    // diagnostics (tools/gemm_sweep.py, persistent kernel): 1 skip output stores, 2 skip TMA, 4 skip MMAs
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 192;

// K-block geometry.  One smem row holds BK contraction elements (K-major) or BK output elements (MN-major) and is
// exactly one swizzle span: BK = 64 -> 128 B rows / SWIZZLE_128B, BK = 32 -> 64 B rows / SWIZZLE_64B (half-size
// stages, twice as many of them: the 6-product parity mode needs the deeper ring to hide the TMA latency).
template <int BK>
struct KCfg {
  static_assert(BK == 64 || BK == 32, "BK must be 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)");
  static constexpr int ROWB = BK * 2;
  static constexpr uint64_t SWZ = BK == 64 ? SWZ_128B : SWZ_64B;
  static constexpr int SBO = 8 * ROWB;     // stride between 8-row groups
  static constexpr int ATOM = BK;          // width (elements) of an MN-major atom = ROWB / 2
  static constexpr int KSTEPS = BK / 16;   // MMAs (K = 16) per product per k-block
};

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

template <int NS, int BROWS, bool A_MN, bool B_MN, bool TWO_CTA, int BK>
__device__ __forceinline__ void issue_kblock(uint32_t sa, uint32_t sb, uint32_t t_main, uint32_t t_corr,
                                             uint32_t idesc, uint32_t& acc_main, uint32_t& acc_corr) {
  using C = KCfg<BK>;
  // K-major : [plane][rows][ROWB]              , k-step = 32 B inside the row
  // MN-major: [atom][plane][BK k-rows][ROWB]   , k-step = 16 rows, LBO = distance between atoms
  constexpr uint32_t A_PLANE = (A_MN ? BK : GEMM_BM) * C::ROWB;
  constexpr uint32_t B_PLANE = (B_MN ? BK : BROWS) * C::ROWB;
  constexpr uint32_t A_KSTEP = A_MN ? 16 * C::ROWB : 32;
  constexpr uint32_t B_KSTEP = B_MN ? 16 * C::ROWB : 32;
  constexpr uint32_t A_LBO = A_MN ? NS * BK * C::ROWB : 0;
  constexpr uint32_t B_LBO = B_MN ? NS * BK * C::ROWB : 0;
  constexpr int NPROD = NS == 1 ? 1 : (NS == 2 ? 3 : 6);
  constexpr int PA[6] = {0, 0, 1, 1, 0, 2};
  constexpr int PB[6] = {0, 1, 0, 1, 2, 0};
  const uint64_t a_base = make_smem_desc(sa, A_LBO, C::SBO, C::SWZ);
  const uint64_t b_base = make_smem_desc(sb, B_LBO, C::SBO, C::SWZ);
#pragma unroll
  for (int k = 0; k < C::KSTEPS; ++k) {   // leading product p0q0 -> "main"
    umma_any<TWO_CTA>(t_main, a_base + ((k * A_KSTEP) >> 4), b_base + ((k * B_KSTEP) >> 4), idesc, acc_main);
    acc_main = 1u;
  }
#pragma unroll
  for (int k = 0; k < C::KSTEPS; ++k) {   // corrections -> "corr"
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
constexpr int GEMM_CHUNK_ELEMS = 512;   // contraction elements per promotion chunk = 32 main MMAs

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
