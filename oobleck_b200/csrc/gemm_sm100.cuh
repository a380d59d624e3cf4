// tcgen05 / TMA / TMEM GEMM for the per-stage transformer layers (sm_100a).
//
//   D[M,N] (fp32) = sum over split products of A_p[M,K] * B_q[K,N]      (+ fused epilogue)
//
// Operands are split 16-bit planes (common.cuh: fp16 pairs = 3 products per MAC, the default; bf16 x 3 = 6 products,
// any range) so that the tensor cores reproduce the reference's fp32 Conv1D / Linear arithmetic
// (oobleck/execution/layer.py:104-105 pins fp32; HF GPT-2 addmm).
// Either operand may be K-major (contraction index contiguous) or MN-major (output index contiguous), which is
// what lets forward (X.W), dgrad (dY.W^T) and wgrad (X^T.dY) all read the SAME natural row-major buffers
// without a transpose pass:
//     forward  Y  = X  . W      A = X  [M,K] K-major    B = W  [K,N] N-major
//     dgrad    dX = dY . W^T    A = dY [M,N] K-major    B = W  [K,N] as [N_out=K][contract=N] K-major
//     wgrad    dW = X^T. dY     A = X  [tok,K] M-major  B = dY [tok,N] N-major
//
// Structure (gemm_sm100_persistent.cuh: a CTA pair walks 256 x BN output tiles; 192 threads per CTA):
//   warp 0    TMA producer  : cp.async.bulk.tensor.3d of all planes of the A and B k-block into a
//                             multi-stage SWIZZLE_128B shared-memory ring, mbarrier complete_tx
//   warp 1    MMA issuer    : one elected thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 in TMEM),
//                             tcgen05.commit releases ring slots / signals the epilogue; owns TMEM alloc
//   warps 2-5 epilogue      : tcgen05.ld 32x32b from TMEM + fp32 promotion, then the slab-transposed fused epilogue
//                             below (bias / residual / GELU / dGELU / accumulate, fp32 and/or plane stores)
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
  int nplanes_out;      // plane-set code (common.cuh): 1..3 bf16 planes, 5 = bf16 x 3 + fp16 pair, PLANES_H2 = pair only
  float alpha;          // value = alpha * acc (before bias etc.)
  int vec4;             // set by gemm_launch: every pointer / stride above allows 16-B (fp32) and 8-B (plane) accesses
  int prefetch;         // set by gemm_launch: L1-prefetch the residual / aux / accumulate rows of a slab in one burst
};

struct GemmParams {
  int M, N, K;
  int nsplit;   // planes used from each operand (1..3)
  GemmEpilogue epi;
  int a_bf16, b_bf16;   // operand element formats (1 = bf16 planes, 0 = fp16 planes); set by gemm_launch
  float corr_scale;     // weight of the correction accumulator (2^-11 for fp16 x 2 planes, else 1); gemm_launch
  int slab;     // epilogue transposition slab: 32 columns when the smem budget allows, else 16; set by the launcher
  int chunk_kb; // promotion chunk in k-blocks (0 = default GEMM_CHUNK_KB)
  int debug;    // diagnostics (tools/gemm_sweep.py, persistent kernel): 1 skip output stores, 2 skip TMA, 4 skip MMAs
};

constexpr int GEMM_BM = 128;
// warp 0 producer, warp 1 MMA, then GEMM_EPI_GROUPS x 4 epilogue warps: each group of 4 covers the 128 TMEM lanes
// (a warp may only touch lanes 32 * (warp % 4) ..) and takes BN / GEMM_EPI_GROUPS of the tile's columns.  Round 1 ran one
// group: 1.5 eligible warps per scheduler, "No Eligible" 76 % of the cycles, the fused epilogues cost as much as the
// whole main loop (profiles/r01_ncu_gemm_fwdfc_v5.txt, r01_gemm_sweep12_compact_epilogue.log).
constexpr int GEMM_EPI_GROUPS = 2;
constexpr int GEMM_THREADS = 64 + 128 * GEMM_EPI_GROUPS;

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

// Fused epilogue for a slab of SLAB consecutive columns of the warp's 32 output rows.
//
// After tcgen05.ld each thread holds ONE row (TMEM lane) -- storing from that layout makes every warp store touch 32
// different 128-B lines with 16 B each (measured: the forward-FC epilogue, 14 B/element, ran at ~1 TB/s and took
// longer than the whole 3-product main loop; profiles/r01_gemm_sweep8_fp16.log).  So the slab is first transposed
// through a small XOR-swizzled shared-memory tile (conflict-free both ways): afterwards SLAB/4 adjacent lanes own 4
// consecutive columns each of the same row, and every global access of the epilogue -- bias / residual / dGELU aux /
// accumulate loads, fp32 stores, plane stores -- is a run of full 32-B sectors (128-B lines for SLAB = 32).
template <int SLAB>
__device__ __forceinline__ uint32_t slab_swz(int row) {
  return SLAB == 32 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
}

__device__ __forceinline__ uint32_t pack_bf16x2(bf16 lo, bf16 hi) {
  return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// Second half of a slab: reads the transposed tile back and does all global traffic.  Deliberately NOT inlined and not
// unrolled: it is called once per slab from the (necessarily unrolled, the accumulators are registers) slab loop, and
// the first version that inlined it grew the kernel to 123k SASS instructions -- instruction-fetch bound, every GEMM
// 60 % slower (profiles/r01_gemm_sweep9_coalesced_epilogue.log).
template <int SLAB>
__device__ __noinline__ void epilogue_rows(const float* stage, const GemmEpilogue& ep, int row0, int col0, int M, int N,
                                           int lane) {
  constexpr int C = SLAB / 4;     // 16-B chunks per row
  constexpr int RPI = 32 / C;     // rows covered by one warp-wide access
  // The descriptor lives in param space behind a generic reference: copy what the loop needs into registers once, or
  // every field is re-loaded after every global store (possible aliasing) and the loop becomes one dependent chain.
  float* const d = ep.d;
  const float* const bias = ep.bias;
  const float* const resid = ep.resid;
  const float* const aux = ep.aux;
  uint16_t* const planes = reinterpret_cast<uint16_t*>(ep.planes);
  const long ldd = ep.ldd, ldr = ep.ldr, ldaux = ep.ldaux, ldp = ep.ldp, pstride = ep.plane_stride;
  const int accumulate = ep.accumulate, act = ep.act, nplanes_out = ep.nplanes_out;
  const float alpha = ep.alpha;

  const int cl = lane % C, rl = lane / C;
  const int gcol = col0 + 4 * cl;
  if (gcol >= N) return;
  const bool vec = ep.vec4 && (gcol + 4 <= N);
  if (!vec) {   // ragged / unaligned edge: plain per-element code, correctness only
#pragma unroll 1
    for (int q = 0; q < C; ++q) {
      const int rr = q * RPI + rl;
      const long grow = row0 + rr;
      if (grow >= M) break;
      const float* src = stage + rr * SLAB + 4 * (cl ^ slab_swz<SLAB>(rr));
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        if (gcol + j >= N) break;
        float y = src[j] * alpha;
        if (bias) y += bias[gcol + j];
        if (resid) y += resid[grow * ldr + gcol + j];
        if (accumulate) y += d[grow * ldd + gcol + j];
        if (act == ACT_DGELU) y *= gelu_new_grad_f(aux[grow * ldaux + gcol + j]);
        if (d) d[grow * ldd + gcol + j] = y;
        if (planes) {
          if (act == ACT_GELU) y = gelu_new_f(y);
          uint16_t w[5];
          split5(y, w[0], w[1], w[2], w[3], w[4]);
#pragma unroll
          for (int pl = 0; pl < 5; ++pl)
            if (pl < planes_count(nplanes_out))
              planes[pl * pstride + grow * ldp + gcol + j] = (nplanes_out == PLANES_H2 && pl < 2) ? w[3 + (pl & 1)] : w[pl];
        }
      }
    }
    return;
  }
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) b = ldg_f4(bias + gcol);
  const bool has_r = resid != nullptr, has_acc = accumulate != 0, has_aux = act == ACT_DGELU;
  // software pipeline, depth 1: the global loads of row-chunk q+1 are in flight while q is computed and stored (one
  // epilogue warp per scheduler: nobody else hides the latency).  The loop body is NOT unrolled -- the whole function
  // must stay resident in the instruction cache next to the MMA-issue and producer loops.
  float4 r_n = make_float4(0.f, 0.f, 0.f, 0.f), a_n = r_n, acc_n = r_n;
  if (ep.prefetch && (has_r || has_acc || has_aux)) {   // burst of L1 prefetches for every row-chunk of this slab: one exposed latency
#pragma unroll 1
    for (int q = 0; q < C; ++q) {
      const long grow = row0 + q * RPI + rl;
      if (grow >= M) break;
      if (has_r) prefetch_l1(resid + grow * ldr + gcol);
      if (has_acc) prefetch_l1(d + grow * ldd + gcol);
      if (has_aux) prefetch_l1(aux + grow * ldaux + gcol);
    }
  }
  {
    const long grow = row0 + rl;
    if (grow < M) {
      if (has_r) r_n = ldg_f4(resid + grow * ldr + gcol);
      if (has_acc) acc_n = ldg_f4(d + grow * ldd + gcol);
      if (has_aux) a_n = ldg_f4(aux + grow * ldaux + gcol);
    }
  }
#pragma unroll 1
  for (int q = 0; q < C; ++q) {
    const int rr = q * RPI + rl;
    const long grow = row0 + rr;
    if (grow >= M) break;
    const float4 v = reinterpret_cast<const float4*>(stage + rr * SLAB)[cl ^ slab_swz<SLAB>(rr)];
    const float4 r = r_n, a = a_n, acc = acc_n;
    if (q + 1 < C && grow + RPI < M) {
      const long gn = grow + RPI;
      if (has_r) r_n = ldg_f4(resid + gn * ldr + gcol);
      if (has_acc) acc_n = ldg_f4(d + gn * ldd + gcol);
      if (has_aux) a_n = ldg_f4(aux + gn * ldaux + gcol);
    }
    float y[4] = {v.x * alpha + b.x + r.x + acc.x, v.y * alpha + b.y + r.y + acc.y, v.z * alpha + b.z + r.z + acc.z,
                  v.w * alpha + b.w + r.w + acc.w};
    if (has_aux) {
      y[0] *= gelu_new_grad_f(a.x); y[1] *= gelu_new_grad_f(a.y); y[2] *= gelu_new_grad_f(a.z); y[3] *= gelu_new_grad_f(a.w);
    }
    if (d) stg_f4(d + grow * ldd + gcol, make_float4(y[0], y[1], y[2], y[3]));
    if (planes) {
      if (act == ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = gelu_new_f(y[j]);
      }
      uint16_t* pp = planes + grow * ldp + gcol;
      if (nplanes_out <= 3 || nplanes_out == 5) {
        bf16 p0[4], p1[4], p2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split3(y[j], p0[j], p1[j], p2[j]);
        *reinterpret_cast<uint2*>(pp) = make_uint2(pack_bf16x2(p0[0], p0[1]), pack_bf16x2(p0[2], p0[3]));
        if (nplanes_out > 1)
          *reinterpret_cast<uint2*>(pp + pstride) = make_uint2(pack_bf16x2(p1[0], p1[1]), pack_bf16x2(p1[2], p1[3]));
        if (nplanes_out > 2)
          *reinterpret_cast<uint2*>(pp + 2 * pstride) = make_uint2(pack_bf16x2(p2[0], p2[1]), pack_bf16x2(p2[2], p2[3]));
      }
      if (nplanes_out == 5 || nplanes_out == PLANES_H2) {
        uint16_t h0[4], h1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_h2(y[j], h0[j], h1[j]);
        uint16_t* ph = nplanes_out == 5 ? pp + 3 * pstride : pp;
        *reinterpret_cast<uint2*>(ph) =
            make_uint2((uint32_t)h0[0] | ((uint32_t)h0[1] << 16), (uint32_t)h0[2] | ((uint32_t)h0[3] << 16));
        *reinterpret_cast<uint2*>(ph + pstride) =
            make_uint2((uint32_t)h1[0] | ((uint32_t)h1[1] << 16), (uint32_t)h1[2] | ((uint32_t)h1[3] << 16));
      }
    }
  }
}

// First half: each thread drops its row's SLAB accumulator values into the swizzled tile.
template <int SLAB>
__device__ __forceinline__ void epilogue_slab(const float* x, float* stage, const GemmEpilogue& e, int row0, int col0,
                                              int M, int N, int lane) {
  constexpr int C = SLAB / 4;
  {
    float4* dst = reinterpret_cast<float4*>(stage + lane * SLAB);
    const uint32_t sw = slab_swz<SLAB>(lane);
#pragma unroll
    for (int k = 0; k < C; ++k) dst[k ^ sw] = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
  }
  __syncwarp();
  epilogue_rows<SLAB>(stage, e, row0, col0, M, N, lane);
  __syncwarp();
}

// Accumulation scheme.  The tensor core's fp32 accumulator TRUNCATES on every accumulate (measured on B200: the
// error of a K=1600 six-product GEMM grew linearly with the number of MMAs, ~2^-25.7 per MMA, and did not depend
// on the split level).  To keep fp32-grade results for any K:
//   * the leading product p0q0 ("main") is accumulated in chunks of GEMM_CHUNK_KB k-blocks; after every chunk the
//     epilogue warps fold the chunk into fp32 registers with round-to-nearest adds (promotion).  Two TMEM buffers
//     let the MMA warp run one chunk ahead of the fold;
//   * the correction products ("corr", 2^-8 of main and smaller) accumulate over the whole contraction in a third
//     TMEM region -- their truncation error is 2^-8 smaller still -- and are folded once at the end.
// TMEM: main[2] + corr[2] (one per tile parity of the persistent loop) = 4 x BN = 512 columns.  (First version folded main+corr every 4 k-blocks: the extra TMEM reads
// cost 35% of the GEMM's throughput -- profiles/README.md.)
constexpr int GEMM_CHUNK_ELEMS = 512;   // contraction elements per promotion chunk = 32 main MMAs

// Host side -----------------------------------------------------------------------------------------------------

// A split matrix as stored in HBM: [nplanes][rows][ld] row-major, 2-byte elements.
struct PlaneMat {
  const bf16* base;
  long rows;          // number of rows of the stored matrix
  long cols;          // number of (valid) columns
  long ld;            // row stride in elements (multiple of 8)
  long plane_stride;  // elements between planes (multiple of 8)
  int nplanes;
  int fp16 = 0;       // 0: bf16 planes (x = p0+p1+p2), 1: fp16 planes (x = h0 + 2^-11 h1), common.cuh
};

// D = A.B with A given as [M,K] (a_mn_major=0) or [K,M] (a_mn_major=1); B as [N,K] (0) or [K,N] (1).
int gemm_launch(const PlaneMat& A, int a_mn_major, const PlaneMat& B, int b_mn_major, const GemmParams& p,
                cudaStream_t stream);

// Cached TMA descriptor over a split matrix [nplanes][rows][ld]: box = {bk columns, box_rows, box_planes}, bk = 64 ->
// SWIZZLE_128B, 32 -> SWIZZLE_64B.  The returned pointer is valid until the next call on this host thread.
int tensor_map_3d(const CUtensorMap** out, const PlaneMat& a, int box_rows, int box_planes, int bk);
long tensor_map_encodes();   // cuTensorMapEncodeTiled calls so far (cache misses)

// CUDA-event instrumentation of GEMM launches (bench.py roofline): see oob_gemm_timing_begin/end
int gemm_timing_begin();
int gemm_timing_end(double* total_ms, double* total_flops, double* executed_flops, long* launches);

}  // namespace oob
