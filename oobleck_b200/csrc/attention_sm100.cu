// Causal multi-head self-attention for the GPT-2 stage layer on the 5th-generation tensor cores (sm_100a):
// softmax(where(causal, QK^T/sqrt(d), finfo.min)) V in fp32-grade arithmetic, flash-style (the T x T score matrix
// never reaches HBM; the reference -- HF GPT2Attention eager, reached from oobleck/execution/layer.py:144-145 --
// materialises mb*H*T^2 fp32 per block).
//
// Operands are the split 16-bit planes the GEMMs use (common.cuh): q|k|v planes [NPL][B*T][3E] straight from the QKV
// GEMM epilogue, dO planes [NPL][B*T][E] from the proj dgrad epilogue.  NPL = 2: fp16 pairs, 3 tensor-core products
// per MAC (default); NPL = 3: bf16 x 3, 6 products.  Every product D = A.B^T or A.B is
//     main  = A_0 B_0                       -> TMEM "main" accumulator
//     corr  = A_0 B_1 + A_1 B_0 (+ ...)     -> TMEM "corr" accumulator, folded as main + CS * corr  (CS = 2^-11 | 1)
// with A a [128 x 64] K-major SWIZZLE_128B tile and B a [64 x 64] tile used either K-major (B^T: contraction over the
// 64 columns) or MN-major (contraction over the 64 rows) -- the same smem tile serves both, so dO / Q / K tiles are
// loaded once per step for the two products that need them.  Tiles arrive by TMA (cp.async.bulk.tensor.3d, all planes
// of a tile per instruction); probabilities / dS are written to shared memory by the softmax threads in the same
// swizzled K-major layout and consumed as the A operand of the next product.
//
// One thread owns one accumulator row (= TMEM lane): a softmax row never crosses threads, no shuffles.  The tensor
// core's fp32 accumulator truncates, so the forward keeps the running O in registers and folds
// each tile's fresh P.V product with round-to-nearest adds; the backward accumulators (<= 64 main MMAs) stay in TMEM.
//
//   forward        grid (H, B, T/128 q tiles)  : per 64-key tile  S = Q K^T, online softmax, O += P V
//   backward dK,dV grid (H, B, T/128 key tiles): per 64-query tile S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
//   backward dQ    grid (H, B, T/128 q tiles)  : per 64-key tile  S = Q K^T, dP = dO V^T, dQ += dS K
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2.. one (forward) or two (backward; each
// half of the columns) groups of 128 row-owning threads.
#include <type_traits>

#include "kernels.h"

namespace oob {
namespace {

constexpr int AD = 64;                     // head dim
constexpr int A_ROWS = 128;                // rows of an A tile (= MMA M = TMEM lanes)
constexpr int B_ROWS = 64;                 // rows of a B tile
constexpr int ROWB = 128;                  // bytes per tile row: 64 two-byte elements = one SWIZZLE_128B span
constexpr int A_PLANE = A_ROWS * ROWB;     // 16 KB
constexpr int B_PLANE = B_ROWS * ROWB;     // 8 KB
constexpr float LOG2E = 1.4426950408889634f;

template <int NPL>
struct Fmt {
  static_assert(NPL == 2 || NPL == 3, "fp16 pair (2 planes) or bf16 x 3");
  static constexpr bool FP16 = NPL == 2;
  static constexpr int A_TILE = NPL * A_PLANE;
  static constexpr int B_TILE = NPL * B_PLANE;
  static constexpr int NPROD = FP16 ? 3 : 6;
};
template <int NPL>
__device__ __forceinline__ constexpr float corr_scale() { return NPL == 2 ? H1_INV_SCALE : 1.0f; }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16 consecutive columns of this thread's row: main + CS * corr
template <int NPL>
__device__ __forceinline__ void ld_combined16(uint32_t t_main, uint32_t t_corr, float (&s)[16]) {
  uint32_t a[16], c[16];
  tmem_ld_x16(t_main, a);
  tmem_ld_x16(t_corr, c);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = fmaf(__uint_as_float(c[i]), corr_scale<NPL>(), __uint_as_float(a[i]));
}

// One split product D[128 x 64] (+)= A[128 x 64] . op(B[64 x 64]) issued by a single thread.
//   B_MN = false: D[m][n] = sum_k A[m][k] B[n][k]   (B K-major: its rows are the n index)
//   B_MN = true : D[m][n] = sum_k A[m][k] B[k][n]   (B MN-major: its rows are the contraction index)
// `acc` = 0 starts fresh accumulators.  The "corr" accumulator must sit right behind "main" (t_main + 64): the leading
// product A0.B0 and the first correction A0.B1 are ONE instruction with N = 128 -- planes 0 and 1 of a B tile are
// adjacent in shared memory, so [B0 | B1] is simply a B operand of 128 rows (K-major) or of two N-atoms one plane apart
// (MN-major), and its two halves land in the main and corr columns.  An N = 64 instruction occupies the tensor pipe for
// 32 clk, less than the single issuing lane needs per instruction (descriptors travel vector -> uniform registers):
// the first version, 3 x N = 64 per k-step, was issue-bound.  Descriptor offsets are compile-time constants.
template <int NPL, bool B_MN>
__device__ __forceinline__ void issue_product(uint32_t sa, uint32_t sb, uint32_t t_main, uint32_t t_corr, uint32_t acc) {
  constexpr int F16 = Fmt<NPL>::FP16 ? 0 : 1;   // instruction-descriptor format code: 0 = F16, 1 = BF16
  constexpr uint32_t idesc64 = make_idesc_f16kind(A_ROWS, B_ROWS, 0, B_MN ? 1 : 0, F16, F16);
  constexpr uint32_t idesc128 = make_idesc_f16kind(A_ROWS, 2 * B_ROWS, 0, B_MN ? 1 : 0, F16, F16);
  constexpr uint32_t B_KSTEP = B_MN ? 16 * ROWB : 32;   // 16 contraction rows, or 16 elements inside the row
  // remaining corrections after A0.[B0|B1]:  fp16 pair: A1.B0      bf16 x 3: A1.B0, A1.B1, A0.B2, A2.B0
  constexpr int NREST = Fmt<NPL>::FP16 ? 1 : 4;
  constexpr int PA[4] = {1, 1, 0, 2};
  constexpr int PB[4] = {0, 1, 2, 0};
  const uint64_t a_base = make_smem_desc(sa, 0, 8 * ROWB, SWZ_128B);
  const uint64_t b_base = make_smem_desc(sb, B_MN ? B_PLANE : 0, 8 * ROWB, SWZ_128B);
  (void)t_corr;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    umma_bf16(t_main, a_base + ((k * 32) >> 4), b_base + ((k * B_KSTEP) >> 4), idesc128, k == 0 ? acc : 1u);
#pragma unroll
    for (int q = 0; q < NREST; ++q)
      umma_bf16(t_main + 64, a_base + ((PA[q] * A_PLANE + k * 32) >> 4),
                b_base + ((PB[q] * B_PLANE + k * B_KSTEP) >> 4), idesc64, 1u);
  }
}

// Two fp32 values -> packed planes (low half = x).  The packed converts (F2FP) produce both halves in one instruction
// and the plane words need no further shuffling (the scalar version spent ~5 LOP3 / PRMT per element on packing).
__device__ __forceinline__ void split_h2_x2(float x, float y, uint32_t& h0, uint32_t& h1) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(y), "f"(x));
  float fx, fy;
  asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}"
      : "=f"(fx), "=f"(fy)
      : "r"(h0));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"((y - fy) * H1_SCALE), "f"((x - fx) * H1_SCALE));
}
__device__ __forceinline__ void split3_x2(float x, float y, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(y), "f"(x));
  float rx = x - __uint_as_float(p0 << 16), ry = y - __uint_as_float(p0 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(ry), "f"(rx));
  rx -= __uint_as_float(p1 << 16);
  ry -= __uint_as_float(p1 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(ry), "f"(rx));
}

// 8 consecutive values of row `row` (columns 8*c8 ..) -> the planes of a thread-written A tile (K-major, SWIZZLE_128B:
// the 16-byte chunk index is XORed with row & 7 -- the pattern TMA produces for the loaded tiles)
template <int NPL>
__device__ __forceinline__ void store_split8(uint32_t s_tile, int row, int c8, const float (&x)[8]) {
  const uint32_t addr = s_tile + row * ROWB + ((c8 ^ (row & 7)) << 4);
  if constexpr (Fmt<NPL>::FP16) {
    uint32_t h0[4], h1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_h2_x2(x[2 * i], x[2 * i + 1], h0[i], h1[i]);
    sts_v4(addr, h0[0], h0[1], h0[2], h0[3]);
    sts_v4(addr + A_PLANE, h1[0], h1[1], h1[2], h1[3]);
  } else {
    uint32_t w[3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split3_x2(x[2 * i], x[2 * i + 1], w[0][i], w[1][i], w[2][i]);
#pragma unroll
    for (int p = 0; p < 3; ++p) sts_v4(addr + p * A_PLANE, w[p][0], w[p][1], w[p][2], w[p][3]);
  }
}
template <int NPL>
__device__ __forceinline__ void store_zero8(uint32_t s_tile, int row, int c8) {
  const uint32_t addr = s_tile + row * ROWB + ((c8 ^ (row & 7)) << 4);
#pragma unroll
  for (int p = 0; p < NPL; ++p) sts_v4(addr + p * A_PLANE, 0u, 0u, 0u, 0u);
}

// 4 consecutive output elements: fp32 and / or split planes (plane-set codes of common.cuh)
__device__ __forceinline__ void store_quad(float* f32, uint16_t* planes, long ps, int code, long off, float4 v) {
  if (f32) *reinterpret_cast<float4*>(f32 + off) = v;
  if (!planes) return;
  const float y[4] = {v.x, v.y, v.z, v.w};
  uint16_t* pp = planes + off;
  if (code != PLANES_H2) {
    bf16 p0[4], p1[4], p2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split3(y[j], p0[j], p1[j], p2[j]);
    auto pk = [](bf16 lo, bf16 hi) { return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16); };
    *reinterpret_cast<uint2*>(pp) = make_uint2(pk(p0[0], p0[1]), pk(p0[2], p0[3]));
    if (code > 1) *reinterpret_cast<uint2*>(pp + ps) = make_uint2(pk(p1[0], p1[1]), pk(p1[2], p1[3]));
    if (code > 2) *reinterpret_cast<uint2*>(pp + 2 * ps) = make_uint2(pk(p2[0], p2[1]), pk(p2[2], p2[3]));
  }
  if (code == 5 || code == PLANES_H2) {
    uint16_t h0[4], h1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_h2(y[j], h0[j], h1[j]);
    uint16_t* ph = code == 5 ? pp + 3 * ps : pp;
    *reinterpret_cast<uint2*>(ph) = make_uint2(h0[0] | ((uint32_t)h0[1] << 16), h0[2] | ((uint32_t)h0[3] << 16));
    *reinterpret_cast<uint2*>(ph + ps) = make_uint2(h1[0] | ((uint32_t)h1[1] << 16), h1[2] | ((uint32_t)h1[3] << 16));
  }
}

// Output staging: a [128 rows x 64 fp32] tile in shared memory (32 KB, 16-byte chunks XOR-swizzled by row & 7 so the
// row-per-thread writes and the row-major reads are both conflict-free); the write-out then moves full 256-B rows.
__device__ __forceinline__ uint32_t stage_addr(uint32_t s_stage, int row, int c16) {
  return s_stage + row * 256 + ((c16 ^ (row & 7)) << 4);
}
// rows [row_base, row_base+128) x 64 columns starting at element offset col0 of a [*, ld] matrix; row validity by limit
__device__ __forceinline__ void stage_writeout(uint32_t s_stage, int tid, int nthreads, float* f32, uint16_t* planes,
                                               long ps, int code, long grow0, int rows_valid, long ld, long col0) {
  for (int idx = tid; idx < A_ROWS * 16; idx += nthreads) {
    const int row = idx >> 4, c = idx & 15;
    if (row >= rows_valid) continue;
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(stage_addr(s_stage, row, c)));
    store_quad(f32, planes, ps, code, (grow0 + row) * ld + col0 + 4 * c, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward.  TMEM (256 columns, two CTAs per SM): S main | S corr | O main | O corr, 64 columns each.
//
// Per 64-key tile j the three actors run a software pipeline (measured on the first tcgen05 version, which kept O in
// registers and folded every tile: 70 us, tensor pipe 19 %, issue slots 41 % -- every tile was one serial round trip
// MMA -> softmax -> MMA -> fold; profiles/r02_ncu_attn_fwd_v1_tcgen05.txt):
//   MMA warp : S(j+1) is issued as soon as the softmax threads have pulled S(j) into registers (SFREE), i.e. it runs
//              under the exponentials of tile j; PV(j) follows when P(j) is in shared memory.  K is double-buffered so
//              that S(j+1) never waits for its TMA.
//   softmax  : ONE pass over S (64 scores stay in registers), lazy running maximum: the reference point m_ref of a row
//              only moves when the row maximum has grown by more than 2^8 (then O, still in TMEM, is rescaled with
//              tcgen05.ld/st -- rare after the first tiles); P = exp(s - m_ref) <= 2^8 fits the fp16 pair.
//   O        : accumulated in TMEM over all key tiles (<= 64 main MMAs per row: truncation bias ~1e-6 relative, the
//              same budget the backward accumulators use), read once in the epilogue.
enum { FB_QFULL = 0, FB_KFULL0, FB_KFULL1, FB_KFREE0, FB_KFREE1, FB_VFULL, FB_VFREE, FB_SREADY, FB_SFREE, FB_PREADY,
       FB_OREADY, FB_COUNT };
constexpr float LAZY_RESCALE_LOG2 = 8.0f;

// the softmax work of one tile for one row; DIAG: the tile crosses the causal diagonal for this warp (masking needed)
template <int NPL, bool DIAG>
__device__ __forceinline__ void fwd_tile_scores(uint32_t tS_main, uint32_t tS_corr, int k0, int qi, float (&s)[64]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float t[16];
    ld_combined16<NPL>(tS_main + g * 16, tS_corr + g * 16, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) s[g * 16 + i] = (DIAG && k0 + g * 16 + i > qi) ? -INFINITY : t[i];
  }
}

template <int NPL>
__global__ void __launch_bounds__(192, NPL == 2 ? 2 : 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                      float* __restrict__ out, uint16_t* __restrict__ planes, long plane_stride, int nplanes,
                      float* __restrict__ lse, int T, int H, float scale) {
  using F = Fmt<NPL>;
  extern __shared__ __align__(1024) uint8_t smem[];   // SWIZZLE_128B tiles need 1024-B alignment (checked below)
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + F::A_TILE, sV = sK + 2 * F::B_TILE, sP = sV + F::B_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * F::A_TILE + 3 * F::B_TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + FB_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heaviest tiles (most keys under the causal mask) first: grid = (H, B, q tiles), z is the slowest index
  const int qt = (int)(gridDim.z - 1 - blockIdx.z), h = blockIdx.x, b = blockIdx.y;
  const int E = H * AD;
  const int q0 = qt * A_ROWS;
  const int kv_len = min(T, q0 + A_ROWS);
  const int nkt = (kv_len + B_ROWS - 1) / B_ROWS;
  const int grow0 = b * T;   // first row of this sequence in the [B*T, *] matrices
  constexpr uint32_t TMEM_COLS = 256;
  if ((sQ & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    for (int i = 0; i < FB_COUNT; ++i) mbar_init(&bars[i], (i == FB_PREADY || i == FB_SFREE) ? 4 : 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS_main = tmem_base, tS_corr = tmem_base + 64, tO_main = tmem_base + 128, tO_corr = tmem_base + 192;

  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int j = 0; j < nkt; ++j) {
      const int st = j & 1;
      if (j >= 2) mbar_wait(&bars[FB_KFREE0 + st], ((j >> 1) - 1) & 1);
      if (elect_one()) {
        if (j == 0) {
          mbar_arrive_expect_tx(&bars[FB_QFULL], F::A_TILE);
          tma_load_3d(smem, &tm_q, &bars[FB_QFULL], h * AD, grow0 + q0, 0);
        }
        mbar_arrive_expect_tx(&bars[FB_KFULL0 + st], F::B_TILE);
        tma_load_3d(smem + F::A_TILE + st * F::B_TILE, &tm_kv, &bars[FB_KFULL0 + st], E + h * AD, grow0 + j * B_ROWS, 0);
      }
      __syncwarp();
      if (j > 0) mbar_wait(&bars[FB_VFREE], (j - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars[FB_VFULL], F::B_TILE);
        tma_load_3d(smem + F::A_TILE + 2 * F::B_TILE, &tm_kv, &bars[FB_VFULL], 2 * E + h * AD, grow0 + j * B_ROWS, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(&bars[FB_QFULL], 0);
    mbar_wait(&bars[FB_KFULL0], 0);
    tc_fence_after();
    if (elect_one()) {
      issue_product<NPL, false>(sQ, sK, tS_main, tS_corr, 0u);
      umma_commit(&bars[FB_KFREE0]);
      umma_commit(&bars[FB_SREADY]);
    }
    __syncwarp();
    for (int j = 0; j < nkt; ++j) {
      if (j + 1 < nkt) {   // S(j+1): under the softmax of tile j
        const int st = (j + 1) & 1;
        mbar_wait(&bars[FB_KFULL0 + st], ((j + 1) >> 1) & 1);
        mbar_wait(&bars[FB_SFREE], j & 1);      // every softmax thread holds S(j) in registers
        tc_fence_after();
        if (elect_one()) {
          issue_product<NPL, false>(sQ, sK + st * F::B_TILE, tS_main, tS_corr, 0u);
          umma_commit(&bars[FB_KFREE0 + st]);
          umma_commit(&bars[FB_SREADY]);
        }
        __syncwarp();
      }
      mbar_wait(&bars[FB_PREADY], j & 1);       // P(j) is in shared memory, O is rescaled if it had to be
      mbar_wait(&bars[FB_VFULL], j & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_product<NPL, true>(sP, sV, tO_main, tO_corr, j > 0 ? 1u : 0u);
        umma_commit(&bars[FB_VFREE]);
        umma_commit(&bars[FB_OREADY]);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / output threads: one query row each =====================
    const int quarter = warp & 3;            // TMEM lane partition this warp may access
    const int r = quarter * 32 + lane;       // row inside the tile
    const int qi = q0 + r;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float sl2 = scale * LOG2E;
    float m_ref = -INFINITY, l = 0.f;        // reference point of the exponentials (raw score units), running sum
    for (int j = 0; j < nkt; ++j) {
      const int k0 = j * B_ROWS;
      mbar_wait(&bars[FB_SREADY], j & 1);
      tc_fence_after();
      const bool active = k0 <= q0 + quarter * 32 + 31;   // some row of this warp sees a key of this tile
      const bool diag = k0 + B_ROWS - 1 > q0 + quarter * 32;
      float s[64];
      if (active) {
        if (diag) fwd_tile_scores<NPL, true>(tS_main + lane_off, tS_corr + lane_off, k0, qi, s);
        else fwd_tile_scores<NPL, false>(tS_main + lane_off, tS_corr + lane_off, k0, qi, s);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[FB_SFREE]);        // the tensor core may overwrite S
      if (j > 0) {                                        // PV(j-1) done: P buffer free, O at rest
        mbar_wait(&bars[FB_OREADY], (j - 1) & 1);
        tc_fence_after();
      }
      if (active) {
        float mx = s[0];
#pragma unroll
        for (int i = 1; i < 64; ++i) mx = fmaxf(mx, s[i]);
        const bool need = (mx - m_ref) * sl2 > LAZY_RESCALE_LOG2;   // first tile: m_ref = -inf
        if (__any_sync(0xffffffffu, need)) {
          const float new_ref = need ? mx : m_ref;
          const float alpha = need ? ex2f((m_ref - new_ref) * sl2) : 1.f;
          l *= alpha;
          if (j > 0) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {   // O main (4 x 16 columns) then O corr
              uint32_t v[16];
              tmem_ld_x16(tO_main + lane_off + g * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_x16(tO_main + lane_off + g * 16, v);
            }
            tmem_st_wait();
          }
          m_ref = new_ref;
        }
        const float negm = -m_ref * sl2;
        float rowsum = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float p = ex2f(fmaf(s[c * 8 + i], sl2, negm));   // masked scores are -inf: p = 0
            s[c * 8 + i] = p;
            rowsum += p;
          }
          store_split8<NPL>(sP, r, c, reinterpret_cast<const float(&)[8]>(s[c * 8]));
        }
        l += rowsum;
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) store_zero8<NPL>(sP, r, c);   // PV accumulates every row: masked rows add zero
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[FB_PREADY]);
    }
    // every MMA has completed (OREADY of the last tile): read O, normalise, stage through the dead Q tile
    mbar_wait(&bars[FB_OREADY], (nkt - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l;
    if (qi < T) lse[((long)b * H + h) * T + qi] = m_ref * scale + __logf(l);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float o[16];
      ld_combined16<NPL>(tO_main + lane_off + g * 16, tO_corr + lane_off + g * 16, o);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t a = stage_addr(sQ, r, g * 4 + c);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(o[4 * c] * inv), "f"(o[4 * c + 1] * inv),
                     "f"(o[4 * c + 2] * inv), "f"(o[4 * c + 3] * inv)
                     : "memory");
      }
    }
    tc_fence_before();
    named_bar_sync(1, 128);
    stage_writeout(sQ, threadIdx.x - 64, 128, out, planes, plane_stride, nplanes, (long)grow0 + q0, min(A_ROWS, T - q0),
                   E, (long)h * AD);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// delta[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d]
__global__ void attention_delta_kernel(const float* __restrict__ o, const float* __restrict__ dout,
                                       float* __restrict__ delta, int B, int T, int H) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = B * T * H;
  if (warp >= total) return;
  const int h = warp % H, bt = warp / H;
  const long off = (long)bt * H * AD + h * AD + lane * 2;
  const float2 a = *reinterpret_cast<const float2*>(o + off), d = *reinterpret_cast<const float2*>(dout + off);
  const float s = warp_sum(a.x * d.x + a.y * d.y);
  if (lane == 0) delta[((long)(bt / T) * H + h) * T + (bt % T)] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// backward, dK and dV of one tile of 128 keys (rows = keys).  Per 64-query tile:
//   S^T = K Q^T, dP^T = V dO^T  ->  P^T = exp(S^T scale - lse[q]),  dS^T = P^T (dP^T - delta[q]) scale
//   dV += P^T dO,  dK += dS^T Q    (Q / dO tiles reused as MN-major B operands)
// TMEM (512 columns): S^T | dP^T | dV | dK, each main + corr of 64 columns.
// 18 warps: producer, MMA, and BWD_NCG groups of 128 row-owning threads, each group taking 64 / BWD_NCG query columns
// (4 warps per scheduler hide the TMEM-load and SFU latencies; the first version ran 2 groups, sequential MMA and
// element phases: 99 us, issue slots 34 % busy -- profiles/r02_ncu_attn_bwd_v1_tcgen05.txt).
// Pipeline (fp16 pairs, two Q / dO stages): the element threads pull S^T / dP^T of tile i into registers and release
// the TMEM regions at once (SFREE), so the tensor core computes S^T / dP^T of tile i+1 under the exponentials of tile i;
// dV / dK of tile i follow when P^T / dS^T are in shared memory.
constexpr int BWD_NCG = 4;                         // column groups
constexpr int BWD_COLS = B_ROWS / BWD_NCG;         // 16 columns per thread
constexpr int BWD_THREADS = 64 + 128 * BWD_NCG;    // 576
enum { KB_KVFULL = 0, KB_SREADY, KB_SFREE, KB_PSREADY, KB_PSFREE, KB_PTDONE, KB_DSREADY, KB_DONE, KB_QFULL0, KB_QFULL1,
       KB_QFREE0, KB_QFREE1, KB_COUNT };

template <int NPL>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kv_sm100_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                         const __grid_constant__ CUtensorMap tm_do, const float* __restrict__ lse,
                         const float* __restrict__ delta, float* __restrict__ dqkv, uint16_t* __restrict__ planes,
                         long plane_stride, int nplanes, int T, int H, float scale) {
  using F = Fmt<NPL>;
  constexpr int STAGES = NPL == 2 ? 2 : 1;         // Q / dO ring depth (smem budget)
  constexpr bool SHARED_PS = NPL == 3;             // P^T and dS^T share one buffer (bf16 x 3 tiles are 1.5x larger)
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + F::A_TILE;
  const uint32_t sQ0 = sV + F::A_TILE;                       // stage s: Q at sQ0 + s * 2 * B_TILE, dO right behind it
  const uint32_t sPT = sQ0 + STAGES * 2 * F::B_TILE;
  const uint32_t sDS = SHARED_PS ? sPT : sPT + F::A_TILE;
  const uint32_t tiles_end = sDS + F::A_TILE;
  float* sNl = reinterpret_cast<float*>(smem + (tiles_end - sK));   // [2][64]  -lse * log2(e) per query of the tile
  float* sDel = sNl + 2 * B_ROWS;                                   // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDel + 2 * B_ROWS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + KB_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.z, h = blockIdx.x, b = blockIdx.y;   // key tile 0 sees every query: heaviest first
  const int E = H * AD;
  const int k0 = kt * A_ROWS;
  const int nq64 = (T + B_ROWS - 1) / B_ROWS;
  const int it0 = k0 / B_ROWS;              // first query tile that can see these keys
  const int ntiles = nq64 - it0;
  const int grow0 = b * T;
  constexpr uint32_t TMEM_COLS = 512;
  if ((sK & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    for (int i = 0; i < KB_COUNT; ++i)
      mbar_init(&bars[i], (i == KB_PSREADY || i == KB_DSREADY || i == KB_SFREE) ? 4 * BWD_NCG : 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tST = tmem_base, tDPT = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 384;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars[KB_KVFULL], 2 * F::A_TILE);
      tma_load_3d(smem, &tm_kv, &bars[KB_KVFULL], E + h * AD, grow0 + k0, 0);
      tma_load_3d(smem + F::A_TILE, &tm_kv, &bars[KB_KVFULL], 2 * E + h * AD, grow0 + k0, 0);
    }
    __syncwarp();
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&bars[KB_QFREE0 + s], ((it / STAGES) - 1) & 1);
      if (elect_one()) {
        uint8_t* dst = smem + 2 * F::A_TILE + (size_t)s * 2 * F::B_TILE;
        const int q0 = (it0 + it) * B_ROWS;
        mbar_arrive_expect_tx(&bars[KB_QFULL0 + s], 2 * F::B_TILE);
        tma_load_3d(dst, &tm_q, &bars[KB_QFULL0 + s], h * AD, grow0 + q0, 0);
        tma_load_3d(dst + F::B_TILE, &tm_do, &bars[KB_QFULL0 + s], h * AD, grow0 + q0, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    auto issue_scores = [&](int it) {     // S^T, dP^T of tile `it` (fresh accumulators)
      const int s = it % STAGES;
      const uint32_t sQ = sQ0 + s * 2 * F::B_TILE, sdO = sQ + F::B_TILE;
      mbar_wait(&bars[KB_QFULL0 + s], (it / STAGES) & 1);
      if (it > 0) mbar_wait(&bars[KB_SFREE], (it - 1) & 1);   // the element threads hold tile it-1 in registers
      tc_fence_after();
      if (elect_one()) {
        issue_product<NPL, false>(sK, sQ, tST, tST + 64, 0u);
        issue_product<NPL, false>(sV, sdO, tDPT, tDPT + 64, 0u);
        umma_commit(&bars[KB_SREADY]);
      }
      __syncwarp();
    };
    mbar_wait(&bars[KB_KVFULL], 0);
    issue_scores(0);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % STAGES;
      const uint32_t sQ = sQ0 + s * 2 * F::B_TILE, sdO = sQ + F::B_TILE;
      if (STAGES == 2 && it + 1 < ntiles) issue_scores(it + 1);   // under the exponentials of tile `it`
      mbar_wait(&bars[KB_PSREADY], it & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_product<NPL, true>(sPT, sdO, tDV, tDV + 64, it > 0 ? 1u : 0u);
        if constexpr (SHARED_PS) umma_commit(&bars[KB_PTDONE]);
      }
      __syncwarp();
      if constexpr (SHARED_PS) {
        mbar_wait(&bars[KB_DSREADY], it & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        issue_product<NPL, true>(sDS, sQ, tDK, tDK + 64, it > 0 ? 1u : 0u);
        umma_commit(&bars[KB_QFREE0 + s]);
        umma_commit(&bars[KB_PSFREE]);
        if (it == ntiles - 1) umma_commit(&bars[KB_DONE]);
      }
      __syncwarp();
      if (STAGES == 1 && it + 1 < ntiles) issue_scores(it + 1);   // one Q / dO stage: its reload needs dK(it) done
    }
  } else {
    // ===================== P^T / dS^T threads: one key row, BWD_COLS query columns each =====================
    const int quarter = warp & 3;
    const int cg = (warp - 2) >> 2;               // query columns [BWD_COLS * cg, +BWD_COLS) of the tile
    const int r = quarter * 32 + lane;
    const int kj = k0 + r;
    const int tid = threadIdx.x - 64;             // 0 .. 128 * BWD_NCG - 1
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t col0 = (uint32_t)(cg * BWD_COLS);
    const float sl2 = scale * LOG2E;
    // per-query softmax statistics: fetched one tile ahead (the global-load latency hides under the previous tile's
    // arithmetic), handed over through shared memory double-buffered by tile parity
    float nl_next = -INFINITY, del_next = 0.f;
    auto fetch_stats = [&](int it) {
      const int qi = (it0 + it) * B_ROWS + tid;
      const bool in = tid < B_ROWS && it < ntiles && qi < T;
      nl_next = in ? -lse[((long)b * H + h) * T + qi] * LOG2E : -INFINITY;
      del_next = in ? delta[((long)b * H + h) * T + qi] : 0.f;
    };
    fetch_stats(0);
    for (int it = 0; it < ntiles; ++it) {
      const int q0 = (it0 + it) * B_ROWS;
      const int pb = it & 1;
      if (tid < B_ROWS) {
        sNl[pb * B_ROWS + tid] = nl_next;
        sDel[pb * B_ROWS + tid] = del_next;
      }
      named_bar_sync(1, 128 * BWD_NCG);
      fetch_stats(it + 1);
      mbar_wait(&bars[KB_SREADY], it & 1);
      tc_fence_after();
      float p[BWD_COLS], dp[BWD_COLS];
      ld_combined16<NPL>(tST + lane_off + col0, tST + 64 + lane_off + col0, p);
      ld_combined16<NPL>(tDPT + lane_off + col0, tDPT + 64 + lane_off + col0, dp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[KB_SFREE]);            // the tensor core may start the next tile's scores
      const bool diag = q0 < k0 + A_ROWS;     // some (key, query) pair of this tile is masked
      auto tile_body = [&](auto diag_c) {     // two instantiations: the mask costs 2 instructions per element
        constexpr bool DIAG = decltype(diag_c)::value;
#pragma unroll
        for (int i = 0; i < BWD_COLS; ++i) {
          const int ql = (int)col0 + i;
          float pv = ex2f(fmaf(p[i], sl2, sNl[pb * B_ROWS + ql]));
          if (DIAG && kj > q0 + ql) pv = 0.f;
          p[i] = pv;
          dp[i] = pv * (dp[i] - sDel[pb * B_ROWS + ql]) * scale;    // dS^T
        }
      };
      if (diag) tile_body(std::true_type{});
      else tile_body(std::false_type{});
      if (it > 0) mbar_wait(&bars[KB_PSFREE], (it - 1) & 1);   // dV / dK of the previous tile have consumed the buffers
      store_split8<NPL>(sPT, r, (int)(col0 >> 3), reinterpret_cast<const float(&)[8]>(p[0]));
      store_split8<NPL>(sPT, r, (int)(col0 >> 3) + 1, reinterpret_cast<const float(&)[8]>(p[8]));
      if constexpr (SHARED_PS) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[KB_PSREADY]);
        mbar_wait(&bars[KB_PTDONE], it & 1);   // dV has consumed P^T: the buffer may take dS^T
      }
      store_split8<NPL>(sDS, r, (int)(col0 >> 3), reinterpret_cast<const float(&)[8]>(dp[0]));
      store_split8<NPL>(sDS, r, (int)(col0 >> 3) + 1, reinterpret_cast<const float(&)[8]>(dp[8]));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[SHARED_PS ? KB_DSREADY : KB_PSREADY]);
    }
    // epilogue: dK -> columns [E, 2E), dV -> [2E, 3E) of dqkv; staged through the (dead) P^T buffer
    mbar_wait(&bars[KB_DONE], 0);
    tc_fence_after();
    const int rows_valid = min(A_ROWS, T - k0);
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t tacc = which == 0 ? tDK : tDV;
      float v[BWD_COLS];
      ld_combined16<NPL>(tacc + lane_off + col0, tacc + 64 + lane_off + col0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a = stage_addr(sPT, r, (int)(col0 >> 2) + q);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                     "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                     : "memory");
      }
      named_bar_sync(1, 128 * BWD_NCG);
      stage_writeout(sPT, tid, 128 * BWD_NCG, dqkv, planes, plane_stride, nplanes, (long)grow0 + k0, rows_valid, 3L * E,
                     (long)(which + 1) * E + h * AD);
      named_bar_sync(1, 128 * BWD_NCG);
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, dQ of one tile of 128 queries (rows = queries).  Per 64-key tile:
//   S = Q K^T, dP = dO V^T  ->  dS = exp(S scale - lse) (dP - delta) scale,   dQ += dS K   (K tile reused MN-major)
// TMEM (512 allocated): S | dP | dQ, each main + corr of 64 columns.  Same pipeline as above.
enum { QB_QFULL = 0, QB_SREADY, QB_SFREE, QB_DSREADY, QB_DSFREE, QB_DONE, QB_KVFULL0, QB_KVFULL1, QB_KVFREE0, QB_KVFREE1,
       QB_COUNT };

template <int NPL>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_q_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_kv, const float* __restrict__ lse,
                        const float* __restrict__ delta, float* __restrict__ dqkv, uint16_t* __restrict__ planes,
                        long plane_stride, int nplanes, int T, int H, float scale) {
  using F = Fmt<NPL>;
  constexpr int STAGES = NPL == 2 ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sdO = sQ + F::A_TILE;
  const uint32_t sK0 = sdO + F::A_TILE;                      // stage s: K at sK0 + s * 2 * B_TILE, V right behind it
  const uint32_t sDS = sK0 + STAGES * 2 * F::B_TILE;
  const uint32_t tiles_end = sDS + F::A_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (tiles_end - sQ));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + QB_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = (int)(gridDim.z - 1 - blockIdx.z), h = blockIdx.x, b = blockIdx.y;   // heaviest first
  const int E = H * AD;
  const int q0 = qt * A_ROWS;
  const int kv_len = min(T, q0 + A_ROWS);
  const int nkt = (kv_len + B_ROWS - 1) / B_ROWS;
  const int grow0 = b * T;
  constexpr uint32_t TMEM_COLS = 512;
  if ((sQ & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_kv);
    for (int i = 0; i < QB_COUNT; ++i) mbar_init(&bars[i], (i == QB_DSREADY || i == QB_SFREE) ? 4 * BWD_NCG : 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDQ = tmem_base + 256;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars[QB_QFULL], 2 * F::A_TILE);
      tma_load_3d(smem, &tm_q, &bars[QB_QFULL], h * AD, grow0 + q0, 0);
      tma_load_3d(smem + F::A_TILE, &tm_do, &bars[QB_QFULL], h * AD, grow0 + q0, 0);
    }
    __syncwarp();
    for (int j = 0; j < nkt; ++j) {
      const int s = j % STAGES;
      if (j >= STAGES) mbar_wait(&bars[QB_KVFREE0 + s], ((j / STAGES) - 1) & 1);
      if (elect_one()) {
        uint8_t* dst = smem + 2 * F::A_TILE + (size_t)s * 2 * F::B_TILE;
        mbar_arrive_expect_tx(&bars[QB_KVFULL0 + s], 2 * F::B_TILE);
        tma_load_3d(dst, &tm_kv, &bars[QB_KVFULL0 + s], E + h * AD, grow0 + j * B_ROWS, 0);
        tma_load_3d(dst + F::B_TILE, &tm_kv, &bars[QB_KVFULL0 + s], 2 * E + h * AD, grow0 + j * B_ROWS, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    auto issue_scores = [&](int j) {
      const int s = j % STAGES;
      const uint32_t sKt = sK0 + s * 2 * F::B_TILE, sVt = sKt + F::B_TILE;
      mbar_wait(&bars[QB_KVFULL0 + s], (j / STAGES) & 1);
      if (j > 0) mbar_wait(&bars[QB_SFREE], (j - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_product<NPL, false>(sQ, sKt, tS, tS + 64, 0u);
        issue_product<NPL, false>(sdO, sVt, tDP, tDP + 64, 0u);
        umma_commit(&bars[QB_SREADY]);
      }
      __syncwarp();
    };
    mbar_wait(&bars[QB_QFULL], 0);
    issue_scores(0);
    for (int j = 0; j < nkt; ++j) {
      const int s = j % STAGES;
      const uint32_t sKt = sK0 + s * 2 * F::B_TILE;
      if (STAGES == 2 && j + 1 < nkt) issue_scores(j + 1);
      mbar_wait(&bars[QB_DSREADY], j & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_product<NPL, true>(sDS, sKt, tDQ, tDQ + 64, j > 0 ? 1u : 0u);
        umma_commit(&bars[QB_KVFREE0 + s]);
        umma_commit(&bars[QB_DSFREE]);
        if (j == nkt - 1) umma_commit(&bars[QB_DONE]);
      }
      __syncwarp();
      if (STAGES == 1 && j + 1 < nkt) issue_scores(j + 1);
    }
  } else {
    const int quarter = warp & 3;
    const int cg = (warp - 2) >> 2;               // key columns [BWD_COLS * cg, +BWD_COLS) of the tile
    const int r = quarter * 32 + lane;
    const int qi = q0 + r;
    const int tid = threadIdx.x - 64;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t col0 = (uint32_t)(cg * BWD_COLS);
    const float sl2 = scale * LOG2E;
    const float nl = qi < T ? -lse[((long)b * H + h) * T + qi] * LOG2E : -INFINITY;
    const float del = qi < T ? delta[((long)b * H + h) * T + qi] : 0.f;
    for (int j = 0; j < nkt; ++j) {
      const int k0 = j * B_ROWS;
      mbar_wait(&bars[QB_SREADY], j & 1);
      tc_fence_after();
      float p[BWD_COLS], dp[BWD_COLS];
      ld_combined16<NPL>(tS + lane_off + col0, tS + 64 + lane_off + col0, p);
      ld_combined16<NPL>(tDP + lane_off + col0, tDP + 64 + lane_off + col0, dp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[QB_SFREE]);
      const bool diag = k0 + B_ROWS - 1 > q0;
      auto tile_body = [&](auto diag_c) {
        constexpr bool DIAG = decltype(diag_c)::value;
#pragma unroll
        for (int i = 0; i < BWD_COLS; ++i) {
          float pv = ex2f(fmaf(p[i], sl2, nl));
          if (DIAG && k0 + (int)col0 + i > qi) pv = 0.f;
          p[i] = pv * (dp[i] - del) * scale;
        }
      };
      if (diag) tile_body(std::true_type{});
      else tile_body(std::false_type{});
      if (j > 0) mbar_wait(&bars[QB_DSFREE], (j - 1) & 1);     // dQ of the previous tile has consumed the dS buffer
      store_split8<NPL>(sDS, r, (int)(col0 >> 3), reinterpret_cast<const float(&)[8]>(p[0]));
      store_split8<NPL>(sDS, r, (int)(col0 >> 3) + 1, reinterpret_cast<const float(&)[8]>(p[8]));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[QB_DSREADY]);
    }
    mbar_wait(&bars[QB_DONE], 0);
    tc_fence_after();
    {
      float v[BWD_COLS];
      ld_combined16<NPL>(tDQ + lane_off + col0, tDQ + 64 + lane_off + col0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a = stage_addr(sDS, r, (int)(col0 >> 2) + q);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                     "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                     : "memory");
      }
    }
    tc_fence_before();
    named_bar_sync(1, 128 * BWD_NCG);
    stage_writeout(sDS, tid, 128 * BWD_NCG, dqkv, planes, plane_stride, nplanes, (long)grow0 + q0, min(A_ROWS, T - q0),
                   3L * E, (long)h * AD);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
int set_smem(const void* fn, size_t bytes) {
  OOB_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int NPL>
int attention_fwd_t(const bf16* qkv, long qkv_ps, float* out, bf16* out_planes, long plane_stride, int nplanes,
                    float* lse, int B, int T, int H, cudaStream_t s) {
  using F = Fmt<NPL>;
  const long M = (long)B * T, E = (long)H * AD;
  const PlaneMat qm{qkv, M, 3 * E, 3 * E, qkv_ps, NPL, F::FP16 ? 1 : 0};
  const CUtensorMap *tq, *tkv;
  int rc;
  if ((rc = tensor_map_3d(&tq, qm, A_ROWS, NPL, 64))) return rc;
  if ((rc = tensor_map_3d(&tkv, qm, B_ROWS, NPL, 64))) return rc;
  const size_t smem = 2 * F::A_TILE + 3 * F::B_TILE + 128;
  static bool once = false;
  if (!once) {
    if (set_smem((const void*)attn_fwd_sm100_kernel<NPL>, smem)) return -1;
    once = true;
  }
  dim3 grid(H, B, (T + A_ROWS - 1) / A_ROWS);
  attn_fwd_sm100_kernel<NPL><<<grid, 192, smem, s>>>(*tq, *tkv, out, reinterpret_cast<uint16_t*>(out_planes),
                                                     plane_stride, nplanes, lse, T, H, 1.0f / sqrtf((float)AD));
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

template <int NPL>
int attention_bwd_t(const bf16* qkv, long qkv_ps, const float* out, const float* dout, const bf16* dout_planes,
                    long do_ps, const float* lse, float* delta, float* dqkv, bf16* dqkv_planes, long plane_stride,
                    int nplanes, int B, int T, int H, cudaStream_t s) {
  using F = Fmt<NPL>;
  constexpr int STAGES = NPL == 2 ? 2 : 1;
  const long M = (long)B * T, E = (long)H * AD;
  const PlaneMat qm{qkv, M, 3 * E, 3 * E, qkv_ps, NPL, F::FP16 ? 1 : 0};
  const PlaneMat dm{dout_planes, M, E, E, do_ps, NPL, F::FP16 ? 1 : 0};
  const CUtensorMap *tq128, *tq64, *tdo128, *tdo64;
  int rc;
  if ((rc = tensor_map_3d(&tq128, qm, A_ROWS, NPL, 64))) return rc;
  if ((rc = tensor_map_3d(&tq64, qm, B_ROWS, NPL, 64))) return rc;
  if ((rc = tensor_map_3d(&tdo128, dm, A_ROWS, NPL, 64))) return rc;
  if ((rc = tensor_map_3d(&tdo64, dm, B_ROWS, NPL, 64))) return rc;
  const size_t smem_kv = 2 * F::A_TILE + STAGES * 2 * F::B_TILE + (NPL == 3 ? 1 : 2) * F::A_TILE + 1024 + 256;
  const size_t smem_q = 3 * F::A_TILE + STAGES * 2 * F::B_TILE + 256;
  static bool once = false;
  if (!once) {
    if (set_smem((const void*)attn_bwd_kv_sm100_kernel<NPL>, smem_kv)) return -1;
    if (set_smem((const void*)attn_bwd_q_sm100_kernel<NPL>, smem_q)) return -1;
    once = true;
  }
  const int total = B * T * H;
  attention_delta_kernel<<<(total + 7) / 8, 256, 0, s>>>(out, dout, delta, B, T, H);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  dim3 grid(H, B, (T + A_ROWS - 1) / A_ROWS);
  const float scale = 1.0f / sqrtf((float)AD);
  uint16_t* pl = reinterpret_cast<uint16_t*>(dqkv_planes);
  attn_bwd_kv_sm100_kernel<NPL><<<grid, BWD_THREADS, smem_kv, s>>>(*tq128, *tq64, *tdo64, lse, delta, dqkv, pl, plane_stride,
                                                           nplanes, T, H, scale);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  attn_bwd_q_sm100_kernel<NPL><<<grid, BWD_THREADS, smem_q, s>>>(*tq128, *tdo128, *tq64, lse, delta, dqkv, pl, plane_stride,
                                                         nplanes, T, H, scale);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

int attention_fwd(const bf16* qkv_planes, long qkv_plane_stride, int operand_fp16, float* out, bf16* out_planes,
                  long plane_stride, int nplanes, float* lse, int B, int T, int H, int D, cudaStream_t s) {
  OOB_CHECK(D == AD, "attention: head_dim must be 64 (got %d)", D);
  OOB_CHECK(B > 0 && T > 0 && H > 0, "attention: empty problem");
  return operand_fp16 ? attention_fwd_t<2>(qkv_planes, qkv_plane_stride, out, out_planes, plane_stride, nplanes, lse, B,
                                           T, H, s)
                      : attention_fwd_t<3>(qkv_planes, qkv_plane_stride, out, out_planes, plane_stride, nplanes, lse, B,
                                           T, H, s);
}

int attention_bwd(const bf16* qkv_planes, long qkv_plane_stride, int operand_fp16, const float* out, const float* dout,
                  const bf16* dout_planes, long dout_plane_stride, const float* lse, float* delta, float* dqkv,
                  bf16* dqkv_planes, long plane_stride, int nplanes, int B, int T, int H, int D, cudaStream_t s) {
  OOB_CHECK(D == AD, "attention: head_dim must be 64 (got %d)", D);
  OOB_CHECK(B > 0 && T > 0 && H > 0, "attention: empty problem");
  return operand_fp16 ? attention_bwd_t<2>(qkv_planes, qkv_plane_stride, out, dout, dout_planes, dout_plane_stride, lse,
                                           delta, dqkv, dqkv_planes, plane_stride, nplanes, B, T, H, s)
                      : attention_bwd_t<3>(qkv_planes, qkv_plane_stride, out, dout, dout_planes, dout_plane_stride, lse,
                                           delta, dqkv, dqkv_planes, plane_stride, nplanes, B, T, H, s);
}

}  // namespace oob
