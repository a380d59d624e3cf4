// Causal multi-head self-attention for the GPT-2 stage layer (HF GPT2Attention eager path:
// softmax(where(causal, QK^T/sqrt(d), finfo.min)) V, fp32), flash-style: the T x T score matrix never
// reaches HBM (the reference materialises mb*H*T^2 fp32 = 210 MB per GPT-2-XL block).
//
// Operands arrive as the same split-bf16 planes the GEMMs use (q|k|v planes straight from the QKV GEMM epilogue, dO
// planes from the proj dgrad epilogue), so tiles are cp.async'ed into shared memory without any ALU work and fed to
// mma.sync.m16n8k16 bf16 through ldmatrix.  Each fp32 product is rebuilt from the six plane products
// (p0q0 p0q1 p1q0 p1q1 p0q2 p2q0) like in the GEMM; probabilities / dS are split into planes in registers.
// (First version: TF32x3 with per-fragment splitting of fp32 smem tiles -- 3.3x more issue slots per MMA; see
// profiles/README.md.)  64 x 64 tiles, head_dim 64, 4 warps x 16 rows; smem tiles are [plane][64][64] bf16 with the
// 16-byte chunks of a row XOR-swizzled by (row & 7), which makes every ldmatrix phase conflict-free.
// The HMMA accumulator truncates, so each tile product goes to a fresh accumulator that is folded with RN adds.
//
//   forward        grid (T/64, H, B): S = QK^T, online softmax, O = PV; writes O (fp32 + split planes) and LSE
//   backward dK,dV grid (T/64 kv tiles, H, B): S^T = KQ^T, dV += P^T dO, dP^T = V dO^T, dK += dS^T Q
//   backward dQ    grid (T/64 q tiles, H, B):  S = QK^T, dP = dO V^T, dQ += dS K
#include <cuda_bf16.h>

#include "kernels.h"

namespace oob {

constexpr int AT = 64;                     // tile (queries or keys)
constexpr int AD = 64;                     // head dim
constexpr int APLANE = AT * AD * 2;        // bytes of one plane of a tile (8 KB)
// Operand formats (template parameter H2 everywhere below):
//   H2 = false  bf16 x 3 planes, six products per MAC (any range)
//   H2 = true   fp16 pair x ~= h0 + 2^-11 h1 (common.cuh): q|k|v, probabilities, and loss-scaled gradients; three
//               products per MAC.  The two correction products go first into the (fresh) accumulator, which is then
//               multiplied by 2^-11 in place before the leading product is added -- no second accumulator.
template <bool H2> constexpr int NPL = H2 ? 2 : 3;                 // planes per tile
template <bool H2> constexpr int ATILE = NPL<H2> * APLANE;         // bytes of a tile (16 / 24 KB)

__device__ __forceinline__ uint32_t tile_off(int plane, int row, int col) {  // col: element index, multiple of 8
  return (uint32_t)(plane * APLANE + row * 128 + ((((col >> 3) ^ row) & 7) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// tile rows [row0, row0+64) x 64 columns of every plane: global planes layout [3][rows][ld] (plane stride ps elements)
template <bool H2>
__device__ __forceinline__ void load_tile_async(uint32_t sbase, const bf16* g, long ld, long ps, int valid_rows) {
  for (int i = threadIdx.x; i < NPL<H2> * AT * 8; i += blockDim.x) {
    const int plane = i >> 9, r = (i >> 3) & 63, c = i & 7;
    const uint32_t dst = sbase + tile_off(plane, r, c * 8);
    if (r < valid_rows) cp_async16(dst, g + (long)plane * ps + (long)r * ld + c * 8);
    else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 values -> packed fp16 pair planes (low half = x): h0 and the 2^11-scaled residual
__device__ __forceinline__ void split_pack_h2(float x, float y, uint32_t& h0, uint32_t& h1) {
  uint16_t a0, a1, b0, b1;
  split_h2(x, a0, a1);
  split_h2(y, b0, b1);
  h0 = (uint32_t)a0 | ((uint32_t)b0 << 16);
  h1 = (uint32_t)a1 | ((uint32_t)b1 << 16);
}
__device__ __forceinline__ void scale_acc(float (&a)[8][4], float f) {
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j][0] *= f; a[j][1] *= f; a[j][2] *= f; a[j][3] *= f; }
}
// c += a*b rebuilt from the six plane products, smallest first.  b[p] holds the fragments of two adjacent n-tiles:
// {b0, b1 of tile 2jp, b0, b1 of tile 2jp+1}; jj selects the tile.
__device__ __forceinline__ void mma6(float (&c)[4], const uint32_t (&a)[3][4], const uint32_t (&b)[3][4], int jj) {
  mma_bf16(c, a[0], b[2][2 * jj], b[2][2 * jj + 1]);
  mma_bf16(c, a[2], b[0][2 * jj], b[0][2 * jj + 1]);
  mma_bf16(c, a[1], b[1][2 * jj], b[1][2 * jj + 1]);
  mma_bf16(c, a[0], b[1][2 * jj], b[1][2 * jj + 1]);
  mma_bf16(c, a[1], b[0][2 * jj], b[0][2 * jj + 1]);
  mma_bf16(c, a[0], b[0][2 * jj], b[0][2 * jj + 1]);
}

// two fp32 values -> one packed bf16x2 per plane (low half = x)
__device__ __forceinline__ void split_pack(float x, float y, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(x, y);
  float2 f = __bfloat1622float2(h0);
  float rx = x - f.x, ry = y - f.y;
  __nv_bfloat162 h1 = __floats2bfloat162_rn(rx, ry);
  f = __bfloat1622float2(h1);
  rx -= f.x; ry -= f.y;
  __nv_bfloat162 h2 = __floats2bfloat162_rn(rx, ry);
  p0 = *reinterpret_cast<uint32_t*>(&h0);
  p1 = *reinterpret_cast<uint32_t*>(&h1);
  p2 = *reinterpret_cast<uint32_t*>(&h2);
}

// acc[16 x 64] = A[r0..r0+16, 0..64] . B^T   with A = rows of smem tile `sa`, B[k][n] = tile_b[n][k].
// `acc` must be zero on entry (every caller builds one tile product in a fresh accumulator -- the HMMA accumulator
// truncates, and the fp16-pair path rescales the accumulator between its two passes).
template <bool H2>
__device__ __forceinline__ void mm_smem_nk(float (&acc)[8][4], uint32_t sa, int r0, uint32_t sb, int lane) {
  // ldmatrix.x4 address lanes: A -> {rows 0-7 | 8-15} x {k 0-7 | 8-15};  B -> {n 0-7: k 0-7, k 8-15 | n 8-15: ...}
  const int a_row = r0 + (lane & 15), a_col = (lane >> 4) * 8;
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_col = ((lane >> 3) & 1) * 8;
  if constexpr (!H2) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[3][4];
#pragma unroll
      for (int p = 0; p < 3; ++p) ldsm_x4(sa + tile_off(p, a_row, kk * 16 + a_col), a[p]);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b[3][4];
#pragma unroll
        for (int p = 0; p < 3; ++p) ldsm_x4(sb + tile_off(p, jp * 16 + b_row, kk * 16 + b_col), b[p]);
        mma6(acc[2 * jp], a, b, 0);
        mma6(acc[2 * jp + 1], a, b, 1);
      }
    }
  } else {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // corrections a0.b1 + a1.b0 (both carry the factor 2^11)
      uint32_t a[2][4];
#pragma unroll
      for (int p = 0; p < 2; ++p) ldsm_x4(sa + tile_off(p, a_row, kk * 16 + a_col), a[p]);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b[2][4];
#pragma unroll
        for (int p = 0; p < 2; ++p) ldsm_x4(sb + tile_off(p, jp * 16 + b_row, kk * 16 + b_col), b[p]);
        mma_f16(acc[2 * jp], a[0], b[1][0], b[1][1]);
        mma_f16(acc[2 * jp], a[1], b[0][0], b[0][1]);
        mma_f16(acc[2 * jp + 1], a[0], b[1][2], b[1][3]);
        mma_f16(acc[2 * jp + 1], a[1], b[0][2], b[0][3]);
      }
    }
    scale_acc(acc, H1_INV_SCALE);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // leading product a0.b0
      uint32_t a0[4];
      ldsm_x4(sa + tile_off(0, a_row, kk * 16 + a_col), a0);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b0[4];
        ldsm_x4(sb + tile_off(0, jp * 16 + b_row, kk * 16 + b_col), b0);
        mma_f16(acc[2 * jp], a0, b0[0], b0[1]);
        mma_f16(acc[2 * jp + 1], a0, b0[2], b0[3]);
      }
    }
  }
}
// acc[16 x 64] = P[16 x 64] . B   with P in accumulator layout (fp32), B[k][n] = tile_b[k][n]; acc zero on entry
template <bool H2>
__device__ __forceinline__ void mm_regs_kn(float (&acc)[8][4], const float (&pm)[8][4], uint32_t sb, int lane) {
  // transposed ldmatrix.x4: {k 0-7 | 8-15} x {n 0-7 | 8-15}
  const int b_row = lane & 15, b_col = (lane >> 4) * 8;
  if constexpr (!H2) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[3][4];   // C tiles 2kk, 2kk+1 -> A fragment (rows g / g+8, k 2t.. / 8+2t..)
      split_pack(pm[2 * kk][0], pm[2 * kk][1], a[0][0], a[1][0], a[2][0]);
      split_pack(pm[2 * kk][2], pm[2 * kk][3], a[0][1], a[1][1], a[2][1]);
      split_pack(pm[2 * kk + 1][0], pm[2 * kk + 1][1], a[0][2], a[1][2], a[2][2]);
      split_pack(pm[2 * kk + 1][2], pm[2 * kk + 1][3], a[0][3], a[1][3], a[2][3]);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b[3][4];
#pragma unroll
        for (int p = 0; p < 3; ++p) ldsm_x4_trans(sb + tile_off(p, kk * 16 + b_row, jp * 16 + b_col), b[p]);
        mma6(acc[2 * jp], a, b, 0);
        mma6(acc[2 * jp + 1], a, b, 1);
      }
    }
  } else {
    uint32_t a0[4][4];   // leading planes of all four k-steps stay in registers for the second pass
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a1[4];
      split_pack_h2(pm[2 * kk][0], pm[2 * kk][1], a0[kk][0], a1[0]);
      split_pack_h2(pm[2 * kk][2], pm[2 * kk][3], a0[kk][1], a1[1]);
      split_pack_h2(pm[2 * kk + 1][0], pm[2 * kk + 1][1], a0[kk][2], a1[2]);
      split_pack_h2(pm[2 * kk + 1][2], pm[2 * kk + 1][3], a0[kk][3], a1[3]);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b[2][4];
#pragma unroll
        for (int p = 0; p < 2; ++p) ldsm_x4_trans(sb + tile_off(p, kk * 16 + b_row, jp * 16 + b_col), b[p]);
        mma_f16(acc[2 * jp], a0[kk], b[1][0], b[1][1]);
        mma_f16(acc[2 * jp], a1, b[0][0], b[0][1]);
        mma_f16(acc[2 * jp + 1], a0[kk], b[1][2], b[1][3]);
        mma_f16(acc[2 * jp + 1], a1, b[0][2], b[0][3]);
      }
    }
    scale_acc(acc, H1_INV_SCALE);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b0[4];
        ldsm_x4_trans(sb + tile_off(0, kk * 16 + b_row, jp * 16 + b_col), b0);
        mma_f16(acc[2 * jp], a0[kk], b0[0], b0[1]);
        mma_f16(acc[2 * jp + 1], a0[kk], b0[2], b0[3]);
      }
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&a)[8][4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0.f;
}
__device__ __forceinline__ void store_pair(float* f32, bf16* planes, long plane_stride, int nplanes, long off, float a,
                                           float c) {
  if (f32) *reinterpret_cast<float2*>(f32 + off) = make_float2(a, c);
  if (planes && nplanes == PLANES_H2) {   // loss-scaled gradient out: fp16 pair only
    uint16_t a0, a1, c0, c1;
    split_h2(a, a0, a1);
    split_h2(c, c0, c1);
    *reinterpret_cast<uint32_t*>(planes + off) = (uint32_t)a0 | ((uint32_t)c0 << 16);
    *reinterpret_cast<uint32_t*>(planes + plane_stride + off) = (uint32_t)a1 | ((uint32_t)c1 << 16);
  } else if (planes) {
    uint32_t p0, p1, p2;
    split_pack(a, c, p0, p1, p2);
    *reinterpret_cast<uint32_t*>(planes + off) = p0;
    if (nplanes > 1) *reinterpret_cast<uint32_t*>(planes + plane_stride + off) = p1;
    if (nplanes > 2) *reinterpret_cast<uint32_t*>(planes + 2 * plane_stride + off) = p2;
    if (nplanes == 5) {   // fp16 x 2 planes for the forward proj GEMM
      uint16_t a0, a1, c0, c1;
      split_h2(a, a0, a1);
      split_h2(c, c0, c1);
      *reinterpret_cast<uint32_t*>(planes + 3 * plane_stride + off) = (uint32_t)a0 | ((uint32_t)c0 << 16);
      *reinterpret_cast<uint32_t*>(planes + 4 * plane_stride + off) = (uint32_t)a1 | ((uint32_t)c1 << 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <bool H2>
__global__ void __launch_bounds__(128, 3)
attention_fwd_kernel(const bf16* __restrict__ qkv, long qkv_ps, float* __restrict__ out, bf16* __restrict__ planes,
                     long plane_stride, int nplanes, float* __restrict__ lse, int T, int H, float scale) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  const uint32_t sQ = (smem_u32(sm_raw) + 1023u) & ~1023u, sK = sQ + ATILE<H2>, sV = sK + ATILE<H2>;
  const int E = H * AD;
  const long ld = 3L * E;
  // heaviest tiles (most keys under the causal mask) are dispatched first: grid = (H, B, tiles), z is the slowest index
  const int qt = (int)(gridDim.z - 1 - blockIdx.z), h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qt * AT;
  const bf16* base = qkv + (long)b * T * ld + h * AD;
  load_tile_async<H2>(sQ, base + (long)q0 * ld, ld, qkv_ps, min(AT, T - q0));

  float o[8][4];
  zero_acc(o);
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const int r0 = warp * 16;
  const int qi0 = q0 + r0 + g, qi1 = qi0 + 8;

  for (int kt = 0; kt <= qt; ++kt) {
    const int k0 = kt * AT;
    __syncthreads();  // previous tile fully consumed
    load_tile_async<H2>(sK, base + E + (long)k0 * ld, ld, qkv_ps, min(AT, T - k0));
    load_tile_async<H2>(sV, base + 2 * E + (long)k0 * ld, ld, qkv_ps, min(AT, T - k0));
    cp_async_wait_all();
    __syncthreads();
    float s[8][4];
    zero_acc(s);
    mm_smem_nk<H2>(s, sQ, r0, sK, lane);
    float mx[2] = {mrow[0], mrow[1]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kj = k0 + j * 8 + 2 * t + (e & 1);
        const int qi = (e < 2) ? qi0 : qi1;
        float v = s[j][e] * scale;
        if (kj > qi || kj >= T) v = -INFINITY;
        s[j][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], psum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) corr[r] = (mrow[r] == -INFINITY) ? 0.f : __expf(mrow[r] - mx[r]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float m = mx[e >> 1];
        const float pv = (m == -INFINITY) ? 0.f : __expf(s[j][e] - m);
        s[j][e] = pv;
        psum[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      psum[r] += __shfl_xor_sync(0xffffffffu, psum[r], 1);
      psum[r] += __shfl_xor_sync(0xffffffffu, psum[r], 2);
      lrow[r] = lrow[r] * corr[r] + psum[r];
      mrow[r] = mx[r];
    }
    float ot[8][4];
    zero_acc(ot);
    mm_regs_kn<H2>(ot, s, sV, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] = o[j][0] * corr[0] + ot[j][0]; o[j][1] = o[j][1] * corr[0] + ot[j][1];
      o[j][2] = o[j][2] * corr[1] + ot[j][2]; o[j][3] = o[j][3] * corr[1] + ot[j][3];
    }
  }

  const float inv0 = 1.f / lrow[0], inv1 = 1.f / lrow[1];
  if (t == 0) {
    if (qi0 < T) lse[((long)b * H + h) * T + qi0] = mrow[0] + logf(lrow[0]);
    if (qi1 < T) lse[((long)b * H + h) * T + qi1] = mrow[1] + logf(lrow[1]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = r ? qi1 : qi0;
      if (qi >= T) continue;
      const float inv = r ? inv1 : inv0;
      const long off = ((long)b * T + qi) * E + h * AD + j * 8 + 2 * t;
      store_pair(out, planes, plane_stride, nplanes, off, o[j][2 * r] * inv, o[j][2 * r + 1] * inv);
    }
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
// dK, dV for one tile of 64 keys; everything is computed transposed so that keys are the MMA row index.
template <bool H2>
__global__ void __launch_bounds__(128)
attention_bwd_kv_kernel(const bf16* __restrict__ qkv, long qkv_ps, const bf16* __restrict__ dout, long do_ps,
                        const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dqkv,
                        bf16* __restrict__ planes, long plane_stride, int nplanes, int T, int H, float scale) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  const uint32_t sK = (smem_u32(sm_raw) + 1023u) & ~1023u, sV = sK + ATILE<H2>, sQ = sV + ATILE<H2>,
                 sdO = sQ + ATILE<H2>;
  __shared__ float sLse[AT], sDel[AT];
  const int E = H * AD;
  const long ld = 3L * E;
  // key tile 0 sees every query tile: heaviest first (grid = (H, B, tiles), z is the slowest index)
  const int kt = blockIdx.z, h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int k0 = kt * AT;
  const bf16* base = qkv + (long)b * T * ld + h * AD;
  const bf16* dobase = dout + (long)b * T * E + h * AD;
  load_tile_async<H2>(sK, base + E + (long)k0 * ld, ld, qkv_ps, min(AT, T - k0));
  load_tile_async<H2>(sV, base + 2 * E + (long)k0 * ld, ld, qkv_ps, min(AT, T - k0));
  const int r0 = warp * 16;
  const int kj0 = k0 + r0 + g, kj1 = kj0 + 8;

  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  const int nqt = (T + AT - 1) / AT;
  for (int qt = kt; qt < nqt; ++qt) {
    const int q0 = qt * AT;
    __syncthreads();
    load_tile_async<H2>(sQ, base + (long)q0 * ld, ld, qkv_ps, min(AT, T - q0));
    load_tile_async<H2>(sdO, dobase + (long)q0 * E, E, do_ps, min(AT, T - q0));
    if (threadIdx.x < AT) {
      const int qi = q0 + threadIdx.x;
      sLse[threadIdx.x] = qi < T ? lse[((long)b * H + h) * T + qi] : INFINITY;
      sDel[threadIdx.x] = qi < T ? delta[((long)b * H + h) * T + qi] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    float s[8][4];
    zero_acc(s);
    mm_smem_nk<H2>(s, sK, r0, sQ, lane);          // S^T = K Q^T  (rows: keys, cols: queries)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ql = j * 8 + 2 * t + (e & 1);
        const int qi = q0 + ql;
        const int kj = (e < 2) ? kj0 : kj1;
        float pv = __expf(s[j][e] * scale - sLse[ql]);
        if (kj > qi || kj >= T) pv = 0.f;
        s[j][e] = pv;  // P^T
      }
    }
    float tmp[8][4];
    zero_acc(tmp);
    mm_regs_kn<H2>(tmp, s, sdO, lane);             // dV += P^T dO
#pragma unroll
    for (int j = 0; j < 8; ++j) { dv[j][0] += tmp[j][0]; dv[j][1] += tmp[j][1]; dv[j][2] += tmp[j][2]; dv[j][3] += tmp[j][3]; }
    float dp[8][4];
    zero_acc(dp);
    mm_smem_nk<H2>(dp, sV, r0, sdO, lane);         // dP^T = V dO^T
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ql = j * 8 + 2 * t + (e & 1);
        dp[j][e] = s[j][e] * (dp[j][e] - sDel[ql]) * scale;   // dS^T
      }
    }
    zero_acc(tmp);
    mm_regs_kn<H2>(tmp, dp, sQ, lane);             // dK += dS^T Q
#pragma unroll
    for (int j = 0; j < 8; ++j) { dk[j][0] += tmp[j][0]; dk[j][1] += tmp[j][1]; dk[j][2] += tmp[j][2]; dk[j][3] += tmp[j][3]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int kj = r ? kj1 : kj0;
      if (kj >= T) continue;
      const long rowoff = ((long)b * T + kj) * 3 * E + h * AD + j * 8 + 2 * t;
      store_pair(dqkv, planes, plane_stride, nplanes, rowoff + E, dk[j][2 * r], dk[j][2 * r + 1]);
      store_pair(dqkv, planes, plane_stride, nplanes, rowoff + 2 * E, dv[j][2 * r], dv[j][2 * r + 1]);
    }
  }
}

// dQ for one tile of 64 queries
template <bool H2>
__global__ void __launch_bounds__(128)
attention_bwd_q_kernel(const bf16* __restrict__ qkv, long qkv_ps, const bf16* __restrict__ dout, long do_ps,
                       const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dqkv,
                       bf16* __restrict__ planes, long plane_stride, int nplanes, int T, int H, float scale) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  const uint32_t sQ = (smem_u32(sm_raw) + 1023u) & ~1023u, sdO = sQ + ATILE<H2>, sK = sdO + ATILE<H2>,
                 sV = sK + ATILE<H2>;
  const int E = H * AD;
  const long ld = 3L * E;
  // heaviest tiles (most keys under the causal mask) are dispatched first: grid = (H, B, tiles), z is the slowest index
  const int qt = (int)(gridDim.z - 1 - blockIdx.z), h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qt * AT;
  const bf16* base = qkv + (long)b * T * ld + h * AD;
  load_tile_async<H2>(sQ, base + (long)q0 * ld, ld, qkv_ps, min(AT, T - q0));
  load_tile_async<H2>(sdO, dout + ((long)b * T + q0) * E + h * AD, E, do_ps, min(AT, T - q0));
  const int r0 = warp * 16;
  const int qi0 = q0 + r0 + g, qi1 = qi0 + 8;
  float lse_r[2], del_r[2];
  lse_r[0] = qi0 < T ? lse[((long)b * H + h) * T + qi0] : INFINITY;
  lse_r[1] = qi1 < T ? lse[((long)b * H + h) * T + qi1] : INFINITY;
  del_r[0] = qi0 < T ? delta[((long)b * H + h) * T + qi0] : 0.f;
  del_r[1] = qi1 < T ? delta[((long)b * H + h) * T + qi1] : 0.f;

  float dq[8][4];
  zero_acc(dq);
  for (int kt = 0; kt <= qt; ++kt) {
    const int k0 = kt * AT;
    __syncthreads();
    load_tile_async<H2>(sK, base + E + (long)k0 * ld, ld, qkv_ps, min(AT, T - k0));
    load_tile_async<H2>(sV, base + 2 * E + (long)k0 * ld, ld, qkv_ps, min(AT, T - k0));
    cp_async_wait_all();
    __syncthreads();
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mm_smem_nk<H2>(s, sQ, r0, sK, lane);     // S  = Q K^T
    mm_smem_nk<H2>(dp, sdO, r0, sV, lane);   // dP = dO V^T
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kj = k0 + j * 8 + 2 * t + (e & 1);
        const int qi = (e < 2) ? qi0 : qi1;
        float pv = __expf(s[j][e] * scale - lse_r[e >> 1]);
        if (kj > qi || kj >= T) pv = 0.f;
        s[j][e] = pv * (dp[j][e] - del_r[e >> 1]) * scale;  // dS
      }
    }
    zero_acc(dp);                        // dp is dead: reuse it as the per-tile accumulator
    mm_regs_kn<H2>(dp, s, sK, lane);         // dQ += dS K
#pragma unroll
    for (int j = 0; j < 8; ++j) { dq[j][0] += dp[j][0]; dq[j][1] += dp[j][1]; dq[j][2] += dp[j][2]; dq[j][3] += dp[j][3]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = r ? qi1 : qi0;
      if (qi >= T) continue;
      const long off = ((long)b * T + qi) * 3 * E + h * AD + j * 8 + 2 * t;
      store_pair(dqkv, planes, plane_stride, nplanes, off, dq[j][2 * r], dq[j][2 * r + 1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
static int set_smem(const void* fn, size_t bytes) {
  OOB_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <bool H2>
static int attention_fwd_t(const bf16* qkv_planes, long qkv_plane_stride, float* out, bf16* out_planes,
                           long plane_stride, int nplanes, float* lse, int B, int T, int H, int D, cudaStream_t s) {
  const size_t smem = (size_t)3 * ATILE<H2> + 1024;
  static bool once = false;
  if (!once) {
    if (set_smem((const void*)attention_fwd_kernel<H2>, smem)) return -1;
    once = true;
  }
  dim3 grid(H, B, (T + AT - 1) / AT);
  attention_fwd_kernel<H2><<<grid, 128, smem, s>>>(qkv_planes, qkv_plane_stride, out, out_planes, plane_stride, nplanes,
                                                   lse, T, H, 1.0f / sqrtf((float)D));
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int attention_fwd(const bf16* qkv_planes, long qkv_plane_stride, int operand_fp16, float* out, bf16* out_planes,
                  long plane_stride, int nplanes, float* lse, int B, int T, int H, int D, cudaStream_t s) {
  OOB_CHECK(D == AD, "attention: head_dim must be 64 (got %d)", D);
  OOB_CHECK((reinterpret_cast<uintptr_t>(qkv_planes) & 15) == 0 && (qkv_plane_stride & 7) == 0,
            "attention: qkv planes misaligned");
  return operand_fp16 ? attention_fwd_t<true>(qkv_planes, qkv_plane_stride, out, out_planes, plane_stride, nplanes, lse,
                                              B, T, H, D, s)
                      : attention_fwd_t<false>(qkv_planes, qkv_plane_stride, out, out_planes, plane_stride, nplanes, lse,
                                               B, T, H, D, s);
}

template <bool H2>
static int attention_bwd_t(const bf16* qkv_planes, long qkv_plane_stride, const float* out, const float* dout,
                           const bf16* dout_planes, long dout_plane_stride, const float* lse, float* delta, float* dqkv,
                           bf16* dqkv_planes, long plane_stride, int nplanes, int B, int T, int H, int D,
                           cudaStream_t s) {
  const size_t smem = (size_t)4 * ATILE<H2> + 1024;
  static bool once = false;
  if (!once) {
    if (set_smem((const void*)attention_bwd_kv_kernel<H2>, smem)) return -1;
    if (set_smem((const void*)attention_bwd_q_kernel<H2>, smem)) return -1;
    once = true;
  }
  const int total = B * T * H;
  attention_delta_kernel<<<(total + 7) / 8, 256, 0, s>>>(out, dout, delta, B, T, H);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  dim3 grid(H, B, (T + AT - 1) / AT);
  const float scale = 1.0f / sqrtf((float)D);
  attention_bwd_kv_kernel<H2><<<grid, 128, smem, s>>>(qkv_planes, qkv_plane_stride, dout_planes, dout_plane_stride, lse,
                                                      delta, dqkv, dqkv_planes, plane_stride, nplanes, T, H, scale);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  attention_bwd_q_kernel<H2><<<grid, 128, smem, s>>>(qkv_planes, qkv_plane_stride, dout_planes, dout_plane_stride, lse,
                                                     delta, dqkv, dqkv_planes, plane_stride, nplanes, T, H, scale);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int attention_bwd(const bf16* qkv_planes, long qkv_plane_stride, int operand_fp16, const float* out, const float* dout,
                  const bf16* dout_planes, long dout_plane_stride, const float* lse, float* delta, float* dqkv,
                  bf16* dqkv_planes, long plane_stride, int nplanes, int B, int T, int H, int D, cudaStream_t s) {
  OOB_CHECK(D == AD, "attention: head_dim must be 64 (got %d)", D);
  OOB_CHECK((reinterpret_cast<uintptr_t>(qkv_planes) & 15) == 0 && (qkv_plane_stride & 7) == 0 &&
                (reinterpret_cast<uintptr_t>(dout_planes) & 15) == 0 && (dout_plane_stride & 7) == 0,
            "attention: operand planes misaligned");
  return operand_fp16 ? attention_bwd_t<true>(qkv_planes, qkv_plane_stride, out, dout, dout_planes, dout_plane_stride,
                                              lse, delta, dqkv, dqkv_planes, plane_stride, nplanes, B, T, H, D, s)
                      : attention_bwd_t<false>(qkv_planes, qkv_plane_stride, out, dout, dout_planes, dout_plane_stride,
                                               lse, delta, dqkv, dqkv_planes, plane_stride, nplanes, B, T, H, D, s);
}

}  // namespace oob
