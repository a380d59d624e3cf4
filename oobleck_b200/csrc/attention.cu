// Causal multi-head self-attention for the GPT-2 stage layer (HF GPT2Attention eager path:
// softmax(where(causal, QK^T/sqrt(d), finfo.min)) V, fp32), flash-style: the T x T score matrix never
// reaches HBM (the reference materialises mb*H*T^2 fp32 = 210 MB per GPT-2-XL block).
//
// Tensor-core path: warp-level mma.sync.m16n8k8 TF32 with the 3-term split (hi*hi + hi*lo + lo*hi) done in
// registers, which reproduces fp32 products to ~2^-21 -- the fp32-parity counterpart of the split-bf16 tcgen05
// GEMMs.  64 x 64 tiles, head_dim == 64, 4 warps x 16 rows.  Tiles are staged in shared memory with a row
// stride of 68 floats, which makes both fragment access patterns below bank-conflict free.
//
//   forward        grid (T/64, H, B): S = QK^T, online softmax, O = PV; writes O (fp32 + split planes) and LSE
//   backward dK,dV grid (T/64 kv tiles, H, B): S^T = KQ^T, dV += P^T dO, dP^T = V dO^T, dK += dS^T Q
//   backward dQ    grid (T/64 q tiles, H, B):  S = QK^T, dP = dO V^T, dQ += dS K
#include "kernels.h"

namespace oob {

constexpr int AT = 64;    // tile (queries or keys)
constexpr int AD = 64;    // head dim
constexpr int ALD = 68;   // smem row stride (floats)

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// c += a * b to fp32 accuracy: lo*hi + hi*lo first, hi*hi last
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}

// A fragment (16 x 8) of a row-major smem tile: rows r0.., columns k0..
__device__ __forceinline__ void load_a_smem(const float* tile, int r0, int k0, int g, int t, uint32_t (&ah)[4],
                                            uint32_t (&al)[4]) {
  split_tf32(tile[(r0 + g) * ALD + k0 + t], ah[0], al[0]);
  split_tf32(tile[(r0 + g + 8) * ALD + k0 + t], ah[1], al[1]);
  split_tf32(tile[(r0 + g) * ALD + k0 + t + 4], ah[2], al[2]);
  split_tf32(tile[(r0 + g + 8) * ALD + k0 + t + 4], ah[3], al[3]);
}
// A fragment from an accumulator n-tile (C layout -> A layout with the contraction index permuted:
// fragment column t <-> C column 2t, column t+4 <-> C column 2t+1; the B side uses the same permutation).
__device__ __forceinline__ void load_a_acc(const float (&c)[4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
  split_tf32(c[0], ah[0], al[0]);
  split_tf32(c[2], ah[1], al[1]);
  split_tf32(c[1], ah[2], al[2]);
  split_tf32(c[3], ah[3], al[3]);
}
// B fragment (8 x 8) where B[k][n] = tile[n0+n][k0+k]  ("transposed" operand, e.g. K in QK^T)
__device__ __forceinline__ void load_b_nk(const float* tile, int n0, int k0, int g, int t, uint32_t (&bh)[2],
                                          uint32_t (&bl)[2]) {
  split_tf32(tile[(n0 + g) * ALD + k0 + t], bh[0], bl[0]);
  split_tf32(tile[(n0 + g) * ALD + k0 + t + 4], bh[1], bl[1]);
}
// B fragment where B[k][n] = tile[k0+perm(k)][n0+n] with the permutation matching load_a_acc
__device__ __forceinline__ void load_b_kn_perm(const float* tile, int k0, int n0, int g, int t, uint32_t (&bh)[2],
                                               uint32_t (&bl)[2]) {
  split_tf32(tile[(k0 + 2 * t) * ALD + n0 + g], bh[0], bl[0]);
  split_tf32(tile[(k0 + 2 * t + 1) * ALD + n0 + g], bh[1], bl[1]);
}

// acc[16 x 64] (+)= A_tile[r0..r0+16, 0..64] * B^T where B rows are the tile's rows  (C = A . tile^T)
__device__ __forceinline__ void mm_a_smem_b_nk(float (&acc)[8][4], const float* a_tile, int r0, const float* b_tile,
                                               int g, int t) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    uint32_t ah[4], al[4];
    load_a_smem(a_tile, r0, ks * 8, g, t, ah, al);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t bh[2], bl[2];
      load_b_nk(b_tile, j * 8, ks * 8, g, t, bh, bl);
      mma3(acc[j], ah, al, bh, bl);
    }
  }
}
// acc[16 x 64] += P[16 x 64] * tile[64 x 64], P given in accumulator layout
__device__ __forceinline__ void mm_a_acc_b_kn(float (&acc)[8][4], const float (&p)[8][4], const float* b_tile, int g,
                                              int t) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    uint32_t ah[4], al[4];
    load_a_acc(p[ks], ah, al);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t bh[2], bl[2];
      load_b_kn_perm(b_tile, ks * 8, j * 8, g, t, bh, bl);
      mma3(acc[j], ah, al, bh, bl);
    }
  }
}

// cooperative load of a [64 x 64] fp32 tile (row stride ld_g in global) into smem, rows >= valid zero-filled
__device__ __forceinline__ void load_tile(float* dst, const float* src, long ld_g, int valid_rows) {
  for (int i = threadIdx.x; i < AT * (AD / 4); i += blockDim.x) {
    const int r = i / (AD / 4), c = (i % (AD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < valid_rows) v = ldg_f4(src + (long)r * ld_g + c);
    *reinterpret_cast<float4*>(dst + r * ALD + c) = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attention_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, bf16* __restrict__ planes,
                     long plane_stride, int nplanes, float* __restrict__ lse, int T, int H, float scale) {
  extern __shared__ float sm[];
  float* sQ = sm;                 // [64][68]
  float* sK = sQ + AT * ALD;
  float* sV = sK + AT * ALD;
  const int E = H * AD;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qt * AT;
  const float* base = qkv + (long)b * T * 3 * E + h * AD;
  load_tile(sQ, base + (long)q0 * 3 * E, 3 * E, min(AT, T - q0));

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const int r0 = warp * 16;
  const int qi0 = q0 + r0 + g, qi1 = qi0 + 8;

  for (int kt = 0; kt <= qt; ++kt) {
    const int k0 = kt * AT;
    __syncthreads();  // previous tile fully consumed (also orders the sQ fill on the first trip)
    load_tile(sK, base + E + (long)k0 * 3 * E, 3 * E, min(AT, T - k0));
    load_tile(sV, base + 2 * E + (long)k0 * 3 * E, 3 * E, min(AT, T - k0));
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    mm_a_smem_b_nk(s, sQ, r0, sK, g, t);
    // scale + causal mask + running max
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
    // P.V goes into a fresh accumulator and is folded with a round-to-nearest add: the HMMA accumulator
    // truncates, so a long-running accumulation over all key tiles would drift (see gemm_sm100.cuh)
    float ot[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) ot[j][0] = ot[j][1] = ot[j][2] = ot[j][3] = 0.f;
    mm_a_acc_b_kn(ot, s, sV, g, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] = o[j][0] * corr[0] + ot[j][0]; o[j][1] = o[j][1] * corr[0] + ot[j][1];
      o[j][2] = o[j][2] * corr[1] + ot[j][2]; o[j][3] = o[j][3] * corr[1] + ot[j][3];
    }
  }

  // epilogue: O / l, LSE
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
      const float a = o[j][2 * r] * (r ? inv1 : inv0), c = o[j][2 * r + 1] * (r ? inv1 : inv0);
      const long off = ((long)b * T + qi) * E + h * AD + j * 8 + 2 * t;
      if (out) *reinterpret_cast<float2*>(out + off) = make_float2(a, c);
      if (planes) {
        bf16 x0, x1, x2, y0, y1, y2;
        split3(a, x0, x1, x2);
        split3(c, y0, y1, y2);
        const bf16 xs[3] = {x0, x1, x2}, ys[3] = {y0, y1, y2};
#pragma unroll
        for (int p = 0; p < 3; ++p)
          if (p < nplanes)
            *reinterpret_cast<uint32_t*>(planes + p * plane_stride + off) =
                (uint32_t)__bfloat16_as_ushort(xs[p]) | ((uint32_t)__bfloat16_as_ushort(ys[p]) << 16);
      }
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
__global__ void __launch_bounds__(128)
attention_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                        const float* __restrict__ delta, float* __restrict__ dqkv, bf16* __restrict__ planes,
                        long plane_stride, int nplanes, int T, int H, float scale) {
  extern __shared__ float sm[];
  float* sK = sm;
  float* sV = sK + AT * ALD;
  float* sQ = sV + AT * ALD;
  float* sdO = sQ + AT * ALD;
  float* sLse = sdO + AT * ALD;  // [64]
  float* sDel = sLse + AT;       // [64]
  const int E = H * AD;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int k0 = kt * AT;
  const float* base = qkv + (long)b * T * 3 * E + h * AD;
  const float* dobase = dout + (long)b * T * E + h * AD;
  load_tile(sK, base + E + (long)k0 * 3 * E, 3 * E, min(AT, T - k0));
  load_tile(sV, base + 2 * E + (long)k0 * 3 * E, 3 * E, min(AT, T - k0));
  const int r0 = warp * 16;
  const int kj0 = k0 + r0 + g, kj1 = kj0 + 8;

  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  const int nqt = (T + AT - 1) / AT;
  for (int qt = kt; qt < nqt; ++qt) {
    const int q0 = qt * AT;
    __syncthreads();
    load_tile(sQ, base + (long)q0 * 3 * E, 3 * E, min(AT, T - q0));
    load_tile(sdO, dobase + (long)q0 * E, E, min(AT, T - q0));
    if (threadIdx.x < AT) {
      const int qi = q0 + threadIdx.x;
      sLse[threadIdx.x] = qi < T ? lse[((long)b * H + h) * T + qi] : INFINITY;
      sDel[threadIdx.x] = qi < T ? delta[((long)b * H + h) * T + qi] : 0.f;
    }
    __syncthreads();
    // S^T = K Q^T  (rows: keys, cols: queries)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    mm_a_smem_b_nk(s, sK, r0, sQ, g, t);
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
    // dV += P^T dO   (fresh accumulator per tile + RN fold, see forward)
    float tmp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) tmp[j][0] = tmp[j][1] = tmp[j][2] = tmp[j][3] = 0.f;
    mm_a_acc_b_kn(tmp, s, sdO, g, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) { dv[j][0] += tmp[j][0]; dv[j][1] += tmp[j][1]; dv[j][2] += tmp[j][2]; dv[j][3] += tmp[j][3]; }
    // dP^T = V dO^T
    float dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    mm_a_smem_b_nk(dp, sV, r0, sdO, g, t);
    // dS^T = P^T * (dP^T - delta) * scale
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ql = j * 8 + 2 * t + (e & 1);
        dp[j][e] = s[j][e] * (dp[j][e] - sDel[ql]) * scale;
      }
    }
    // dK += dS^T Q
#pragma unroll
    for (int j = 0; j < 8; ++j) tmp[j][0] = tmp[j][1] = tmp[j][2] = tmp[j][3] = 0.f;
    mm_a_acc_b_kn(tmp, dp, sQ, g, t);
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
      const float vals[2][2] = {{dk[j][2 * r], dk[j][2 * r + 1]}, {dv[j][2 * r], dv[j][2 * r + 1]}};
#pragma unroll
      for (int w = 0; w < 2; ++w) {  // 0: dK (column block E), 1: dV (column block 2E)
        const long off = rowoff + (w + 1) * E;
        if (dqkv) *reinterpret_cast<float2*>(dqkv + off) = make_float2(vals[w][0], vals[w][1]);
        if (planes) {
          bf16 x0, x1, x2, y0, y1, y2;
          split3(vals[w][0], x0, x1, x2);
          split3(vals[w][1], y0, y1, y2);
          const bf16 xs[3] = {x0, x1, x2}, ys[3] = {y0, y1, y2};
#pragma unroll
          for (int p = 0; p < 3; ++p)
            if (p < nplanes)
              *reinterpret_cast<uint32_t*>(planes + p * plane_stride + off) =
                  (uint32_t)__bfloat16_as_ushort(xs[p]) | ((uint32_t)__bfloat16_as_ushort(ys[p]) << 16);
        }
      }
    }
  }
}

// dQ for one tile of 64 queries
__global__ void __launch_bounds__(128)
attention_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                       const float* __restrict__ delta, float* __restrict__ dqkv, bf16* __restrict__ planes,
                       long plane_stride, int nplanes, int T, int H, float scale) {
  extern __shared__ float sm[];
  float* sQ = sm;
  float* sdO = sQ + AT * ALD;
  float* sK = sdO + AT * ALD;
  float* sV = sK + AT * ALD;
  const int E = H * AD;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qt * AT;
  const float* base = qkv + (long)b * T * 3 * E + h * AD;
  load_tile(sQ, base + (long)q0 * 3 * E, 3 * E, min(AT, T - q0));
  load_tile(sdO, dout + ((long)b * T + q0) * E + h * AD, E, min(AT, T - q0));
  const int r0 = warp * 16;
  const int qi0 = q0 + r0 + g, qi1 = qi0 + 8;
  float lse_r[2], del_r[2];
  lse_r[0] = qi0 < T ? lse[((long)b * H + h) * T + qi0] : INFINITY;
  lse_r[1] = qi1 < T ? lse[((long)b * H + h) * T + qi1] : INFINITY;
  del_r[0] = qi0 < T ? delta[((long)b * H + h) * T + qi0] : 0.f;
  del_r[1] = qi1 < T ? delta[((long)b * H + h) * T + qi1] : 0.f;

  float dq[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
  for (int kt = 0; kt <= qt; ++kt) {
    const int k0 = kt * AT;
    __syncthreads();
    load_tile(sK, base + E + (long)k0 * 3 * E, 3 * E, min(AT, T - k0));
    load_tile(sV, base + 2 * E + (long)k0 * 3 * E, 3 * E, min(AT, T - k0));
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
    mm_a_smem_b_nk(s, sQ, r0, sK, g, t);     // S  = Q K^T
    mm_a_smem_b_nk(dp, sdO, r0, sV, g, t);   // dP = dO V^T
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
    // dQ += dS K   (dp is dead here: reuse it as the per-tile accumulator)
#pragma unroll
    for (int j = 0; j < 8; ++j) dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    mm_a_acc_b_kn(dp, s, sK, g, t);
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
      const float a = dq[j][2 * r], c = dq[j][2 * r + 1];
      if (dqkv) *reinterpret_cast<float2*>(dqkv + off) = make_float2(a, c);
      if (planes) {
        bf16 x0, x1, x2, y0, y1, y2;
        split3(a, x0, x1, x2);
        split3(c, y0, y1, y2);
        const bf16 xs[3] = {x0, x1, x2}, ys[3] = {y0, y1, y2};
#pragma unroll
        for (int p = 0; p < 3; ++p)
          if (p < nplanes)
            *reinterpret_cast<uint32_t*>(planes + p * plane_stride + off) =
                (uint32_t)__bfloat16_as_ushort(xs[p]) | ((uint32_t)__bfloat16_as_ushort(ys[p]) << 16);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
static int set_smem(const void* fn, size_t bytes) {
  OOB_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

int attention_fwd(const float* qkv, float* out, bf16* out_planes, long plane_stride, int nplanes, float* lse, int B,
                  int T, int H, int D, cudaStream_t s) {
  OOB_CHECK(D == AD, "attention: head_dim must be 64 (got %d)", D);
  const size_t smem = (size_t)3 * AT * ALD * sizeof(float);
  static bool once = false;
  if (!once) {
    if (set_smem((const void*)attention_fwd_kernel, smem)) return -1;
    once = true;
  }
  dim3 grid((T + AT - 1) / AT, H, B);
  attention_fwd_kernel<<<grid, 128, smem, s>>>(qkv, out, out_planes, plane_stride, nplanes, lse, T, H,
                                               1.0f / sqrtf((float)D));
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int attention_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* delta, float* dqkv,
                  bf16* dqkv_planes, long plane_stride, int nplanes, int B, int T, int H, int D, cudaStream_t s) {
  OOB_CHECK(D == AD, "attention: head_dim must be 64 (got %d)", D);
  const size_t smem_kv = (size_t)(4 * AT * ALD + 2 * AT) * sizeof(float);
  const size_t smem_q = (size_t)4 * AT * ALD * sizeof(float);
  static bool once = false;
  if (!once) {
    if (set_smem((const void*)attention_bwd_kv_kernel, smem_kv)) return -1;
    if (set_smem((const void*)attention_bwd_q_kernel, smem_q)) return -1;
    once = true;
  }
  const int total = B * T * H;
  attention_delta_kernel<<<(total + 7) / 8, 256, 0, s>>>(out, dout, delta, B, T, H);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  dim3 grid((T + AT - 1) / AT, H, B);
  const float scale = 1.0f / sqrtf((float)D);
  attention_bwd_kv_kernel<<<grid, 128, smem_kv, s>>>(qkv, dout, lse, delta, dqkv, dqkv_planes, plane_stride, nplanes,
                                                     T, H, scale);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  attention_bwd_q_kernel<<<grid, 128, smem_q, s>>>(qkv, dout, lse, delta, dqkv, dqkv_planes, plane_stride, nplanes, T,
                                                   H, scale);
  OOB_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace oob
