// Shared device/host helpers for the sm_100a kernels: error plumbing, inline-PTX wrappers for
// mbarrier / TMA / tcgen05 / TMEM, and the split-bf16 ("bf16x3") representation of fp32 operands.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace oob {

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI returns int status; message via oob_last_error())
void set_error(const char* fmt, ...);
// number of kernels this library has launched (bench.py reports it as gpu_launches)
void count_launch(int n = 1);
#define OOB_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      oob::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                           \
    }                                                                                      \
  } while (0)
#define OOB_CHECK(cond, ...)                                                               \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      oob::set_error(__VA_ARGS__);                                                         \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Split representation.  An fp32 value x is carried to the tensor cores as up to three bf16 planes
// p0 + p1 + p2 == x to 24 mantissa bits (p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1)).
// A GEMM with nsplit planes issues the products whose combined weight is >= 2^-8(nsplit-1):
//   nsplit 1: p0q0                      (bf16 accuracy,  1 MMA)
//   nsplit 2: + p0q1 + p1q0             (~2^-16,        3 MMAs)
//   nsplit 3: + p1q1 + p0q2 + p2q0      (~2^-23, fp32-grade, 6 MMAs)   <- parity mode
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void split3(float x, bf16& p0, bf16& p1, bf16& p2) {
  p0 = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(p0);
  p1 = __float2bfloat16_rn(r);
  r -= __bfloat162float(p1);
  p2 = __float2bfloat16_rn(r);
}

// Second operand format: TWO fp16 planes, x ~= h0 + 2^-11 * h1 with h0 = fp16(x), h1 = fp16((x - h0) * 2^11).
// 22 significand bits (fp16 carries 11 per plane; the residual is pre-scaled so it never falls into fp16 subnormals
// before x itself is below ~2^-14), so the three products h0g0 + 2^-11 (h0g1 + h1g0) are fp32-grade at half the
// tensor work of the six bf16 products.  fp16 has a 5-bit exponent: only tensors with a bounded range use it
// (weights, LayerNorm / GELU / attention outputs -- the forward operands); gradients stay in bf16 x 3.
// Conversions saturate (no inf) so an outlier degrades precision instead of poisoning the GEMM.
constexpr float H1_SCALE = 2048.f;
constexpr float H1_INV_SCALE = 1.f / 2048.f;
__device__ __forceinline__ uint16_t f2h_sat(float x) {
  uint16_t h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
__device__ __forceinline__ float h2f(uint16_t h) {
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
  return f;
}
__device__ __forceinline__ void split_h2(float x, uint16_t& h0, uint16_t& h1) {
  h0 = f2h_sat(x);
  h1 = f2h_sat((x - h2f(h0)) * H1_SCALE);
}
// raw 16-bit patterns of all five planes of a "5-plane" buffer: [bf16 p0 p1 p2 | fp16 h0 h1]
__device__ __forceinline__ void split5(float x, uint16_t& q0, uint16_t& q1, uint16_t& q2, uint16_t& q3, uint16_t& q4) {
  bf16 a, b, c;
  split3(x, a, b, c);
  q0 = __bfloat16_as_ushort(a); q1 = __bfloat16_as_ushort(b); q2 = __bfloat16_as_ushort(c);
  split_h2(x, q3, q4);
}

// tanh(u) = 1 - 2 / (1 + e^{2u}) on the SFU (ex2 + rcp, ~2 ulp each; saturates correctly at +-inf).  The libm
// tanhf is ~40 instructions with a branch; the GEMM epilogues evaluate it for every FC element and their code has
// to stay small enough for the instruction cache (profiles/README.md, epilogue section).
__device__ __forceinline__ float tanh_fast(float u) {
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * u));
}
// Plane-set codes accepted by every producer's `nplanes` argument:
//   1..3        bf16 planes p0..  (x = p0 + p1 + p2)
//   5           bf16 x 3 followed by the fp16 pair
//   PLANES_H2   the fp16 pair only, at planes 0 and 1 (loss-scaled gradients, backward GEMM operands)
constexpr int PLANES_H2 = 22;
__host__ __device__ __forceinline__ int planes_count(int code) { return code == PLANES_H2 ? 2 : code; }
// which of split5's five outputs goes to stored plane i
__host__ __device__ __forceinline__ int planes_src(int code, int i) { return code == PLANES_H2 ? 3 + i : i; }

__device__ __forceinline__ float gelu_new_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanh_fast(u));
}
__device__ __forceinline__ float gelu_new_grad_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  float t = tanh_fast(u);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x * x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 3-D tiled load: coordinates (c0 = innermost element index, c1 = row, c2 = plane)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation: whole warp, writes base address to *smem_slot.  ncols power of two >= 32.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit: arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive 32-bit columns; thread i of the warp gets lane (base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (tcgen05), see PTX ISA "Matrix Descriptor" / cute::UMMA::SmemDescriptor:
//  [0,14) start>>4  [16,30) LBO>>4  [32,46) SBO>>4  [46,48) version=1  [49,52) base offset  [61,64) swizzle
enum : uint64_t { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint64_t swizzle) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= swizzle << 61;
  return d;
}
// Instruction descriptor for kind::f16 with fp32 accumulation (cute::UMMA::InstrDescriptor); operand formats:
// 1 = BF16, 0 = F16 (independent fields for A and B)
__host__ __device__ constexpr uint32_t make_idesc_f16kind(int M, int N, int a_mn_major, int b_mn_major,
                                                          int a_bf16 = 1, int b_bf16 = 1) {
  return (1u << 4)                       // c_format  = F32
         | ((uint32_t)a_bf16 << 7)       // a_format
         | ((uint32_t)b_bf16 << 10)      // b_format
         | ((uint32_t)a_mn_major << 15)  // a_major   (0 = K-major, 1 = MN-major)
         | ((uint32_t)b_mn_major << 16)  // b_major
         | ((uint32_t)(N >> 3) << 17)    // n_dim
         | ((uint32_t)(M >> 4) << 24);   // m_dim
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// 128-bit streaming global accesses
__device__ __forceinline__ float4 ldg_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void stg_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace oob
