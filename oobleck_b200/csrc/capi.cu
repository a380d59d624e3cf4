// extern "C" surface of liboobleck_b200.so (declared in include/oobleck_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/oobleck_b200.h"
#include "kernels.h"

namespace oob {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace oob

using namespace oob;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline PlaneMat PM(const oob_planes* p) {
  return PlaneMat{reinterpret_cast<const bf16*>(p->base), p->rows, p->cols, p->ld, p->plane_stride, p->nplanes,
                  p->format};
}

extern "C" {

int oob_version(void) { return 100; }
long oob_launch_count(void) { return g_launches.load(); }
int oob_gemm_timing_begin(void) { return gemm_timing_begin(); }
int oob_gemm_timing_end(double* total_ms, double* total_flops, double* executed_flops, long* launches) {
  return gemm_timing_end(total_ms, total_flops, executed_flops, launches);
}
const char* oob_last_error(void) { return g_err; }
long oob_tensor_map_encodes(void) { return tensor_map_encodes(); }
long oob_ln_bwd_partials_floats(int n_embd) { return (long)LN_BWD_MAX_GRID * 2 * n_embd; }
long oob_colsum_partials_floats(int cols) { return (long)COLSUM_MAX_PARTS * cols; }

int oob_split_planes(const float* x, void* planes, long n, long plane_stride, int nplanes, void* stream) {
  return split_planes(x, reinterpret_cast<bf16*>(planes), n, plane_stride, nplanes, S(stream));
}

int oob_gemm(const oob_planes* a, int a_mn, const oob_planes* b, int b_mn, int M, int N, int K, int nsplit,
             const oob_gemm_epilogue* e, void* stream) {
  OOB_CHECK(a && b && e, "oob_gemm: null argument");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.nsplit = nsplit;
  p.epi.d = e->d; p.epi.ldd = e->ldd; p.epi.bias = e->bias; p.epi.resid = e->resid; p.epi.ldr = e->ldr;
  p.epi.accumulate = e->accumulate; p.epi.act = e->act; p.epi.aux = e->aux; p.epi.ldaux = e->ldaux;
  p.epi.planes = reinterpret_cast<bf16*>(e->planes); p.epi.ldp = e->ldp; p.epi.plane_stride = e->plane_stride;
  p.epi.nplanes_out = e->nplanes_out; p.epi.alpha = e->alpha;
  OOB_CHECK(p.epi.d || p.epi.planes, "oob_gemm: no output");
  OOB_CHECK(!(p.epi.accumulate && !p.epi.d), "oob_gemm: accumulate needs d");
  OOB_CHECK(!(p.epi.act == ACT_DGELU && !p.epi.aux), "oob_gemm: dGELU needs aux");
  return gemm_launch(PM(a), a_mn, PM(b), b_mn, p, S(stream));
}

int oob_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, void* y_planes,
                      long plane_stride, int nplanes, float* mean, float* rstd, int rows, int n_embd, float eps,
                      void* stream) {
  return layernorm_fwd(x, gamma, beta, y, reinterpret_cast<bf16*>(y_planes), plane_stride, nplanes, mean, rstd, rows,
                       n_embd, eps, S(stream));
}
int oob_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                      const float* dres, float* dx, void* dx_planes, long plane_stride, int nplanes, float* dgamma,
                      float* dbeta, float* partials, int rows, int n_embd, float param_grad_scale, void* stream) {
  return layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, reinterpret_cast<bf16*>(dx_planes), plane_stride, nplanes,
                       dgamma, dbeta, partials, rows, n_embd, param_grad_scale, S(stream));
}
int oob_colsum_accumulate(const float* a, long lda, int rows, int cols, float* out, float* partials, float scale,
                          void* stream) {
  return colsum_accumulate(a, lda, rows, cols, out, partials, scale, S(stream));
}

int oob_attention_fwd(const void* qkv_planes, long qkv_plane_stride, int operand_fp16, float* out, void* out_planes,
                      long plane_stride, int nplanes, float* lse, int batch, int seq, int n_head, int head_dim,
                      void* stream) {
  return attention_fwd(reinterpret_cast<const bf16*>(qkv_planes), qkv_plane_stride, operand_fp16, out,
                       reinterpret_cast<bf16*>(out_planes), plane_stride, nplanes, lse, batch, seq, n_head, head_dim,
                       S(stream));
}
int oob_attention_bwd(const void* qkv_planes, long qkv_plane_stride, int operand_fp16, const float* out, const float* dout,
                      const void* dout_planes, long dout_plane_stride, const float* lse, float* delta, float* dqkv,
                      void* dqkv_planes, long plane_stride, int nplanes, int batch, int seq, int n_head, int head_dim,
                      void* stream) {
  return attention_bwd(reinterpret_cast<const bf16*>(qkv_planes), qkv_plane_stride, operand_fp16, out, dout,
                       reinterpret_cast<const bf16*>(dout_planes), dout_plane_stride, lse, delta, dqkv,
                       reinterpret_cast<bf16*>(dqkv_planes), plane_stride, nplanes, batch, seq, n_head, head_dim,
                       S(stream));
}

int oob_embedding_fwd(const long long* ids, const float* wte, const float* wpe, float* hidden, int rows, int seq,
                      int n_embd, void* stream) {
  return embedding_fwd(ids, wte, wpe, hidden, rows, seq, n_embd, S(stream));
}
int oob_embedding_bwd(const long long* ids, const float* dhidden, float* dwte, float* dwpe, int batch, int seq,
                      int n_embd, float scale, void* stream) {
  return embedding_bwd(ids, dhidden, dwte, dwpe, batch, seq, n_embd, scale, S(stream));
}

int oob_cross_entropy(const float* logits, long ldl, const long long* labels, int batch, int seq, int vocab,
                      float* row_loss, float* loss, float* total_loss, void* dlogits_planes, long ldp,
                      long plane_stride, int nplanes, float grad_scale, void* stream) {
  return cross_entropy(logits, ldl, labels, batch, seq, vocab, row_loss, loss, total_loss,
                       reinterpret_cast<bf16*>(dlogits_planes), ldp, plane_stride, nplanes, grad_scale, S(stream));
}

int oob_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* planes,
                   long plane_stride, int nplanes, long n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, void* stream) {
  return adamw_step(param, grad, exp_avg, exp_avg_sq, reinterpret_cast<bf16*>(planes), plane_stride, nplanes, n, lr,
                    beta1, beta2, eps, weight_decay, step, S(stream));
}

}  // extern "C"
