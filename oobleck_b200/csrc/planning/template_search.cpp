// Pipeline-template search, dependency-free (the control-plane-facing half of SURVEY 8(f2)).
//
// Rebuilds what oobleck/csrc/planning/pipeline_template.cpp:82-339 + execution_result.h:60-205 compute -- for every
// node count n in [min_nodes, max_nodes] the stage split of the layer list (and the GPU split inside a node) that
// minimises the 1F1B iteration-time estimate T = t1 + t2 + t3 -- without cppcoro / oneTBB / pybind11 (none of which
// exist in this image): a plain memoised recursion behind a C entry point.  Cost algebra as in the reference:
//   stage(layers [a, b), g GPUs):  fwd = sum layer.fwd / g (+ allreduce_in_node[g] per layer if g > 1), same for bwd
//   leaf:     t1 = t3 = fwd + bwd, t2 = 2 (fwd + bwd), kstar = 0
//   combine:  kstar = left's if left.kstar_latency > right.kstar_latency else right.kstar + |left|
//             t1 = l.t1 + r.t1;  t2 = (2 (|l| + |r|) + kstar + 1) * kstar_latency;  t3 = sum of fwd + bwd from kstar on
//   feasibility: stages <= layers; one node: stages <= gpus, a single stage needs a power-of-two GPU count, GPUs split
//   evenly between the two halves; several nodes: nodes <= stages.
// Differences, on purpose: every accumulator starts at zero (StageExecutionResult::forward_ / backward_ /
// mem_required_ are read uninitialised in the reference, execution_result.h:78-112); memory is 64-bit.
#include <cmath>
#include <cstdint>
#include <map>
#include <unordered_map>
#include <memory>
#include <tuple>
#include <vector>

#include "../../../include/oobleck_b200.h"

namespace {

struct Stage {
  int begin = 0, end = 0, gpus = 1;
  double fwd = 0.0, bwd = 0.0;
  long long mem = 0;
};

struct Plan {   // DCExecutionResult
  std::vector<std::shared_ptr<Stage>> stages;
  int kstar = 0;
  double t1 = 0.0, t2 = 0.0, t3 = 0.0;
  double t() const { return t1 + t2 + t3; }
  double kstar_latency() const { return stages[kstar]->fwd + stages[kstar]->bwd; }
};

struct Searcher {
  const oob_layer_profile* layers;
  int num_layers;
  const double* ar_in_node;   // [num_layers][ar_stride], column g = all-reduce time over g GPUs of a node; may be null
  int ar_stride;
  using Key = uint64_t;   // stages, begin, end, nodes, gpus per node: 12 bits each
  std::unordered_map<Key, std::shared_ptr<Plan>> cache;
  static Key make_key(int stages, int begin, int end, int nodes, int gpn) {
    return ((Key)stages << 48) | ((Key)begin << 36) | ((Key)end << 24) | ((Key)nodes << 12) | (Key)gpn;
  }

  std::shared_ptr<Plan> leaf(int begin, int end, int gpus) {
    auto st = std::make_shared<Stage>();
    st->begin = begin; st->end = end; st->gpus = gpus;
    for (int i = begin; i < end; ++i) {
      st->fwd += layers[i].forward / gpus;
      st->bwd += layers[i].backward / gpus;
      if (gpus > 1 && ar_in_node && gpus < ar_stride) {
        st->fwd += ar_in_node[(size_t)i * ar_stride + gpus];
        st->bwd += ar_in_node[(size_t)i * ar_stride + gpus];
      }
      st->mem += 6 * layers[i].mem_params + layers[i].mem_activations;
    }
    auto p = std::make_shared<Plan>();
    p->stages = {st};
    p->t1 = p->t3 = st->fwd + st->bwd;
    p->t2 = 2 * (st->fwd + st->bwd);
    return p;
  }

  static std::shared_ptr<Plan> combine(const std::shared_ptr<Plan>& l, const std::shared_ptr<Plan>& r) {
    auto p = std::make_shared<Plan>();
    p->stages = l->stages;
    const bool left_k = l->kstar_latency() > r->kstar_latency();
    p->kstar = left_k ? l->kstar : r->kstar + (int)l->stages.size();
    p->t1 = l->t1 + r->t1;
    const double n_mb = 2.0 * (l->stages.size() + r->stages.size()) + p->kstar + 1;
    double latency = 0.0;
    if (left_k) {
      p->t2 = n_mb * l->kstar_latency();
      for (size_t i = l->kstar; i < l->stages.size(); ++i) latency += l->stages[i]->fwd + l->stages[i]->bwd;
      for (auto& s : r->stages) latency += s->fwd + s->bwd;
    } else {
      p->t2 = n_mb * r->kstar_latency();
      for (size_t i = r->kstar; i < r->stages.size(); ++i) latency += r->stages[i]->fwd + r->stages[i]->bwd;
    }
    p->t3 = latency;
    p->stages.insert(p->stages.end(), r->stages.begin(), r->stages.end());
    return p;
  }

  std::shared_ptr<Plan> solve(int stages, int begin, int end, int nodes, int gpn) {
    const Key key = make_key(stages, begin, end, nodes, gpn);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    // (stages > nodes * gpn can never bottom out -- every leaf is one stage on >= 1 GPU of one node -- so the reference
    // would return null for it after exploring the whole subtree; cut it here)
    bool infeasible = stages > end - begin || stages > nodes * gpn;
    if (nodes == 1) {
      if (gpn < stages) infeasible = true;
      const double lg = std::log2((double)gpn);
      if (stages == 1 && lg != std::trunc(lg)) infeasible = true;
    } else if (nodes > stages) {
      infeasible = true;
    }
    std::shared_ptr<Plan> best;
    if (!infeasible) {
      if (stages == 1) {
        best = leaf(begin, end, gpn);
      } else {
        for (int k = begin + 1; k < end; ++k) {
          if (nodes == 1) {
            for (int gl = 1; gl < gpn; ++gl) {
              if (gl != gpn - gl) continue;   // GPUs of a node are split evenly
              for (int sl = 1; sl < stages; ++sl) {
                auto l = solve(sl, begin, k, 1, gl);
                auto r = solve(stages - sl, k, end, 1, gpn - gl);
                if (!l || !r) continue;
                auto c = combine(l, r);
                if (!best || c->t() < best->t()) best = c;
              }
            }
          } else {
            for (int nl = 1; nl < nodes; ++nl) {
              for (int sl = 1; sl < stages; ++sl) {
                auto l = solve(sl, begin, k, nl, gpn);
                auto r = solve(stages - sl, k, end, nodes - nl, gpn);
                if (!l || !r) continue;
                auto c = combine(l, r);
                if (!best || c->t() < best->t()) best = c;
              }
            }
          }
        }
      }
    }
    cache.emplace(key, best);
    return best;
  }
};

}  // namespace

extern "C" int oob_plan_pipeline_templates(const oob_layer_profile* layers, int num_layers, const double* allreduce_in_node,
                                           int allreduce_stride, int num_gpus_per_node, int min_nodes, int max_nodes,
                                           int* out, int out_capacity, double* iteration_time, int* num_templates) {
  if (num_layers >= 4096 || max_nodes >= 4096 || num_gpus_per_node >= 4096) return -2;
  if (!layers || num_layers <= 0 || !out || !num_templates || !iteration_time || num_gpus_per_node < 1 || min_nodes < 1)
    return -2;
  for (int i = 0; i < num_layers; ++i)
    if (!(layers[i].forward > 0) || !(layers[i].backward > 0)) return -2;   // execution_result.h:74-75
  Searcher s{layers, num_layers, allreduce_in_node, allreduce_stride, {}};
  int pos = 0, count = 0;
  for (int n = min_nodes; n <= max_nodes; ++n) {
    std::shared_ptr<Plan> best;
    for (int stages = n; stages <= num_layers; ++stages) {   // pipeline_template.cpp:96-107
      auto p = s.solve(stages, 0, num_layers, n, num_gpus_per_node);
      if (p && (!best || p->t() < best->t())) best = p;
    }
    if (!best) continue;                                      // "All results are invalid"
    const int need = 2 + 3 * (int)best->stages.size();
    if (pos + need > out_capacity) return -3;
    out[pos++] = n;
    out[pos++] = (int)best->stages.size();
    for (auto& st : best->stages) {
      out[pos++] = st->begin;
      out[pos++] = st->end;
      out[pos++] = st->gpus;
    }
    iteration_time[count++] = best->t();
  }
  *num_templates = count;
  return 0;
}
