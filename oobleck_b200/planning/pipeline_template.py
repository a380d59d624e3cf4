"""The reference's C++ planner objects (oobleck/csrc/planning/*, bound through pipeline_template.pyi) with the same
class names and attributes.  The hot path needs the *shape* of the planner's result:

* ``StageExecutionResult._layer_indices / _num_gpus``     (execution_result.h:60-112)
* ``PipelineTemplate.get_stages() / get_rank_grid(ranks)`` (pipeline_template.h:20-90)

so that ``OobleckPipeline`` can be constructed exactly as ``HeterogeneousPipelinesExecutionPlan.instantiate``
does (planning/instantiator.py:135-152).  ``even_template`` / ``balanced_template`` build templates without the
planner for benchmarks and tests.

``LayerExecutionResult(s)`` and ``PipelineTemplateGenerator.create_pipeline_templates`` are the dependency-free rebuild of
the template search itself: the divide-and-conquer of pipeline_template.cpp:82-339 with the cost algebra
of execution_result.h:60-205, in C++ behind the C ABI (csrc/planning/template_search.cpp; the reference's build needs
cppcoro and oneTBB, neither of which exists here), fed by ``planning/profiler.py``'s measured latencies.
"""
from __future__ import annotations

from typing import Sequence


class LayerExecutionResult:
    """pipeline_template.pyi:1-16 / execution_result.h:15-38."""

    def __init__(self, layer_index: int, forward: float, backward: float, allreduce_in_node: dict[int, float],
                 allreduce_across_nodes: dict[int, float], mem_required: tuple[int, int]):
        self._index = layer_index
        self._forward = float(forward)
        self._backward = float(backward)
        self._allreduce_in_node = dict(allreduce_in_node)
        self._allreduce_across_nodes = dict(allreduce_across_nodes)
        self._mem_required = (int(mem_required[0]), int(mem_required[1]))


class _Size(int):
    """``LayerExecutionResults.size``: a read-only PROPERTY in the reference's binding (bind.cpp:38, and how its tests use
    it: ``profile.size``), a METHOD in its stub file (pipeline_template.pyi:21).  Both spellings work here."""

    def __call__(self) -> int:
        return int(self)


class LayerExecutionResults:
    """pipeline_template.pyi:18-21 / bind.cpp:33-38."""

    def __init__(self, data: list[LayerExecutionResult]):
        self._data = list(data)

    def get(self) -> list[LayerExecutionResult]:
        return self._data

    def at(self, index: int) -> LayerExecutionResult:
        return self._data[index]

    @property
    def size(self) -> _Size:
        return _Size(len(self._data))


class StageExecutionResult:
    """execution_result.h:60-112.  Two ways to build one: the reference's ``StageExecutionResult(layer_results,
    (begin, end), num_gpus)`` (bind.cpp:40-43), which aggregates the layers' times and memory like the C++ constructor,
    and the bookkeeping-only ``StageExecutionResult(layer_indices, num_gpus)`` used where only the shape matters."""

    def __init__(self, layers_or_indices, layer_indices_or_num_gpus=1, num_gpus: int | None = None):
        self._forward = 0.0
        self._backward = 0.0
        self._mem_required = 0
        if isinstance(layers_or_indices, LayerExecutionResults):
            begin, end = layer_indices_or_num_gpus
            assert num_gpus is not None and end <= layers_or_indices.size
            self._num_gpus = int(num_gpus)
            self._layer_indices = []
            for i in range(begin, end):
                layer = layers_or_indices.at(i)
                assert layer._forward > 0 and layer._backward > 0                      # execution_result.h:74-75
                self._layer_indices.append(layer._index)
                self._forward += layer._forward / self._num_gpus
                self._backward += layer._backward / self._num_gpus
                if self._num_gpus > 1:                                                  # :81-84 (``.at``: must exist)
                    self._forward += layer._allreduce_in_node[self._num_gpus]
                    self._backward += layer._allreduce_in_node[self._num_gpus]
                self._mem_required += 6 * layer._mem_required[0] + layer._mem_required[1]
        else:
            self._layer_indices = list(layers_or_indices)
            self._num_gpus = int(layer_indices_or_num_gpus if num_gpus is None else num_gpus)
        self._size = len(self._layer_indices)

    def num_layers(self) -> int:
        return len(self._layer_indices)

    @property
    def _num_layers(self) -> int:          # bind.cpp:48
        return len(self._layer_indices)


class PipelineTemplate:
    def __init__(self, stages: list[StageExecutionResult], iteration_time: float, num_layers: int, num_nodes: int,
                 num_gpus_per_node: int):
        assert sum(s._num_gpus for s in stages) == num_nodes * num_gpus_per_node    # pipeline_template.h:36-41
        assert sum(s.num_layers() for s in stages) == num_layers                     # :43-47
        self._stages = stages
        self._iteration_time = iteration_time
        self._num_nodes = num_nodes
        self._num_gpus_per_node = num_gpus_per_node

    def get_stages(self) -> list[StageExecutionResult]:
        return self._stages

    def get_rank_grid(self, ranks: list[int]) -> dict[int, list[int]]:
        """layer index -> ``num_gpus_per_node`` ranks; each stage rank repeated ``gpn / stage_gpus`` times
        (pipeline_template.h:57-84)."""
        assert len(ranks) == sum(s._num_gpus for s in self._stages)
        grid: dict[int, list[int]] = {}
        cursor = 0
        for stage in self._stages:
            mine = ranks[cursor: cursor + stage._num_gpus]
            cursor += stage._num_gpus
            repeat = self._num_gpus_per_node // stage._num_gpus
            row = [r for r in mine for _ in range(repeat)]
            row += [0] * (self._num_gpus_per_node - len(row))   # std::vector<int>(gpn) zero-fill when not divisible
            for layer_index in stage._layer_indices:
                grid[layer_index] = list(row)
        return dict(sorted(grid.items()))


def even_template(num_layers: int, num_stages: int, num_nodes: int | None = None,
                  num_gpus_per_node: int = 1) -> PipelineTemplate:
    """Contiguous, as-even-as-possible split (earlier stages get the extra layers)."""
    num_nodes = num_stages if num_nodes is None else num_nodes
    base, extra = divmod(num_layers, num_stages)
    stages, start = [], 0
    for s in range(num_stages):
        n = base + (1 if s < extra else 0)
        stages.append(StageExecutionResult(range(start, start + n), 1))
        start += n
    return PipelineTemplate(stages, 0.0, num_layers, num_nodes, num_gpus_per_node)


def balanced_template(layer_costs: Sequence[float], num_stages: int, num_gpus_per_node: int = 1) -> PipelineTemplate:
    """Contiguous split minimising the most expensive stage (linear-partition DP over per-layer costs)."""
    n = len(layer_costs)
    assert 1 <= num_stages <= n
    prefix = [0.0]
    for c in layer_costs:
        prefix.append(prefix[-1] + c)
    INF = float("inf")
    best = [[INF] * (n + 1) for _ in range(num_stages + 1)]
    cut = [[0] * (n + 1) for _ in range(num_stages + 1)]
    best[0][0] = 0.0
    for s in range(1, num_stages + 1):
        for j in range(s, n + 1):
            for i in range(s - 1, j):
                cost = max(best[s - 1][i], prefix[j] - prefix[i])
                if cost < best[s][j]:
                    best[s][j], cut[s][j] = cost, i
    bounds, j = [], n
    for s in range(num_stages, 0, -1):
        i = cut[s][j]
        bounds.append((i, j))
        j = i
    stages = [StageExecutionResult(range(a, b), 1) for a, b in reversed(bounds)]
    return PipelineTemplate(stages, best[num_stages][n], n, num_stages, num_gpus_per_node)


class PipelineTemplateGenerator:
    """pipeline_template.pyi:53-62: ``create_pipeline_templates(layer_execution_results, (min_nodes, max_nodes),
    num_gpus_per_node) -> list[PipelineTemplate]``, one template per feasible node count, stages chosen by the
    divide-and-conquer search (``oob_plan_pipeline_templates``)."""

    def create_pipeline_templates(self, layer_execution_results: LayerExecutionResults, num_nodes: tuple[int, int],
                                  num_gpus_per_node: int) -> list[PipelineTemplate]:
        import ctypes as C

        from .. import lib as L
        lib = L.load()
        n = layer_execution_results.size()
        prof = (L.LayerProfile * n)()
        stride = num_gpus_per_node + 1
        ar = (C.c_double * (n * stride))()
        for i, r in enumerate(layer_execution_results.get()):
            prof[i].forward, prof[i].backward = r._forward, r._backward
            prof[i].mem_params, prof[i].mem_activations = r._mem_required
            for g, v in r._allreduce_in_node.items():
                if 0 <= int(g) < stride:
                    ar[i * stride + int(g)] = float(v)
        min_nodes, max_nodes = num_nodes
        cap = max(1, max_nodes - min_nodes + 1) * (2 + 3 * n)
        out = (C.c_int * cap)()
        times = (C.c_double * max(1, max_nodes - min_nodes + 1))()
        count = C.c_int(0)
        rc = lib.oob_plan_pipeline_templates(prof, n, ar, stride, num_gpus_per_node, min_nodes, max_nodes, out, cap, times,
                                             C.byref(count))
        if rc != 0:
            raise L.OobleckB200Error(f"oob_plan_pipeline_templates failed ({rc}): layer times must be positive")
        templates, pos = [], 0
        for t in range(count.value):
            nodes, nstages = out[pos], out[pos + 1]
            pos += 2
            stages = []
            for _ in range(nstages):
                begin, end, gpus = out[pos], out[pos + 1], out[pos + 2]
                pos += 3
                st = StageExecutionResult(range(begin, end), gpus)
                for i in range(begin, end):            # the aggregates the search itself used (execution_result.h:66-100)
                    layer = layer_execution_results.at(i)
                    ar = float(layer._allreduce_in_node.get(gpus, 0.0)) if gpus > 1 else 0.0
                    st._forward += layer._forward / gpus + ar
                    st._backward += layer._backward / gpus + ar
                    st._mem_required += 6 * layer._mem_required[0] + layer._mem_required[1]
                stages.append(st)
            templates.append(PipelineTemplate(stages, times[t], n, nodes, num_gpus_per_node))
        return templates
