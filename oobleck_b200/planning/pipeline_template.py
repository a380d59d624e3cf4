"""Bookkeeping stand-in for the reference's C++ planner objects (oobleck/csrc/planning/*, bound through
pipeline_template.pyi).  The planner itself (divide-and-conquer template search) is control plane and out of
scope; the hot path only needs the *shape* of its result:

* ``StageExecutionResult._layer_indices / _num_gpus``     (execution_result.h:60-112)
* ``PipelineTemplate.get_stages() / get_rank_grid(ranks)`` (pipeline_template.h:20-90)

so that ``OobleckPipeline`` can be constructed exactly as ``HeterogeneousPipelinesExecutionPlan.instantiate``
does (planning/instantiator.py:135-152).  ``even_template`` / ``balanced_template`` build templates without the
planner for benchmarks and tests.
"""
from __future__ import annotations

from typing import Sequence


class StageExecutionResult:
    def __init__(self, layer_indices: Sequence[int], num_gpus: int = 1):
        self._layer_indices = list(layer_indices)
        self._num_gpus = num_gpus
        self._size = len(self._layer_indices)

    def num_layers(self) -> int:
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
