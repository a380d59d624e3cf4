"""Oracle: the integer bookkeeping of the hot path, restated with plain lists/dicts.

Every function cites the reference lines it follows.  Pinned bit-exactly by
tests/golden/*.json (generated from the reference's own Python, see tests/golden/gen_golden.py)
and by the tables in the reference's tests (ported in tests/test_bookkeeping.py).
"""
from __future__ import annotations

import copy
import math
from collections import defaultdict

import torch

# oobleck/execution/utils.py:4-18 -- wire enum for the P2P meta handshake
ID_TO_DTYPE = [
    torch.float32, torch.float64, torch.complex64, torch.complex128, torch.float16,
    torch.bfloat16, torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64, torch.bool,
]
DTYPE_TO_ID = {dt: i for i, dt in enumerate(ID_TO_DTYPE)}


class Stage:
    """StageExecutionResult's bookkeeping fields (csrc/planning/execution_result.h:60-112)."""

    def __init__(self, layer_indices: list[int], num_gpus: int):
        self._layer_indices = list(layer_indices)
        self._num_gpus = num_gpus


class Template:
    """PipelineTemplate's bookkeeping (csrc/planning/pipeline_template.h:20-90)."""

    def __init__(self, stages: list[Stage], num_nodes: int, num_gpus_per_node: int):
        self._stages = stages
        self._num_nodes = num_nodes
        self._num_gpus_per_node = num_gpus_per_node

    def get_stages(self):
        return self._stages

    def get_rank_grid(self, ranks: list[int]) -> dict[int, list[int]]:
        return get_rank_grid(self._stages, self._num_gpus_per_node, ranks)


def get_rank_grid(stages: list[Stage], num_gpus_per_node: int, ranks: list[int]) -> dict[int, list[int]]:
    """pipeline_template.h:57-84."""
    ranks = list(ranks)
    grid: dict[int, list[int]] = {}
    for stage in stages:
        stage_ranks, ranks = ranks[: stage._num_gpus], ranks[stage._num_gpus:]
        repeat = num_gpus_per_node // stage._num_gpus
        layer_ranks = [0] * num_gpus_per_node
        pos = 0
        for r in stage_ranks:
            for _ in range(repeat):
                layer_ranks[pos] = r
                pos += 1
        for li in stage._layer_indices:
            grid[li] = list(layer_ranks)
    assert len(ranks) == 0
    return dict(sorted(grid.items()))  # std::map iteration order


def dummy_template(num_layers: int, num_stages: int, num_gpus_per_node: int, num_nodes: int) -> Template:
    """tests/conftest.py:144-213 (get_dummy_pipeline_template): even layer slices, power-of-two GPU split."""
    length_chunk = math.ceil(num_layers / num_stages)
    slices = []
    for i in range(0, num_layers, length_chunk):
        slices.append((i, min(i + length_chunk, num_layers)))
    per_stage: dict[int, int] = defaultdict(int)
    per_stage[1] = num_nodes * num_gpus_per_node
    while sum(per_stage.values()) > num_stages:
        m = min(n for n in per_stage if per_stage[n] >= 2)
        per_stage[m] -= 2
        per_stage[m * 2] += 1
    sub = []
    for n in sorted(per_stage):
        sub.extend([n] * per_stage[n])
    assert len(slices) == len(sub)
    return Template([Stage(list(range(a, b)), g) for (a, b), g in zip(slices, sub)], num_nodes, num_gpus_per_node)


def pipeline_neighbours(rank_grid: dict[int, list[int]], my_rank: int):
    """pipeline.py:593-611: per shard column, group = list(set(ranks)); prev/next by index.
    Returns (shard_id, prev_rank, next_rank) of the LAST column containing my_rank (the reference
    overwrites ``self.communication`` in the loop), or None."""
    found = None
    for shard_id in range(len(rank_grid[next(iter(rank_grid))])):
        ranks = [per_layer[shard_id] for per_layer in rank_grid.values()]
        if my_rank in ranks:
            unique = list(set(ranks))
            i = unique.index(my_rank)
            found = (shard_id, unique[i - 1] if i > 0 else None, unique[i + 1] if i < len(unique) - 1 else None)
    return found


def my_layers(rank_grid: dict[int, list[int]], my_rank: int) -> list[int]:
    """pipeline.py:503-524: layers whose per-layer group (set(ranks), :575-578) contains me, grid order."""
    return [li for li, ranks in rank_grid.items() if my_rank in set(ranks)]


def my_stage_index(stages: list[Stage], rank_grid: dict[int, list[int]], my_rank: int) -> int:
    """pipeline.py:532-546."""
    first = next(li for li, ranks in rank_grid.items() if my_rank in ranks)
    return next(i for i, s in enumerate(stages) if first in s._layer_indices)


def dp_grid(rank_grids: list[dict[int, list[int]]], num_gpus_per_node: int) -> dict[int, dict[int, list[int]]]:
    """engine.py:374-398: layer -> fsdp_index -> ranks (one group each, insertion order)."""
    grid: dict[int, dict[int, list[int]]] = defaultdict(dict)
    for rg in rank_grids:
        for layer_index, ranks in rg.items():
            assert len(ranks) == num_gpus_per_node
            for fsdp_index, rank in enumerate(ranks):
                grid[layer_index].setdefault(fsdp_index, []).append(rank)
    return grid


def dp_groups_for_rank(grid: dict[int, dict[int, list[int]]], layer_id: int, my_rank: int) -> dict[int, list[int]]:
    """engine.py:404-412: the {fsdp_index: group} dict handed to Layer.reduce_gradients."""
    return {fi: ranks for fi, ranks in grid[layer_id].items() if my_rank in ranks}


def shard_param_sizes(numel: int, number: int) -> list[int]:
    """layer.py:262-269 (_shard_param): chunk sizes after padding the last chunk."""
    chunk = math.ceil(numel / number)
    n_chunks = math.ceil(numel / chunk) if numel else 0
    sizes = [chunk] * max(n_chunks, 0)
    if len(sizes) < number:
        sizes += [chunk] * (number - len(sizes))
    return sizes


# ---------------------------------------------------------------------------------------------
# Reconfiguration policy (engine.py:91-166, 311-360)


def reconfigure_ranks(pipeline_ranks: list[list[int]], lost_ranks: list[int], min_num_ranks: int) -> list[list[int]]:
    """Returns the new rank lists in final order.  Mutation order follows the reference exactly
    (``ranks`` lists are shared with the pipelines and mutated by ``pop``/``append``)."""
    pipes = [[r for r in ranks if r not in lost_ranks] for ranks in pipeline_ranks]  # :111-114

    def find_biggest():  # :346-360
        biggest = None
        for p in pipes:
            if biggest is None or len(p) >= len(biggest):
                biggest = p
        if biggest is not None and len(biggest) > min_num_ranks:
            return biggest
        return None

    need_merge = False
    new_list: list[list[int]] = []
    for ranks in pipes:  # :119-146
        if len(ranks) == 0:
            continue
        if len(ranks) >= min_num_ranks:
            new_list.append(ranks)
            continue
        while len(ranks) < min_num_ranks:
            biggest = find_biggest()
            if biggest is None:
                need_merge = True
                break
            while len(biggest) > min_num_ranks and len(ranks) < min_num_ranks:
                ranks.append(biggest.pop())
        new_list.append(ranks)

    if need_merge:  # :311-344
        to_merge, results = [], []
        for ranks in new_list:
            (to_merge if len(ranks) < min_num_ranks else results).append(ranks)
        try:
            while to_merge:
                ranks = to_merge.pop(0)
                try:
                    while len(ranks) < min_num_ranks:
                        ranks.extend(to_merge.pop(0))
                except IndexError:
                    ranks.extend(results.pop(0))
                assert len(ranks) >= min_num_ranks
                results.append(ranks)
        except IndexError:
            raise RuntimeError("Ranks are insufficient")
        new_list = results

    for ranks in new_list:  # :153-158
        ranks.sort()
    new_list.sort(key=lambda r: (len(r), r[0]))
    return new_list


def copy_plan(old_rank_grids: list[dict[int, list[int]]], new_rank_grids: list[dict[int, list[int]]]):
    """engine.py:250-306: per layer, (ranks_to_send, [ranks_recv...]) or None if nothing moves.
    Raises the reference's RuntimeError when no old rank list survives unchanged."""
    plan = {}
    for layer_index in range(len(old_rank_grids[0])):
        old_ranks = [g[layer_index] for g in old_rank_grids]
        new_ranks = [g[layer_index] for g in new_rank_grids]
        if all(r in old_ranks for r in new_ranks):
            plan[layer_index] = None
            continue
        alive = [r for r in old_ranks if r in new_ranks]
        if not alive:
            raise RuntimeError(f"No alive ranks for the layer {layer_index}. Terminating.")
        plan[layer_index] = (alive[0], copy.deepcopy(new_ranks))
    return plan


# ---------------------------------------------------------------------------------------------
# Sampler (dataloader.py:43-100)


def sampler_batches(num_samples: int, microbatch_size: int, pipeline_index: int, num_microbatches: list[int],
                    epoch: int = 0, shuffle: bool = True, seed: int = 0) -> list[list[int]]:
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        indices = torch.randperm(num_samples, generator=g).tolist()
    else:
        indices = list(range(num_samples))
    bucket = microbatch_size * sum(num_microbatches)
    offset = sum(num_microbatches[:pipeline_index]) * microbatch_size
    out = []
    for it in range(num_samples // bucket):
        if num_samples - it * bucket < bucket:
            break
        for mb in range(num_microbatches[pipeline_index]):
            lo = it * bucket + mb * microbatch_size + offset
            out.append(indices[lo: lo + microbatch_size])
    return out
