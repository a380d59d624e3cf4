"""The dependency-free template search (csrc/planning/template_search.cpp behind ``PipelineTemplateGenerator``) against
(a) the properties the reference's own tests assert (tests/planning/test_pipeline_template.py:15-89) and (b) a pure-Python
brute force of the same cost algebra (execution_result.h:141-190) over every contiguous stage split."""
import itertools

import pytest

from oobleck_b200.planning.pipeline_template import (LayerExecutionResult, LayerExecutionResults,
                                                     PipelineTemplateGenerator)


def profiles(n, seed=0, head=2.5):
    import random
    rnd = random.Random(seed)
    out = []
    for i in range(n):
        f = 1.0 + rnd.random() if 0 < i < n - 1 else (0.05 if i == 0 else head)
        out.append(LayerExecutionResult(i, f, 2 * f, {1: 0.0, 2: 0.1, 4: 0.2}, {}, (1000 * (i + 1), 500)))
    return LayerExecutionResults(out)


def brute_force(costs, stages):
    """min over contiguous splits of T = t1 + t2 + t3 built exactly like DCExecutionResult's combine, left to right."""
    n = len(costs)
    best = None
    for cuts in itertools.combinations(range(1, n), stages - 1):
        b = [0, *cuts, n]
        lat = [sum(costs[b[i]:b[i + 1]]) for i in range(stages)]
        # the recursion can associate differently, but T only depends on the stage latencies through kstar = the FIRST
        # stage with the largest latency ... (strict > keeps the left one), t2 = (2 S + kstar + 1) lat[kstar]
        k = max(range(stages), key=lambda i: (lat[i], -i))
        t = sum(lat) + (2 * stages + k + 1) * lat[k] + sum(lat[k:])
        if best is None or t < best[0] - 1e-12:
            best = (t, b)
    return best


def test_single_node_single_gpu():                      # test_pipeline_template.py:15-26
    t = PipelineTemplateGenerator().create_pipeline_templates(profiles(8), (1, 1), 1)
    assert len(t) == 1 and t[0]._num_nodes == 1 and t[0]._num_gpus_per_node == 1
    assert len(t[0].get_stages()) == 1 and t[0]._iteration_time > 0


@pytest.mark.parametrize("num_nodes", [2, 3, 4, 6])
def test_one_gpu_per_node(num_nodes):                   # test_pipeline_template.py:28-40
    t = PipelineTemplateGenerator().create_pipeline_templates(profiles(10), (num_nodes, num_nodes), 1)
    assert len(t) == 1 and t[0]._num_nodes == num_nodes
    st = t[0].get_stages()
    assert len(st) >= num_nodes and sum(s._num_gpus for s in st) == num_nodes
    assert [i for s in st for i in s._layer_indices] == list(range(10))      # contiguous, complete, ordered


def test_too_many_nodes_is_infeasible():                # test_pipeline_template.py:42-52
    assert PipelineTemplateGenerator().create_pipeline_templates(profiles(4), (5, 5), 1) == []


def test_node_range_and_gpus_per_node():                # test_pipeline_template.py:54-89
    t = PipelineTemplateGenerator().create_pipeline_templates(profiles(12), (1, 4), 4)
    assert 1 <= len(t) <= 4
    for tpl in t:
        assert tpl._num_gpus_per_node == 4 and tpl._iteration_time > 0
        assert sum(s._num_gpus for s in tpl.get_stages()) == tpl._num_nodes * 4
        grid = tpl.get_rank_grid(list(range(tpl._num_nodes * 4)))
        assert sorted(grid) == list(range(12)) and all(len(r) == 4 for r in grid.values())


@pytest.mark.parametrize("n,nodes,seed", [(6, 2, 0), (7, 3, 1), (9, 4, 2), (10, 5, 3)])
def test_matches_brute_force_one_gpu_per_node(n, nodes, seed):
    res = profiles(n, seed)
    costs = [r._forward + r._backward for r in res.get()]
    tpl = PipelineTemplateGenerator().create_pipeline_templates(res, (nodes, nodes), 1)[0]
    # with one GPU per node every stage is one node: the search ranges over stage counts == node count only
    assert len(tpl.get_stages()) == nodes
    want_t, want_b = brute_force(costs, nodes)
    got_b = [s._layer_indices[0] for s in tpl.get_stages()] + [n]
    assert abs(tpl._iteration_time - want_t) < 1e-9 * want_t, (tpl._iteration_time, want_t, got_b, want_b)


def test_rejects_non_positive_times():
    from oobleck_b200.lib import OobleckB200Error
    bad = LayerExecutionResults([LayerExecutionResult(0, 0.0, 1.0, {}, {}, (1, 1))])
    with pytest.raises(OobleckB200Error):
        PipelineTemplateGenerator().create_pipeline_templates(bad, (1, 1), 1)
