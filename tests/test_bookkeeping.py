"""Oracle bookkeeping vs (a) golden vectors generated from the reference's own Python
(tests/golden/gen_golden.py) and (b) the tables of the reference's own tests."""
import json
import os

import pytest
import torch

from oracle import bookkeeping as bk
from oracle import schedule as osched

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


NUM_LAYERS = 34


def init_ranks(gpn):
    out, used = [], 0
    for i in range(2, 6):
        out.append(list(range(used, used + i * gpn)))
        used += i * gpn
    return out


def test_reconfigure_golden():
    cases = load("reconfigure")
    assert len(cases) >= 100
    for c in cases:
        gpn = c["gpus_per_node"]
        got = bk.reconfigure_ranks(init_ranks(gpn), c["failed"], 2 * gpn)
        assert got == c["result"]["ranks"], c


# tests/execution/test_reconfiguration.py:151-217 (no FSDP) and :252-383 (FSDP) -- the reference's tables
REF_TABLE = [
    (1, [2], [[0, 1], [3, 4], [5, 6, 7, 8], [9, 10, 11, 12, 13]]),
    (1, [6, 8], [[0, 1], [5, 7], [2, 3, 4], [9, 10, 11, 12, 13]]),
    (1, [10, 11], [[0, 1], [2, 3, 4], [9, 12, 13], [5, 6, 7, 8]]),
    (1, [1], [[0, 13], [2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]),
    (1, [1, 3, 4], [[0, 13], [2, 12], [9, 10, 11], [5, 6, 7, 8]]),
    (1, [2, 4, 6, 7, 8], [[0, 1], [3, 13], [5, 12], [9, 10, 11]]),
    (1, [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11], [[0, 12, 13]]),
    (1, [1, 2, 3, 5, 6, 7, 9, 11, 12, 13], [[0, 4], [8, 10]]),
    (1, [1, 2, 3, 5, 6, 7, 9, 10, 11], [[0, 4], [8, 12, 13]]),
    (2, [6, 7], [list(range(0, 4)), [4, 5, 8, 9], list(range(10, 18)), list(range(18, 28))]),
    (2, [10, 11, 18, 19], [list(range(0, 4)), list(range(4, 10)), [12, 13, 14, 15, 16, 17],
                           [20, 21, 22, 23, 24, 25, 26, 27]]),
    (4, [8, 9, 10, 11], [list(range(0, 8)), list(range(12, 20)), list(range(20, 36)), list(range(36, 56))]),
    (4, [20, 21, 22, 23, 28, 29, 30, 31], [list(range(0, 8)), [24, 25, 26, 27, 32, 33, 34, 35],
                                           list(range(8, 20)), list(range(36, 56))]),
    (2, [2, 3], [[0, 1, 26, 27], list(range(4, 10)), list(range(10, 18)), list(range(18, 26))]),
    (2, [2, 3, 4, 5, 8, 9], [[0, 1, 26, 27], [6, 7, 24, 25], list(range(18, 24)), list(range(10, 18))]),
    (2, [2, 3, 10, 11, 14, 15, 16, 17], [[0, 1, 26, 27], [12, 13, 24, 25], list(range(4, 10)),
                                         list(range(18, 24))]),
    (4, [4, 5, 6, 7], [[0, 1, 2, 3, 52, 53, 54, 55], list(range(8, 20)), list(range(20, 36)),
                       list(range(36, 52))]),
    (4, [4, 5, 6, 7, 8, 9, 10, 11, 16, 17, 18, 19], [[0, 1, 2, 3, 52, 53, 54, 55],
                                                     [12, 13, 14, 15, 48, 49, 50, 51], list(range(36, 48)),
                                                     list(range(20, 36))]),
    (4, [4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 36, 37, 38, 39],
     [[0, 1, 2, 3, 52, 53, 54, 55], [16, 17, 18, 19, 32, 33, 34, 35], list(range(20, 32)), list(range(40, 52))]),
    (2, [2, 3, 8, 9, 14, 15, 16, 17, 22, 23, 24, 25, 26, 27], [[10, 11, 12, 13], [18, 19, 20, 21],
                                                               [0, 1, 4, 5, 6, 7]]),
    (2, [2, 3, 6, 7, 8, 9, 10, 11, 14, 15, 16, 17, 22, 23, 24, 25, 26, 27], [[0, 1, 4, 5],
                                                                             [12, 13, 18, 19, 20, 21]]),
]


@pytest.mark.parametrize("gpn,failed,expected", REF_TABLE)
def test_reconfigure_reference_tables(gpn, failed, expected):
    assert bk.reconfigure_ranks(init_ranks(gpn), failed, 2 * gpn) == expected


def test_reconfigure_insufficient_raises():
    with pytest.raises(RuntimeError, match="insufficient"):
        bk.reconfigure_ranks([[0, 1]], [1], 2)


def build_rank_grids(gpn, nodes, npipes, nstages):
    grids, used = [], 0
    for n, k, s in zip(nodes, npipes, nstages):
        t = bk.dummy_template(NUM_LAYERS, s, gpn, n)
        for _ in range(k):
            cnt = n * gpn
            grids.append(t.get_rank_grid(list(range(used, used + cnt))))
            used += cnt
    return grids


def test_dp_groups_golden():
    for c in load("dp_groups"):
        grids = build_rank_grids(c["gpus_per_node"], c["nodes"], c["num_pipelines"], c["stages"])
        grid = bk.dp_grid(grids, c["gpus_per_node"])
        got = {str(l): {str(f): r for f, r in d.items()} for l, d in grid.items()}
        assert got == c["groups"]
        order = [r for d in grid.values() for r in d.values()]
        assert order == c["order"]


# tests/execution/test_engine.py:135-225: every (ranks, fsdp_index) pair below must be an all-reduce group
DP_TABLE = [
    (4, [1, 2], [1, 1], [2, 2], [((0, 4), 0), ((0, 5), 1), ((1, 6), 2), ((1, 7), 3), ((2, 8), 0), ((2, 9), 1),
                                 ((3, 10), 2), ((3, 11), 3)]),
    (4, [3], [2], [4], [((0, 12), 0), ((0, 12), 1), ((1, 13), 2), ((1, 13), 3), ((2, 14), 0), ((3, 15), 3),
                        ((4, 16), 0), ((5, 17), 1), ((6, 18), 2), ((7, 19), 3), ((8, 20), 0), ((11, 23), 3)]),
    (4, [3, 5], [2, 1], [4, 5], [((0, 12, 24), 0), ((0, 12, 25), 1), ((1, 13, 26), 2), ((1, 13, 27), 3),
                                 ((0, 12, 28), 0), ((2, 14, 28), 0), ((3, 15, 31), 3), ((2, 14, 32), 0),
                                 ((4, 16, 32), 0), ((7, 19, 35), 3), ((4, 16, 36), 0), ((8, 20, 36), 0),
                                 ((11, 23, 39), 3), ((8, 20, 40), 0), ((11, 23, 43), 3)]),
]


@pytest.mark.parametrize("gpn,nodes,npipes,nstages,expected", DP_TABLE)
def test_dp_groups_reference_tables(gpn, nodes, npipes, nstages, expected):
    grid = bk.dp_grid(build_rank_grids(gpn, nodes, npipes, nstages), gpn)
    have = {(tuple(sorted(set(r))), f) for d in grid.values() for f, r in d.items()}
    for key in expected:
        assert key in have, key


def test_sampler_golden():
    for c in load("sampler"):
        for pi, want in enumerate(c["batches"]):
            got = bk.sampler_batches(c["num_samples"], c["microbatch_size"], pi, c["num_microbatches"],
                                     c["epoch"], c["shuffle"])
            assert got == want
        # tests/execution/test_dataloader.py:146-166: no sample is shared between pipelines
        flat = [i for p in c["batches"] for b in p for i in b]
        assert len(flat) == len(set(flat))


def test_schedule_golden():
    for c in load("schedule"):
        M, P, s = c["micro_batches"], c["stages"], c["stage_id"]
        assert osched.num_pipe_buffers(M, P, s) == c["num_pipe_buffers"]
        got = [[[n, b] for n, b in cmds] for cmds in osched.steps(M, P, s)]
        assert got == c["steps"]


def test_schedule_survey_table():
    # SURVEY.md 8(c) golden table, M=4 P=2 stage 0 (first rows)
    st = osched.steps(4, 2, 0)
    assert st[0] == [("LoadMicroBatch", 0), ("ForwardPass", 0)]
    assert st[1] == [("SendActivation", 0)]
    assert st[3] == [("RecvGrad", 0), ("SendActivation", 1), ("BackwardPass", 0)]
    assert st[9] == [("RecvGrad", 1), ("BackwardPass", 1)]
    assert len(st) == 2 * (4 + 2 - 1)


def test_schedule_every_microbatch_once():
    for M, P in [(64, 8), (5, 3), (1, 4)]:
        for s in range(P):
            flat = [c for cmds in osched.steps(M, P, s) for c in cmds]
            assert sum(1 for n, _ in flat if n == "ForwardPass") == M
            assert sum(1 for n, _ in flat if n == "BackwardPass") == M


def test_dtype_ids_golden():
    want = load("dtype_ids")
    got = {str(k).replace("torch.", ""): v for k, v in bk.DTYPE_TO_ID.items()}
    assert got == want
    assert bk.DTYPE_TO_ID[torch.float32] == 0 and bk.DTYPE_TO_ID[torch.int64] == 10


def test_rank_grid_and_neighbours():
    # pipeline_template.h:57-84 and tests/execution/test_pipeline.py:193-198
    t = bk.dummy_template(NUM_LAYERS, 4, 1, 4)
    grid = t.get_rank_grid([0, 1, 2, 3])
    assert list(grid) == list(range(NUM_LAYERS))
    assert grid[0] == [0] and grid[NUM_LAYERS - 1] == [3]
    for r in range(4):
        _, prev, nxt = bk.pipeline_neighbours(grid, r)
        assert prev == (r - 1 if r > 0 else None) and nxt == (r + 1 if r < 3 else None)
    assert sum(len(bk.my_layers(grid, r)) for r in range(4)) == NUM_LAYERS
    # stage with 2 GPUs on a 4-GPU node: each rank repeated twice
    t2 = bk.dummy_template(NUM_LAYERS, 2, 4, 1)
    g2 = t2.get_rank_grid([0, 1, 2, 3])
    assert g2[0] == [0, 0, 1, 1] and g2[NUM_LAYERS - 1] == [2, 2, 3, 3]


def test_copy_plan_single_pipeline_loss_raises():
    # tests/execution/test_engine.py:1015-1019: a lone 4-stage pipeline losing a node cannot recover
    t4 = bk.dummy_template(NUM_LAYERS, 4, 1, 4)
    t3 = bk.dummy_template(NUM_LAYERS, 3, 1, 3)
    with pytest.raises(RuntimeError, match="No alive ranks"):
        bk.copy_plan([t4.get_rank_grid([0, 1, 2, 3])], [t3.get_rank_grid([0, 1, 2])])


def test_model_layers_are_callable_objects_for_the_profiler():
    """planning/profiler.py:66-91, 272-274 treats ``model.layers`` entries as modules: init_tensors(layer, device),
    copy.deepcopy(layer).to("cuda"), layer(*input) -> tuple, layer.parameters().  Without a GPU the call must fail
    loudly (no CPU fallback)."""
    import copy

    import pytest
    import torch

    from oobleck_b200.execution.layer import init_tensors
    from oobleck_b200.lib import OobleckB200Error
    from oobleck_b200.module.model import OobleckModel, StageLayer
    m = OobleckModel("gpt2", {"input_ids": None}, None, "t", dict(n_embd=64, n_head=1, n_layer=2, n_positions=32,
                                                                   vocab_size=101))
    assert len(m.layers) == 4 and all(isinstance(l, StageLayer) and callable(l) for l in m.layers)
    assert sum(p.numel() for l in m.layers for p in l.parameters()) == m.total_num_params
    clone = copy.deepcopy(m.layers[1]).to("cuda")
    assert clone is not m.layers[1] and clone.spec is m.layers[1].spec and clone.kind == "block"
    with pytest.raises(OobleckB200Error):
        init_tensors(m.layers[0], torch.device("cpu"))
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            clone(torch.zeros(1, 32, 64), torch.zeros(1, 32, dtype=torch.int64))
