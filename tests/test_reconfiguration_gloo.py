"""Reconfiguration mechanism on CPU/gloo with 4 ranks: 2 replicas x 2 stages lose rank 3; the survivors re-plan with
the reference's policy ([[2], [0, 1]]), rebuild pipelines without touching the world group, copy the missing layers
from the surviving replica (engine.py:238-309) and keep training.  Also: a lone pipeline losing a stage raises the
reference's RuntimeError (tests/execution/test_engine.py:1015-1019)."""
import os
import socket
import sys
import time

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
MARGS = dict(n_embd=64, n_head=1, num_hidden_layers=2, n_positions=32, vocab_size=211)


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    try:
        import torch.distributed as dist
        from oracle_layer import OracleLayer

        from oobleck_b200.execution.dataloader import SyntheticTokenDataset
        from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine
        from oobleck_b200.planning.pipeline_template import even_template
        M, mb = 4, 1
        args = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=mb * M, steps=1),
                                model=ModelArguments(model_name="gpt2", model_tag="t", model_args=dict(MARGS)))
        ds = SyntheticTokenDataset(num_samples=256, seq_len=32, vocab_size=211, pin_memory=False)
        templates = [even_template(4, 1), even_template(4, 2)]
        eng = OobleckEngine(rank, world, 1, None, args, dataset=ds, layer_cls=OracleLayer, templates=templates)
        eng.initialize_distributed("gloo")
        eng.instantiate_pipelines(M, plan=[templates[1], templates[1]])     # ranks [0,1] and [2,3]
        for _ in range(2):
            eng._train_step()
        before = {l.layer_id: l.flat_param.clone() for l in eng._pipeline.execution._layers}
        dist.barrier()
        if rank == 3:                  # the "lost node": leaves without taking part in anything below
            q.put((rank, "gone", None))
            return
        t0 = time.perf_counter()
        eng._reconfiguration.on_reconfigure([3])
        dt = time.perf_counter() - t0
        new_ranks = [p._ranks for p in eng._reconfiguration._pipelines]
        assert new_ranks == [[2], [0, 1]], new_ranks
        pipe = eng._pipeline
        mine = sorted(l.layer_id for l in pipe.execution._layers)
        assert mine == {0: [0, 1], 1: [2, 3], 2: [0, 1, 2, 3]}[rank]
        # layers this rank already owned are reused untouched; rank 2 received layers 2,3 from rank 1
        for l in pipe.execution._layers:
            if l.layer_id in before:
                assert torch.equal(l.flat_param, before[l.layer_id])
        for _ in range(2):
            eng._train_step()
        out = {l.layer_id: l.flat_param.numpy().copy() for l in pipe.execution._layers}
        q.put((rank, out, dt))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None))
        raise


@pytest.mark.timeout(300)
def test_lose_one_rank_and_continue():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(4):
        r = q.get(timeout=240)
        results[r[0]] = r
    for p in procs:
        p.join(timeout=60)
    for r in results.values():
        assert not isinstance(r[1], str) or r[1] == "gone", r[1]
    # both replicas hold identical parameters after continuing (DP all-reduce over the rebuilt groups)
    a = {**results[0][1], **results[1][1]}
    b = results[2][1]
    assert sorted(a) == sorted(b) == [0, 1, 2, 3]
    for lid in a:
        torch.testing.assert_close(torch.from_numpy(a[lid]), torch.from_numpy(b[lid]), rtol=1e-6, atol=1e-8)
    print("reconfiguration seconds per rank:", {k: v[2] for k, v in results.items()})


def test_lone_pipeline_loss_raises():
    import types

    from oobleck_b200.execution.engine import ReconfigurationEngine
    from oobleck_b200.planning.pipeline_template import even_template
    t4, t3 = even_template(6, 4), even_template(6, 3)
    eng = types.SimpleNamespace(_pipeline_templates=[t3, t4], _agent_pipe=None, _num_gpus_per_node=1)
    re = ReconfigurationEngine.__new__(ReconfigurationEngine)
    re._engine = lambda: eng
    old = [t4.get_rank_grid([0, 1, 2, 3])]
    new = [t3.get_rank_grid([0, 1, 2])]
    with pytest.raises(RuntimeError, match="No alive ranks for the layer"):
        re._copy_model_states(old, new, None)
