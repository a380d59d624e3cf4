"""BASELINE config 1: GPT-2 2-stage pipeline, world_size 2, CPU + gloo -- the plumbing (1F1B interpreter, the
reference's blocking wire protocol incl. the meta handshake, buffer lifecycle) with the oracle as stage compute,
checked against a single-process run of the same model.  Also the 2-replica DP all-reduce path (config 4 shape)."""
import itertools
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MARGS = dict(n_embd=64, n_head=1, num_hidden_layers=2, n_positions=32, vocab_size=211)


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def make_engine(rank, world, num_gpus_per_node, M, mb, steps, templates=None):
    from oracle_layer import OracleLayer

    from oobleck_b200.execution.dataloader import SyntheticTokenDataset
    from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine
    args = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=mb * M, steps=steps),
                            model=ModelArguments(model_name="gpt2", model_tag="t", model_args=dict(MARGS)))
    ds = SyntheticTokenDataset(num_samples=128, seq_len=32, vocab_size=211, pin_memory=False)
    eng = OobleckEngine(rank, world, num_gpus_per_node, None, args, dataset=ds, layer_cls=OracleLayer,
                        templates=templates)
    eng.initialize_distributed("gloo")
    return eng


def reference_run(M, mb, steps, num_pipelines=1):
    """single process: all layers, same batches in the sampler's order, grads summed over every pipeline."""
    from oobleck_b200.execution.dataloader import SyntheticTokenDataset
    from oobleck_b200.execution.dataloader import OobleckSampler
    from oobleck_b200.module.model import OobleckModel
    from oracle import gpt2 as og
    from oracle import optim as oo
    model = OobleckModel("gpt2", {"input_ids": None, "attention_mask": None, "labels": None}, None, "t", dict(MARGS))
    d = og.GPT2Dims(n_embd=64, n_head=1, n_layer=2, n_positions=32, vocab_size=211)
    layers = og.build_layers(d)
    flats = [spec.init_flat() for spec in model.layers]
    for l, f in zip(layers, flats):
        og.load_flat_(l, f)
    ds = SyntheticTokenDataset(num_samples=128, seq_len=32, vocab_size=211, pin_memory=False)
    per = M // num_pipelines
    iters = [iter(OobleckSampler(ds, mb, pi, [per] * num_pipelines, 0)) for pi in range(num_pipelines)]
    ms, vs = [torch.zeros_like(f) for f in flats], [torch.zeros_like(f) for f in flats]
    lrs = oo.lr_sequence(steps, warmup_min_lr=0)
    losses = []
    for step in range(steps):
        for l in layers:
            l.zero_grad()
        tot = [0.0] * num_pipelines
        for pi in range(num_pipelines):
            for _ in range(per):
                ids = ds.input_ids[next(iters[pi])]
                x = (ids, torch.ones_like(ids), ids)
                for l in layers:
                    x = l(*x)
                x[0].backward()
                tot[pi] += float(x[0].detach())
        losses.append(tot)
        grads = [og.flat_grads(l) for l in layers]
        for i, l in enumerate(layers):
            oo.adamw_step_(flats[i], grads[i], ms[i], vs[i], step + 1, lrs[step])
            og.load_flat_(l, flats[i])
    return flats, losses, grads


def worker_pp2(rank, world, port, M, mb, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    try:
        eng = make_engine(rank, world, 1, M, mb, steps)
        eng.instantiate_pipelines(M)
        pipe = eng._pipeline
        # tests/execution/test_pipeline.py:193-198 neighbours
        assert pipe.communication.prev_rank == (None if rank == 0 else rank - 1)
        assert pipe.communication.next_rank == (None if rank == world - 1 else rank + 1)
        assert pipe.is_first_stage() == (rank == 0) and pipe.is_last_stage() == (rank == world - 1)
        totals = []
        for _ in range(steps):
            eng._train_step()
            assert all(b is None for bufs in pipe.pipe_buffers.values() for b in bufs)   # test_pipeline.py:344-350
            totals.append(float(pipe.execution.total_loss) if pipe.is_last_stage() else None)
        assert pipe._global_step == steps
        if rank > 0:
            assert pipe.communication.activation_recv_buf is not None
        if rank < world - 1:
            assert pipe.communication.sent_activation_meta and pipe.communication.grad_recv_buf is not None
        out = {l.layer_id: l.flat_param.numpy().copy() for l in pipe.execution._layers}
        q.put((rank, out, totals, None))
        dist.barrier()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


RENDEZVOUS_TROUBLE = ("Address already in use", "EADDRINUSE", "Connection refused", "Connection reset by peer")


def run_spawn(fn, world, *a):
    """One retry, on a fresh port, if the rendezvous itself failed (``free_port`` releases the port before rank 0's
    store binds it: another process can take it in between).  Anything else fails the test."""
    for attempt in range(2):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = free_port()
        procs = [ctx.Process(target=fn, args=(r, world, port, *a, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            results = [q.get(timeout=240) for _ in range(world)]
            for p in procs:
                p.join(timeout=60)
        finally:
            for p in procs:          # never leave workers behind: they would load the host for the tests that follow
                if p.is_alive():
                    p.terminate()
        errors = [r[3] for r in results if r[3] is not None]
        if errors and attempt == 0 and any("init_process_group" in e and any(t in e for t in RENDEZVOUS_TROUBLE)
                                           for e in errors):
            continue
        for r in results:
            assert r[3] is None, r[3]
        return sorted(results, key=lambda r: r[0])


@pytest.mark.timeout(300)
def test_two_stage_pipeline_gloo_matches_single_process():
    M, mb, steps = 4, 1, 3
    results = run_spawn(worker_pp2, 2, M, mb, steps)
    flats, losses, _ = reference_run(M, mb, steps)
    seen = {}
    for _, out, totals, _ in results:
        seen.update(out)
    assert sorted(seen) == [0, 1, 2, 3]                       # layer counts sum to the model (test_pipeline.py:213)
    for lid, f in seen.items():
        torch.testing.assert_close(torch.from_numpy(f), flats[lid], rtol=1e-5, atol=1e-7)
    totals = results[-1][2]
    ref_cum = list(itertools.accumulate(sum(t) for t in losses))
    for got, want in zip(totals, ref_cum):
        assert abs(got - want) < 1e-5 * abs(want)


def worker_dp2(rank, world, port, M, mb, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    try:
        from oobleck_b200.planning.pipeline_template import even_template
        eng = make_engine(rank, world, 1, M, mb, steps, templates=[even_template(4, 1)])
        eng.instantiate_pipelines(M)      # two single-stage replicas
        assert len(eng._dp_engine.group_creation_order) == 1 and eng._dp_engine.group_creation_order[0] == [0, 1]
        for _ in range(steps):
            eng._train_step()
        out = {l.layer_id: l.flat_param.numpy().copy() for l in eng._pipeline.execution._layers}
        q.put((rank, out, float(eng._pipeline.execution.total_loss), None))
        dist.barrier()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


@pytest.mark.timeout(300)
def test_two_replicas_dp_allreduce_gloo():
    """SUM (not mean) all-reduce of every layer's flat gradient across replicas (layer.py:290-291)."""
    M, mb, steps = 4, 1, 2
    results = run_spawn(worker_dp2, 2, M, mb, steps)
    flats, losses, _ = reference_run(M, mb, steps, num_pipelines=2)
    for rank, out, total, _ in results:
        for lid, f in out.items():
            torch.testing.assert_close(torch.from_numpy(f), flats[lid], rtol=1e-5, atol=1e-7)
        want = sum(t[rank] for t in losses)
        assert abs(total - want) < 1e-5 * abs(want)


def worker_dp2_pp4(rank, world, port, M, mb, steps, q):
    """The construction bench.py uses for ``--gpus 8 --replicas 2`` (BASELINE config 4): the engine is told the
    pipeline depth (4 "nodes"), the world size says how many replicas fit."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    try:
        stages = world // 2
        eng = make_engine(rank, stages, 1, M, mb, steps)
        eng._world_size = world
        eng.initialize_distributed("gloo")
        eng.instantiate_pipelines(M)
        pipes = eng._reconfiguration._pipelines
        assert [p._ranks for p in pipes] == [list(range(stages)), list(range(stages, world))]
        # one cross-replica communicator per stage position, created in the same order on every rank (engine.py:374-392)
        assert eng._dp_engine.group_creation_order == [[s, s + stages] for s in range(stages)]
        sampler = eng._pipeline._dataloader.batch_sampler
        assert sampler.num_microbatches == [M // 2, M // 2] and sampler.pipeline_index == rank // stages
        pipe = eng._pipeline
        me = rank % stages
        assert pipe.communication.prev_rank == (None if me == 0 else rank - 1)
        assert pipe.communication.next_rank == (None if me == stages - 1 else rank + 1)
        for _ in range(steps):
            eng._train_step()
        out = {l.layer_id: l.flat_param.numpy().copy() for l in pipe.execution._layers}
        q.put((rank, out, float(pipe.execution.total_loss) if pipe.is_last_stage() else None, None))
        dist.barrier()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


@pytest.mark.timeout(400)
def test_two_replicas_of_four_stages_on_eight_ranks():
    """BASELINE config 4's shape, built the way bench.py builds it, against the single-process oracle run."""
    M, mb, steps = 8, 1, 2
    results = run_spawn(worker_dp2_pp4, 8, M, mb, steps)
    flats, losses, _ = reference_run(M, mb, steps, num_pipelines=2)
    seen = {}
    for rank, out, total, _ in results:
        for lid, f in out.items():
            torch.testing.assert_close(torch.from_numpy(f), flats[lid], rtol=1e-5, atol=1e-7)
            seen.setdefault(lid, []).append(rank)
        if total is not None:
            want = sum(t[rank // 4] for t in losses)
            assert abs(total - want) < 1e-5 * abs(want)
    assert sorted(seen) == [0, 1, 2, 3] and all(len(r) == 2 for r in seen.values())   # every layer lives on two ranks
