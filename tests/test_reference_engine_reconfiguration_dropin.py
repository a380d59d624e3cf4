"""A lost node handled by the REFERENCE's own reconfiguration code on top of this package's execution layer.

The reference's test of this path (tests/execution/test_engine.py:887-1053, ``test_distribued_engine_reconfiguration``)
builds its engine on four workers under the fake agent, lets the fourth leave, and calls
``engine._reconfiguration._on_receive_reconfiguration_notification()`` on the survivors by hand (the listener thread is
patched out: the reference rebuilds pipelines from that thread with no lock against the training thread).  Same script
here, with everything of the reference imported unmodified (the bindings of tests/test_reference_engine_dropin.py):

    ReconfigurationEngine._on_receive_reconfiguration_notification   lost IP off the pipe, remove_lost_node_from_dist_info
    OobleckEngine.initialize_distributed                             the old world is DESTROYED, a new TCPStore / world of
                                                                     three comes up through the agent's port round trip
    ReconfigurationEngine.on_reconfigure / _reinstantiate            re-planning policy, get_new_execution_plan, instantiate,
                                                                     initialize_execution(model, existing_pipeline)
    ReconfigurationEngine._copy_model_states                         broadcasts straight into ``layer._param_handle.flat_param``
    DataParallelEngine (new groups on the new world)

and, underneath, this package's ``OobleckPipeline`` re-using its layers through ``create_layer_from_layer``.  The test then
goes further than the reference's: the moved parameters are compared, and the survivors take a training step on the rebuilt
pipelines that must land where a never-failed single-process run lands (wherever the optimizer state survived the move).
"""
import os
import sys
import threading

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_reference_engine_dropin as base  # noqa: E402
from test_pipeline_gloo import MARGS  # noqa: E402

# three steps first: WarmupLR publishes lr = 0 for the first two.  ONE step afterwards: the reference moves parameters
# without their Adam moments, so the moved copies (rank 2's layers 2 and 3) take a different first update than their
# replicas on rank 1; from the second step on their gradients differ and every replica drifts off the never-failed run.
WORLD, M, STEPS_BEFORE, STEPS_AFTER = 4, 12, 3, 1
IPS = ["127.0.0.1", "127.0.0.2", "127.0.0.3", "127.0.0.4"]


def process(rank, pipe, q):
    torch.set_num_threads(1)
    try:
        from unittest.mock import patch

        import torch.distributed as dist
        from oracle_layer import OracleLayer
        OracleLayer.reload_every_forward = True       # the reference writes flat_param behind the layer's back (see there)
        worker, tu, engine_mod = base.bind_reference()
        patch("socket.gethostbyname", return_value=IPS[rank]).start()
        real_tcpstore = torch.distributed.TCPStore
        patch("torch.distributed.TCPStore", lambda host_name, *a, **kw: real_tcpstore("127.0.0.1", *a, **kw)).start()
        # the test calls the reconfiguration by hand, like the reference's (test_engine.py:947-951)
        patch.object(engine_mod.ReconfigurationEngine, "_reconfiguration_listener_fn", lambda self: None).start()

        args = tu.OobleckArguments(
            dist=tu.DistributedArguments(master_ip="127.0.0.1", master_port=0, node_ips=list(IPS)),
            job=tu.JobArguments(microbatch_size=1, global_microbatch_size=M, steps=STEPS_BEFORE + STEPS_AFTER),
            model=tu.ModelArguments(model_name="gpt2", model_tag=base.TAG, dataset_path="synthetic",
                                    model_args=dict(MARGS)))
        engine = engine_mod.OobleckEngine(0, WORLD, 1, pipe, args)
        engine.initialize_distributed()
        engine.instantiate_pipelines(M)
        assert engine._dp_engine and engine._reconfiguration
        assert engine._dist_info.agent_ips == IPS and engine._dist_info.world_size == 4
        assert dist.get_world_size() == 4
        before = [p._ranks for p in engine._reconfiguration._pipelines]
        for _ in range(STEPS_BEFORE):
            engine._train_step()
        if rank == 3:
            q.put((rank, "left", (before, None), None))
            return                                                        # test_engine.py:960-961

        engine._reconfiguration._on_receive_reconfiguration_notification()

        assert engine._dist_info.agent_ips == IPS[:3] and engine._dist_info.world_size == 3    # :1021-1022
        assert dist.get_world_size() == 3 and dist.get_rank() == rank
        after = [p._ranks for p in engine._reconfiguration._pipelines]
        assert sorted(r for p in after for r in p) == [0, 1, 2]
        model_layers = len(engine._model.layers)
        for pipeline in engine._reconfiguration._pipelines:               # :978-993, generalised to any template
            assert sorted(pipeline.rank_grid) == list(range(model_layers))
            assert all(len(ranks) == 1 and ranks[0] in pipeline._ranks for ranks in pipeline.rank_grid.values())
        for layer_id, ranks_per_layer in engine._pipeline.rank_grid.items():    # :995-1011
            mine = [l for l in engine._pipeline.execution._layers if l.layer_id == layer_id]
            if rank in ranks_per_layer:
                assert len(mine) == 1 and mine[0]._param_handle.flat_param is not None
                assert mine[0]._param_handle.world_size == len(ranks_per_layer)
            else:
                assert not mine

        layers = engine._pipeline.execution._layers
        post_copy = {l.layer_id: (l.flat_param.numpy().copy(), bool(l.exp_avg.abs().max() > 0)) for l in layers}
        for _ in range(STEPS_AFTER):
            engine._train_step()
        out = {l.layer_id: l.flat_param.numpy().copy() for l in layers}
        q.put((rank, out, (before, after), post_copy))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None, None))
        raise


def never_failed_reference():
    """Single process: STEPS_BEFORE steps on global batches 0.., then -- the rebuilt loaders start a fresh iterator, the
    optimizer of a rebuilt ``PipelineExecution`` a fresh LR schedule position of the step count it was given -- the rest."""
    from oobleck_b200.execution.dataloader import OobleckSampler, SyntheticTokenDataset
    from oobleck_b200.module.model import OobleckModel
    from oracle import gpt2 as og
    from oracle import optim as oo
    model = OobleckModel("gpt2", {"input_ids": None, "attention_mask": None, "labels": None}, None, "t", dict(MARGS))
    layers = og.build_layers(og.GPT2Dims(n_embd=64, n_head=1, n_layer=2, n_positions=32, vocab_size=211))
    flats = [spec.init_flat() for spec in model.layers]
    for l, f in zip(layers, flats):
        og.load_flat_(l, f)
    ds = SyntheticTokenDataset(num_samples=128, seq_len=32, vocab_size=211, pin_memory=False)
    ms, vs = [torch.zeros_like(f) for f in flats], [torch.zeros_like(f) for f in flats]
    lrs = oo.lr_sequence(STEPS_BEFORE + STEPS_AFTER, warmup_min_lr=0)
    step = 0
    snapshots = []
    for nsteps in (STEPS_BEFORE, STEPS_AFTER):
        it = iter(OobleckSampler(ds, 1, 0, [M], 0))
        for _ in range(nsteps):
            for l in layers:
                l.zero_grad()
            for _ in range(M):
                ids = ds.input_ids[next(it)]
                x = (ids, torch.ones_like(ids), ids)
                for l in layers:
                    x = l(*x)
                x[0].backward()
            grads = [og.flat_grads(l) for l in layers]
            for i, l in enumerate(layers):
                oo.adamw_step_(flats[i], grads[i], ms[i], vs[i], step + 1, lrs[step])
                og.load_flat_(l, flats[i])
            step += 1
        snapshots.append([f.clone() for f in flats])
    return snapshots      # parameters after the steps before the loss, and at the end


@pytest.mark.timeout(400)
def test_reference_reconfiguration_code_rebuilds_this_packages_pipelines():
    if not base.available():
        pytest.skip("needs /root/reference and oracle/_ref (make -C oracle)")
    from oobleck_b200.execution.engine import DistributionInfo
    # cross-replica all-reduce that gets expensive beyond two replicas: get_best_execution_plan then prefers two 2-stage
    # pipelines to four single-stage ones, the layout of the reference's own test (test_engine.py:975-993)
    base.write_profile_files({1: 1e-3, 2: 2e-3, 3: 50.0, 4: 100.0})
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pipes = [ctx.Pipe(duplex=True) for _ in range(WORLD)]
    procs = [ctx.Process(target=process, args=(r, pipes[r][1], q)) for r in range(WORLD)]
    for p in procs:
        p.start()

    def rebroadcast(ps):
        port = ps[0][0].recv()
        for pipe, _ in ps:
            pipe.send(port)

    def agent():                                                          # test_engine.py:1034-1051
        for pipe, _ in pipes:
            pipe.send(DistributionInfo(list(IPS), WORLD))
        rebroadcast(pipes)
        for pipe, _ in pipes[:3]:
            pipe.send(IPS[3])                                             # "we lost node 3"
        rebroadcast(pipes[:3])

    t = threading.Thread(target=agent, daemon=True)
    t.start()
    results = {}
    for _ in range(WORLD):
        r = q.get(timeout=300)
        results[r[0]] = r
    t.join(timeout=60)
    for p in procs:
        p.join(timeout=60)
    assert results[3][1] == "left"
    for r in range(3):
        assert not isinstance(results[r][1], str), results[r][1]
    before, after = results[0][2]
    print("pipelines before / after (the reference's planner and policy):", before, "->", after)
    assert before == [[0, 1], [2, 3]] and after == [[2], [0, 1]]          # "one 1-stage pipeline and one 2-stage pipeline"
    at_loss, at_end = never_failed_reference()
    # 1. right after the reference's _copy_model_states: every layer a survivor holds carries the trained parameters --
    #    rank 2 now owns the whole model, layers 2 and 3 arrived by broadcast from rank 1 (without their Adam moments: the
    #    reference moves flat_param only, engine.py:284-306)
    assert sorted(results[2][3]) == [0, 1, 2, 3] and sorted(results[0][3]) == [0, 1] and sorted(results[1][3]) == [2, 3]
    for r in range(3):
        for lid, (param, has_moments) in results[r][3].items():
            torch.testing.assert_close(torch.from_numpy(param), at_loss[lid], rtol=1e-5, atol=1e-7)
            assert has_moments == (not (r == 2 and lid in (2, 3)))
    # 2. a training step on the rebuilt pipelines: wherever the optimizer state survived, the result is the never-failed
    #    run's (the copies that restarted their moments take a different update -- the reference's semantics)
    for r in range(3):
        for lid, f in results[r][1].items():
            assert bool(torch.isfinite(torch.from_numpy(f)).all())
            if not (r == 2 and lid in (2, 3)):
                torch.testing.assert_close(torch.from_numpy(f), at_end[lid], rtol=1e-4, atol=2e-6)
