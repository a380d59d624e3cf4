"""The engine driven purely through the agent pipe, like ``worker_main`` (oobleck/elastic/worker.py:13-34) drives the
reference's, with the reference's own fake-agent harness (tests/execution/test_engine.py:650-657, 1037-1053): a thread
that sends ``DistributionInfo`` down every worker pipe, re-broadcasts rank 0's TCPStore port, later announces a lost
node IP to the survivors and re-broadcasts the port again.

4 workers (gloo / CPU, oracle layers): 2 replicas x 2 stages; after two steps the process of rank 3 dies without a
word (``os._exit``), the agent announces its IP, the listener threads queue the reconfiguration, the training threads
drop the step in flight, re-plan with the reference's policy ([[2], [0, 1]]), move the missing layers INCLUDING the
Adam moments, and train on.  The parameters after the last step must match a never-failed single-process oracle run
that consumes the same global batches (the sampler restarts its epoch on reconfiguration, exactly like the
reference's: dataloader.py:43-100 ignores ``num_iterations_done``)."""
import os
import socket
import sys
import threading
from unittest.mock import patch

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
MARGS = dict(n_embd=64, n_head=1, num_hidden_layers=2, n_positions=32, vocab_size=211)
AGENT_IPS = ["127.0.0.1", "127.0.0.2", "127.0.0.3", "127.0.0.4"]
M, MB, STEPS_BEFORE, STEPS_TOTAL = 4, 1, 2, 4
# gloo's per-operation timeout.  Nothing in a passing run waits for it (a dead peer closes its sockets: the survivors'
# operations fail at once, and the listener's abort releases everything else); it only decides how long a merely SLOW
# neighbour -- a loaded CI host, eight single-threaded workers on fewer cores -- is waited for before the step is given up.
COMM_TIMEOUT_S = 90


def scenario(mode):
    """(world, model args, global micro-batches, templates factory, initial plan, pipelines expected after the loss)"""
    from oobleck_b200.planning.pipeline_template import even_template
    mode = mode.split("+")[0]
    if mode == "replicas":
        t = [even_template(4, 1), even_template(4, 2)]
        return 4, MARGS, 4, t, [t[1], t[1]], [[2], [0, 1]]
    if mode == "lone":        # one 4-stage pipeline without a replica: survives through the peer shadows
        t = [even_template(4, 3), even_template(4, 4)]
        return 4, MARGS, 4, t, [t[1]], [[0, 1, 2]]
    big = dict(MARGS, num_hidden_layers=6)          # 8 stage layers
    if mode == "replicas8":   # BASELINE config 4's shape: 2 replicas x 4 stages on 8 ranks -> 4 + 3
        t = [even_template(8, 3), even_template(8, 4)]
        return 8, big, 8, t, [t[1], t[1]], [[4, 5, 6], [0, 1, 2, 3]]
    if mode == "lone8":       # BASELINE config 5's shape: one 8-stage pipeline -> the 7-stage template
        t = [even_template(8, 7), even_template(8, 8)]
        return 8, big, 8, t, [t[1]], [[0, 1, 2, 3, 4, 5, 6]]
    raise ValueError(mode)


def ips_of(world):
    return [f"127.0.0.{i + 1}" for i in range(world)]


def worker(rank, pipe, q, ready, mode):
    torch.set_num_threads(1)
    try:
        from oracle_layer import OracleLayer

        from oobleck_b200.execution.dataloader import SyntheticTokenDataset
        from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine
        from oobleck_b200.planning.pipeline_template import even_template
        world, margs, M, templates, plan, pipelines_after = scenario(mode)
        AGENT_IPS = ips_of(world)
        victim = world - 1
        # every worker believes it runs on its own node: the reference's tests patch the same call (test_engine.py:676)
        patch("socket.gethostbyname", return_value=AGENT_IPS[rank]).start()
        real_tcpstore = torch.distributed.TCPStore

        def local_store(host_name, *a, **kw):      # the fake node IPs all live on this host
            return real_tcpstore("127.0.0.1", *a, **kw)
        patch("torch.distributed.TCPStore", local_store).start()

        args = OobleckArguments(job=JobArguments(microbatch_size=MB, global_microbatch_size=MB * M, steps=STEPS_TOTAL),
                                model=ModelArguments(model_name="gpt2", model_tag="t", model_args=dict(margs)))
        ds = SyntheticTokenDataset(num_samples=256, seq_len=32, vocab_size=211, pin_memory=False)
        # worker_main's call sequence: ctor(local_rank, num_nodes, gpus_per_node, pipe, args) -> initialize_distributed
        # -> instantiate_pipelines -> train
        eng = OobleckEngine(0, len(AGENT_IPS), 1, pipe, args, dataset=ds, layer_cls=OracleLayer, templates=templates,
                            backend="gloo", comm_timeout_s=COMM_TIMEOUT_S, peer_shadow=mode.startswith("lone"))
        eng.initialize_distributed()
        assert eng._rank == rank and eng._world_size == world and eng._rank_map[AGENT_IPS[rank]] == [rank]
        eng.instantiate_pipelines(M, plan=plan)
        assert eng._reconfiguration._reconfiguration_listener is not None    # engine.py:50-53

        orig_step = eng._guarded_train_step
        count = {"n": 0}

        def step_hook():
            if count["n"] == STEPS_BEFORE:
                if rank == victim:
                    def die(*_a, **_k):
                        q.put((rank, "gone", None, None))
                        q.close(); q.join_thread()      # noqa: E702
                        os._exit(0)                     # the node dies: no goodbye, no barrier
                    if mode.endswith("+allreduce"):
                        # dies AFTER the first vote has passed, inside the gradient exchange: only its all-reduce
                        # partner sees a failed collective; the second vote must keep everybody else from committing
                        eng._dp_engine.do_allreduce = die
                    else:
                        die()
                else:
                    ready.put(rank)                     # survivors: tell the "agent" that the failure may be announced
            count["n"] += 1
            return orig_step()
        eng._guarded_train_step = step_hook
        eng.train()

        new_ranks = [p._ranks for p in eng._reconfiguration._pipelines]
        assert new_ranks == pipelines_after, new_ranks
        assert eng._dist_info.agent_ips == AGENT_IPS[:victim] and eng._dist_info.world_size == world - 1
        layers = eng._pipeline.execution._layers
        # ownership follows the template of the pipeline this rank ended up in
        mine = next(p for p in eng._reconfiguration._pipelines if rank in p._ranks)
        stage = mine._template.get_stages()[mine._ranks.index(rank)]
        assert sorted(l.layer_id for l in layers) == list(stage._layer_indices)
        out = {l.layer_id: (l.flat_param.numpy().copy(), l.exp_avg.numpy().copy(), l.opt_step) for l in layers}
        q.put((rank, out, eng._reconfiguration.last_reconfiguration_seconds, len(eng.step_seconds)))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None, None))
        raise


def never_failed_reference(margs=MARGS, M=M, grad_scale=1.0):
    """Single process, all layers: steps 0..STEPS_BEFORE-1 on global batches 0.., then the sampler restarts (the new
    loaders start a fresh epoch-0 iterator) and the remaining steps consume global batches 0.. again."""
    from oobleck_b200.execution.dataloader import OobleckSampler, SyntheticTokenDataset
    from oobleck_b200.module.model import OobleckModel
    from oracle import gpt2 as og
    from oracle import optim as oo
    model = OobleckModel("gpt2", {"input_ids": None, "attention_mask": None, "labels": None}, None, "t", dict(margs))
    layers = og.build_layers(og.GPT2Dims(n_embd=64, n_head=1, n_layer=margs["num_hidden_layers"], n_positions=32,
                                         vocab_size=211))
    flats = [spec.init_flat() for spec in model.layers]
    for l, f in zip(layers, flats):
        og.load_flat_(l, f)
    ds = SyntheticTokenDataset(num_samples=256, seq_len=32, vocab_size=211, pin_memory=False)
    ms, vs = [torch.zeros_like(f) for f in flats], [torch.zeros_like(f) for f in flats]
    lrs = oo.lr_sequence(STEPS_TOTAL, warmup_min_lr=0)
    step = 0
    for nsteps in (STEPS_BEFORE, STEPS_TOTAL - STEPS_BEFORE):
        it = iter(OobleckSampler(ds, MB, 0, [M], 0))           # one pipeline taking the whole global batch
        for _ in range(nsteps):
            for l in layers:
                l.zero_grad()
            for _ in range(M):
                ids = ds.input_ids[next(it)]
                x = (ids, torch.ones_like(ids), ids)
                for l in layers:
                    x = l(*x)
                x[0].backward()
            grads = [og.flat_grads(l) * grad_scale for l in layers]
            for i, l in enumerate(layers):
                oo.adamw_step_(flats[i], grads[i], ms[i], vs[i], step + 1, lrs[step])
                og.load_flat_(l, flats[i])
            step += 1
    return flats, ms


@pytest.mark.timeout(600)
@pytest.mark.parametrize("mode", ["replicas", "lone", "replicas8", "lone8", "replicas+allreduce"])
def test_engine_driven_through_agent_pipe_survives_a_dead_node(mode):
    """``lone``: the same death in a single 4-stage pipeline.  The reference raises "No alive ranks for the layer"
    (engine.py:263-269, its test at tests/execution/test_engine.py:1015-1019); with peer shadows the survivors re-split
    into 3 stages -- layers 1 and 2 move between survivors, layer 3 (parameters AND Adam moments) comes out of the mirror
    its neighbour kept -- and the run still matches the never-failed oracle."""
    from oobleck_b200.execution.engine import DistributionInfo
    world, margs, M_, _, _, _ = scenario(mode)
    AGENT_IPS = ips_of(world)
    victim = world - 1
    ctx = mp.get_context("spawn")
    q, ready = ctx.Queue(), ctx.Queue()
    pipes = [ctx.Pipe(duplex=True) for _ in AGENT_IPS]
    procs = [ctx.Process(target=worker, args=(r, pipes[r][1], q, ready, mode)) for r in range(world)]
    for p in procs:
        p.start()

    def broadcast_rank0_port(ps):                      # tests/execution/test_engine.py:650-657
        port = ps[0][0].recv()
        for pipe, _ in ps:
            pipe.send(port)

    def agent():
        for pipe, _ in pipes:
            pipe.send(DistributionInfo(list(AGENT_IPS), len(AGENT_IPS)))
        broadcast_rank0_port(pipes)
        for _ in range(world - 1):                     # the survivors are inside the step the dead node misses
            ready.get(timeout=300)
        procs[victim].join(timeout=60)                 # the node is really gone
        for pipe, _ in pipes[:victim]:
            pipe.send(AGENT_IPS[victim])               # test_engine.py:1045-1047
        broadcast_rank0_port(pipes[:victim])

    t = threading.Thread(target=agent, daemon=True)
    t.start()
    results = {}
    try:
        for _ in range(world):
            r = q.get(timeout=500)
            results[r[0]] = r
        t.join(timeout=60)
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:              # never leave workers behind: they would load the host for the tests that follow
            if p.is_alive():
                p.terminate()
    assert results[victim][1] == "gone"
    survivors = range(victim)
    for r in survivors:
        assert not isinstance(results[r][1], str), results[r][1]
        assert results[r][3] == STEPS_TOTAL            # every survivor completed all steps (one of them twice started)
        assert results[r][2] is not None and results[r][2] < 60
    flats, ms = never_failed_reference(margs, M_)
    for r in survivors:
        for lid, (param, exp_avg, opt_step) in results[r][1].items():
            assert opt_step == STEPS_TOTAL             # moved layers brought their AdamW step count along
            # summation order differs (replicas reduce, one process accumulates): Adam turns 1e-8 gradient noise into ~1e-7 steps
            torch.testing.assert_close(torch.from_numpy(param), flats[lid], rtol=1e-4, atol=2e-6)
            torch.testing.assert_close(torch.from_numpy(exp_avg), ms[lid], rtol=1e-3, atol=1e-7)
    print("reconfiguration seconds (notification -> pipelines rebuilt):", {r: results[r][2] for r in survivors})
