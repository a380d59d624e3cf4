"""Two nodes lost one after the other (``JobArguments.fault_threshold`` exists because Oobleck expects more than one
failure per job): 8 workers driven through the agent pipe with the reference's fake-agent harness
(tests/execution/test_engine.py:650-657, 1037-1053), gloo / CPU, oracle layers.

    2 replicas x 4 stages   [0,1,2,3] [4,5,6,7]
    rank 7 dies in step 2   -> [4,5,6] [0,1,2,3]          (policy: the reference's, engine.py:91-166)
    rank 6 dies in step 4   -> [0,1,2] [3,4,5]            ([4,5] is below the smallest template: it borrows rank 3 from
                                                           the biggest pipeline, engine.py:117-131)

The second loss re-splits BOTH pipelines, so for the last layers nobody "keeps the layer in place": the reference gives up
there (``RuntimeError("No alive ranks for the layer")``, engine.py:263-269) although rank 3 still holds them.  With peer
shadows enabled the engine takes a surviving old owner as the source (``_move_without_replica``).  Parameters, Adam
moments and step counts after the last step must match a never-failed single-process oracle run that consumes the same
global batches (every reconfiguration restarts the sampler's epoch, like the reference's: dataloader.py:43-100)."""
import os
import sys
import threading
from unittest.mock import patch

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
MARGS = dict(n_embd=64, n_head=1, num_hidden_layers=6, n_positions=32, vocab_size=211)     # 8 stage layers
WORLD, M, MB = 8, 8, 1
LOSSES = [(2, 7), (4, 6)]             # (training step in which the node dies, its rank)
STEPS_TOTAL = 6
AFTER = [[[4, 5, 6], [0, 1, 2, 3]], [[0, 1, 2], [3, 4, 5]]]


def ips_of(world):
    return [f"127.0.0.{i + 1}" for i in range(world)]


def worker(rank, pipe, q, ready):
    torch.set_num_threads(1)
    try:
        from oracle_layer import OracleLayer

        from oobleck_b200.execution.dataloader import SyntheticTokenDataset
        from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine
        from oobleck_b200.planning.pipeline_template import even_template
        ips = ips_of(WORLD)
        patch("socket.gethostbyname", return_value=ips[rank]).start()       # test_engine.py:676
        real_tcpstore = torch.distributed.TCPStore
        patch("torch.distributed.TCPStore", lambda host_name, *a, **kw: real_tcpstore("127.0.0.1", *a, **kw)).start()
        templates = [even_template(8, 3), even_template(8, 4)]
        args = OobleckArguments(job=JobArguments(microbatch_size=MB, global_microbatch_size=MB * M, steps=STEPS_TOTAL),
                                model=ModelArguments(model_name="gpt2", model_tag="t", model_args=dict(MARGS)))
        ds = SyntheticTokenDataset(num_samples=256, seq_len=32, vocab_size=211, pin_memory=False)
        eng = OobleckEngine(0, WORLD, 1, pipe, args, dataset=ds, layer_cls=OracleLayer, templates=templates,
                            backend="gloo", comm_timeout_s=90, peer_shadow=True)
        eng.initialize_distributed()
        eng.instantiate_pipelines(M, plan=[templates[1], templates[1]])
        orig_step = eng._guarded_train_step
        seen = {"announced": 0}

        def step_hook():
            done = len(eng.step_seconds)                    # completed steps so far = index of the step about to run
            for i, (at, victim) in enumerate(LOSSES):
                if done == at and seen["announced"] == i:
                    seen["announced"] = i + 1
                    if rank == victim:
                        q.put((rank, "gone", None))
                        q.close(); q.join_thread()          # noqa: E702
                        os._exit(0)
                    ready.put((i, rank))
            return orig_step()
        eng._guarded_train_step = step_hook
        eng.train()
        layers = eng._pipeline.execution._layers
        out = {l.layer_id: (l.flat_param.numpy().copy(), l.exp_avg.numpy().copy(), l.opt_step) for l in layers}
        q.put((rank, out, [p._ranks for p in eng._reconfiguration._pipelines]))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None))
        raise


def never_failed_reference():
    from oobleck_b200.execution.dataloader import OobleckSampler, SyntheticTokenDataset
    from oobleck_b200.module.model import OobleckModel
    from oracle import gpt2 as og
    from oracle import optim as oo
    model = OobleckModel("gpt2", {"input_ids": None, "attention_mask": None, "labels": None}, None, "t", dict(MARGS))
    layers = og.build_layers(og.GPT2Dims(n_embd=64, n_head=1, n_layer=6, n_positions=32, vocab_size=211))
    flats = [spec.init_flat() for spec in model.layers]
    for l, f in zip(layers, flats):
        og.load_flat_(l, f)
    ds = SyntheticTokenDataset(num_samples=256, seq_len=32, vocab_size=211, pin_memory=False)
    ms, vs = [torch.zeros_like(f) for f in flats], [torch.zeros_like(f) for f in flats]
    lrs = oo.lr_sequence(STEPS_TOTAL, warmup_min_lr=0)
    bounds = [0] + [at for at, _ in LOSSES] + [STEPS_TOTAL]
    step = 0
    for lo, hi in zip(bounds, bounds[1:]):
        it = iter(OobleckSampler(ds, MB, 0, [M], 0))       # every rebuild starts a fresh epoch-0 iterator
        for _ in range(hi - lo):
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
    return flats, ms


@pytest.mark.timeout(900)
def test_two_nodes_lost_one_after_the_other():
    from oobleck_b200.execution.engine import DistributionInfo
    ips = ips_of(WORLD)
    ctx = mp.get_context("spawn")
    q, ready = ctx.Queue(), ctx.Queue()
    pipes = [ctx.Pipe(duplex=True) for _ in ips]
    procs = [ctx.Process(target=worker, args=(r, pipes[r][1], q, ready)) for r in range(WORLD)]
    for p in procs:
        p.start()

    def broadcast_rank0_port(ps):
        port = ps[0][0].recv()
        for pipe, _ in ps:
            pipe.send(port)

    def agent():
        for pipe, _ in pipes:
            pipe.send(DistributionInfo(list(ips), len(ips)))
        broadcast_rank0_port(pipes)
        alive = list(range(WORLD))
        for i, (_, victim) in enumerate(LOSSES):
            for _ in range(len(alive) - 1):                # the survivors are inside the step the dead node misses
                assert ready.get(timeout=400)[0] == i
            procs[victim].join(timeout=60)
            alive.remove(victim)
            for r in alive:
                pipes[r][0].send(ips[victim])
            broadcast_rank0_port([pipes[r] for r in alive])

    t = threading.Thread(target=agent, daemon=True)
    t.start()
    results = {}
    try:
        for _ in range(WORLD):
            r = q.get(timeout=800)
            results[r[0]] = r
        t.join(timeout=60)
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
    victims = [v for _, v in LOSSES]
    for v in victims:
        assert results[v][1] == "gone"
    survivors = [r for r in range(WORLD) if r not in victims]
    for r in survivors:
        assert not isinstance(results[r][1], str), results[r][1]
        assert results[r][2] == AFTER[-1], results[r][2]
    flats, ms = never_failed_reference()
    covered = set()
    for r in survivors:
        for lid, (param, exp_avg, opt_step) in results[r][1].items():
            covered.add(lid)
            assert opt_step == STEPS_TOTAL
            torch.testing.assert_close(torch.from_numpy(param), flats[lid], rtol=1e-4, atol=2e-6)
            torch.testing.assert_close(torch.from_numpy(exp_avg), ms[lid], rtol=1e-3, atol=1e-7)
    assert covered == set(range(8))
