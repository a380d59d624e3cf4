"""Intra-stage sharding (reference layer.py:96-225, 262-291, engine.py:363-412) on CPU + gloo.

The host logic under test is the product's (``oobleck_b200/execution/sharding.py``, the stage communicators of
``OobleckPipeline.initialize_distributed_fsdp``, ``DataParallelEngine``'s per-``fsdp_index`` groups); stage compute is the
oracle.  Expected values follow the reference's semantics: every shard column of a pipeline runs the pipeline's
micro-batches (one dataloader per pipeline, engine.py:618-628) and the post-backward reduce-scatter SUMs without
pre-division (layer.py:202-206), so a stage of ``k`` GPUs contributes ``k`` times its gradient."""
import os
import sys

import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_pipeline_gloo import MARGS, make_engine, run_spawn  # noqa: E402


class FakeGroup:
    def __init__(self, ranks, me, comm="comm"):
        self.ranks, self.me, self.group = ranks, me, comm

    def size(self):
        return len(self.ranks)

    def rank_index(self):
        return self.ranks.index(self.me)


def test_shard_param_is_the_reference_chunking():
    """layer.py:262-270: ``number`` chunks, missing ones zero-filled, the last zero-padded to the first's size."""
    from oobleck_b200.execution.sharding import shard_param
    t = torch.arange(10.0)
    got = shard_param(t, 4)                      # torch.chunk(4) of 10 -> 3, 3, 3, 1
    assert [c.tolist() for c in got] == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 0, 0]]
    got = shard_param(torch.arange(4.0).view(2, 2), 3)     # chunk(3) of 4 -> 2, 2 : one chunk is missing
    assert [c.tolist() for c in got] == [[0, 1], [2, 3], [0, 0]]
    got = shard_param(torch.arange(8.0), 2)
    assert [c.tolist() for c in got] == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_shard_layout_nests_across_stage_widths():
    """A layer held by 1, 2 or 4 ranks of a 4-column grid: shard boundaries are multiples of the same unit, so the
    cross-replica groups of column c reduce the same elements whatever the width of the stage on the other side."""
    from oobleck_b200.execution.sharding import ShardedFlatState, shard_unit
    n, columns = 1000, 4
    unit = shard_unit(n, columns)
    assert unit == 256 and unit % 8 == 0 and unit * columns >= n
    for k in (1, 2, 4):
        for index in range(k):
            st = ShardedFlatState(n, FakeGroup(list(range(k)), index), columns, "cpu")
            assert st.padded == 1024 and st.per_rank == 1024 // k and st.lo == index * st.per_rank
            assert st.compute_param.numel() == n and st.full_grad.numel() == 1024
            if k > 1:   # the shard is a view into the gathered buffer: optimizer and gather work in place
                st.param_shard.fill_(7.0)
                assert torch.equal(st.full_param[st.lo:st.hi], torch.full((st.per_rank,), 7.0))
                assert st.exp_avg.numel() == st.per_rank and st.grad_shard.numel() == st.per_rank
            else:
                assert st.param_shard.numel() == n and st.exp_avg.numel() == n
            # cross-replica groups: one per column this rank owns; the slices tile its reduce buffer
            first = index * (columns // k)
            groups = {c: FakeGroup([0, 9], 0, comm=f"c{c}") for c in range(first, first + columns // k)}
            chunks = st.dp_chunks(groups)
            assert [c.numel() for c, _ in chunks] == [unit] * (columns // k)
            base = st.reduce_buffer.data_ptr()
            assert [(c.data_ptr() - base) // 4 for c, _ in chunks] == [i * unit for i in range(columns // k)]
    # consecutive columns on ONE communicator are reduced in one call; single-member groups are skipped
    st = ShardedFlatState(n, FakeGroup([0], 0), columns, "cpu")
    same = {c: FakeGroup([0, 9], 0, comm="one") for c in range(4)}
    assert [c.numel() for c, _ in st.dp_chunks(same)] == [1024]
    lonely = {0: FakeGroup([0, 9], 0, comm="a"), 1: FakeGroup([0], 0, comm=None), 2: FakeGroup([0, 9], 0, comm="a"),
              3: FakeGroup([0, 9], 0, comm="a")}
    assert [c.numel() for c, _ in st.dp_chunks(lonely)] == [256, 512]
    with pytest.raises(ValueError):
        ShardedFlatState(n, FakeGroup([0, 1, 2], 0), 4, "cpu")            # 4 columns over 3 ranks
    # NO_SHARD allocates exactly the flat vector
    st = ShardedFlatState(n, FakeGroup([0], 0), 1, "cpu")
    assert st.padded == n and st.param_shard is st.full_param and st.param_shard.grad is st.full_grad


# ---- engine level --------------------------------------------------------------------------------------------------
def weighted_reference(M_per_pipeline, weights, mb, steps):
    """Single process, all layers.  Pipeline ``p`` consumes its sampler's micro-batches; its gradient enters the update
    ``weights[p]`` times (the width of its stages)."""
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
    iters = [iter(OobleckSampler(ds, mb, pi, list(M_per_pipeline), 0)) for pi in range(len(M_per_pipeline))]
    ms, vs = [torch.zeros_like(f) for f in flats], [torch.zeros_like(f) for f in flats]
    lrs = oo.lr_sequence(steps, warmup_min_lr=0)
    totals = [0.0] * len(M_per_pipeline)
    for step in range(steps):
        grads = [torch.zeros_like(f) for f in flats]
        for pi, m in enumerate(M_per_pipeline):
            for l in layers:
                l.zero_grad()
            for _ in range(m):
                ids = ds.input_ids[next(iters[pi])]
                x = (ids, torch.ones_like(ids), ids)
                for l in layers:
                    x = l(*x)
                x[0].backward()
                totals[pi] += float(x[0].detach())
            for g, l in zip(grads, layers):
                g.add_(og.flat_grads(l), alpha=float(weights[pi]))
        for i, l in enumerate(layers):
            oo.adamw_step_(flats[i], grads[i], ms[i], vs[i], step + 1, lrs[step])
            og.load_flat_(l, flats[i])
    return flats, ms, totals


def wide_template(num_layers, stages, gpus_per_stage, nodes, gpn):
    from oobleck_b200.planning.pipeline_template import PipelineTemplate, StageExecutionResult
    base, extra = divmod(num_layers, stages)
    out, start = [], 0
    for s in range(stages):
        n = base + (1 if s < extra else 0)
        out.append(StageExecutionResult(range(start, start + n), gpus_per_stage))
        start += n
    return PipelineTemplate(out, 0.0, num_layers, nodes, gpn)


def worker_sharded(rank, world, port, mode, M, mb, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    try:
        gpn = 2
        wide = wide_template(4, 1, 2, 1, gpn)            # one stage on both GPUs of a node: every layer sharded 2-way
        narrow = wide_template(4, 2, 1, 1, gpn)          # two 1-GPU stages on a node: rows [a, a], [b, b]
        plan = {"lone": [wide], "replicas": [wide, wide], "mixed": [wide, narrow]}[mode]
        eng = make_engine(rank % gpn, world // gpn, gpn, M, mb, steps, templates=[wide, narrow])
        assert eng._rank == rank
        eng.instantiate_pipelines(M, plan=plan)
        layers = eng._pipeline.execution._layers
        sharded_here = mode != "mixed" or rank < 2
        for l in layers:
            assert l.sharded == sharded_here
            assert l._param_handle._sharding_strategy == ("FULL_SHARD" if sharded_here else "NO_SHARD")
            assert l._state.columns == gpn
        if sharded_here:   # ONE communicator for the stage, shared by its layers
            assert len({id(l._state.comm) for l in layers}) == 1
        for _ in range(steps):
            eng._train_step()
        out = {}
        for l in layers:
            st = l._state
            # one reduce-scatter per step, one all-gather per step after the first (initial values are deterministic)
            assert (st.scatters, st.gathers) == ((steps, steps - 1) if sharded_here else (0, 0))
            out[l.layer_id] = (st.lo, st.param_shard.numpy().copy(), st.exp_avg.numpy().copy(), st.sharded)
        q.put((rank, out, float(eng._pipeline.execution.total_loss) if eng._pipeline.is_last_stage() else None, None))
        dist.barrier()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


def check_against(results, flats, ms):
    seen = {}
    for rank, out, _, _ in results:
        for lid, (lo, shard, m, sharded) in out.items():
            n = flats[lid].numel()
            want_p = torch.zeros(max(n, lo + shard.size)); want_p[:n] = flats[lid]
            want_m = torch.zeros_like(want_p); want_m[:n] = ms[lid]
            torch.testing.assert_close(torch.from_numpy(shard), want_p[lo:lo + shard.size], rtol=1e-4, atol=2e-6)
            # moments of near-zero gradient entries: the summation order (threads, reduce-scatter) leaves ~1e-8 of noise
            torch.testing.assert_close(torch.from_numpy(m), want_m[lo:lo + m.size], rtol=1e-3, atol=1e-7)
            seen.setdefault(lid, []).append((rank, lo, shard.size))
    return seen


@pytest.mark.timeout(300)
def test_one_stage_on_two_gpus_shards_every_layer():
    """1 node x 2 GPUs, one stage: FULL_SHARD (layer.py:100-102).  Both columns run the pipeline's 4 micro-batches;
    the reduce-scatter sums them: the update sees 2 x the gradient."""
    M, mb, steps = 4, 1, 3
    results = run_spawn(worker_sharded, 2, "lone", M, mb, steps)
    flats, ms, totals = weighted_reference([M], [2], mb, steps)
    seen = check_against(results, flats, ms)
    for lid, parts in seen.items():         # two disjoint halves that tile the (padded) vector
        assert sorted(p[1] for p in parts) == [0, parts[0][2]] and 2 * parts[0][2] >= flats[lid].numel()
    for _, _, total, _ in results:          # both columns are last stages and saw the same losses
        assert abs(total - totals[0]) < 1e-5 * abs(totals[0])


@pytest.mark.timeout(300)
def test_two_sharded_replicas_reduce_shard_by_shard():
    """2 nodes x 2 GPUs, two replicas of the wide pipeline: cross-replica groups exist per fsdp_index ([0, 2] and
    [1, 3], engine.py:374-392) and reduce the reduce-scattered SHARDS."""
    M, mb, steps = 4, 1, 2
    results = run_spawn(worker_sharded, 4, "replicas", M, mb, steps)
    flats, ms, totals = weighted_reference([2, 2], [2, 2], mb, steps)
    check_against(results, flats, ms)
    for rank, _, total, _ in results:
        assert abs(total - totals[rank // 2]) < 1e-5 * abs(totals[rank // 2])


@pytest.mark.timeout(300)
def test_sharded_pipeline_next_to_an_unsharded_one():
    """Heterogeneous replicas: node 0 runs one 2-GPU stage (sharded), node 1 two 1-GPU stages.  Rank 2 / 3 hold whole
    layers and reduce them column by column -- column 0 with rank 0, column 1 with rank 1 (layer.py:279-291)."""
    M, mb, steps = 4, 1, 2
    results = run_spawn(worker_sharded, 4, "mixed", M, mb, steps)
    flats, ms, totals = weighted_reference([2, 2], [2, 1], mb, steps)
    seen = check_against(results, flats, ms)
    assert sorted(seen) == [0, 1, 2, 3]
    for lid, parts in seen.items():          # two half shards on node 0 + the whole layer on rank 2 or 3
        assert len(parts) == 3 and sorted(p[0] for p in parts)[:2] == [0, 1]


# ---- elastic: sharded stages lose a node -----------------------------------------------------------------------------
NODE_IPS = ["127.0.0.1", "127.0.0.2", "127.0.0.3", "127.0.0.4"]
GPN, STEPS_BEFORE, STEPS_TOTAL = 2, 2, 4


def worker_nodes(rank, pipe, q, ready):
    torch.set_num_threads(1)
    try:
        from unittest.mock import patch

        from oracle_layer import OracleLayer

        from oobleck_b200.execution.dataloader import SyntheticTokenDataset
        from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine
        node, local_rank = divmod(rank, GPN)
        patch("socket.gethostbyname", return_value=NODE_IPS[node]).start()      # test_engine.py:676
        real_tcpstore = torch.distributed.TCPStore
        patch("torch.distributed.TCPStore", lambda host_name, *a, **kw: real_tcpstore("127.0.0.1", *a, **kw)).start()
        M, mb = 4, 1
        args = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=mb * M, steps=STEPS_TOTAL),
                                model=ModelArguments(model_name="gpt2", model_tag="t", model_args=dict(MARGS)))
        ds = SyntheticTokenDataset(num_samples=256, seq_len=32, vocab_size=211, pin_memory=False)
        one_node = wide_template(4, 1, GPN, 1, GPN)          # one 2-GPU stage
        two_nodes = wide_template(4, 2, GPN, 2, GPN)         # two 2-GPU stages
        # worker_main's call sequence (elastic/worker.py:23-34), two workers per node
        eng = OobleckEngine(local_rank, len(NODE_IPS), GPN, pipe, args, dataset=ds, layer_cls=OracleLayer,
                            templates=[one_node, two_nodes], backend="gloo", comm_timeout_s=90)
        eng.initialize_distributed()
        assert eng._rank == rank and eng._rank_map[NODE_IPS[node]] == [node * GPN, node * GPN + 1]
        eng.instantiate_pipelines(M, plan=[two_nodes, two_nodes])
        assert all(l.sharded for l in eng._pipeline.execution._layers)
        orig_step = eng._guarded_train_step
        count = {"n": 0}

        def step_hook():
            if count["n"] == STEPS_BEFORE:
                if node == len(NODE_IPS) - 1:
                    q.put((rank, "gone", None))
                    q.close(); q.join_thread()          # noqa: E702
                    os._exit(0)                         # the whole node dies: both of its workers
                ready.put(rank)
            count["n"] += 1
            return orig_step()
        eng._guarded_train_step = step_hook
        eng.train()
        assert [p._ranks for p in eng._reconfiguration._pipelines] == [[4, 5], [0, 1, 2, 3]]
        assert eng._dist_info.world_size == 6 and NODE_IPS[3] not in eng._rank_map
        layers = eng._pipeline.execution._layers
        assert sorted(l.layer_id for l in layers) == ([0, 1, 2, 3] if rank in (4, 5) else ([0, 1] if rank < 2 else [2, 3]))
        out = {l.layer_id: (l._state.lo, l._state.param_shard.numpy().copy(), l.exp_avg.numpy().copy(), l.sharded,
                            l.opt_step) for l in layers}
        q.put((rank, out, len(eng.step_seconds)))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None))
        raise


@pytest.mark.timeout(600)
def test_sharded_stages_survive_a_lost_node():
    """4 nodes x 2 GPUs, two replicas of a 2-stage pipeline whose stages own a node each (every layer sharded 2-way).
    Node 3 dies inside a step: the reference's policy shrinks the second replica to ONE 2-GPU stage on node 2
    ([4, 5]); the layers node 3 held arrive shard by shard from the first replica (rank 2 -> 4, rank 3 -> 5:
    engine.py:278-303 pairs by fsdp_index), parameters and moments, and are gathered inside the new stage on the next
    forward.  The run matches the never-failed oracle (2 x the gradient: both shard columns run the micro-batches)."""
    import threading

    import torch.multiprocessing as mp

    from oobleck_b200.execution.engine import DistributionInfo
    world = len(NODE_IPS) * GPN
    ctx = mp.get_context("spawn")
    q, ready = ctx.Queue(), ctx.Queue()
    pipes = [ctx.Pipe(duplex=True) for _ in range(world)]
    procs = [ctx.Process(target=worker_nodes, args=(r, pipes[r][1], q, ready)) for r in range(world)]
    for p in procs:
        p.start()

    def broadcast_rank0_port(ps):                      # tests/execution/test_engine.py:650-657
        port = ps[0][0].recv()
        for pipe, _ in ps:
            pipe.send(port)

    def agent():
        for pipe, _ in pipes:
            pipe.send(DistributionInfo(list(NODE_IPS), world))
        broadcast_rank0_port(pipes)
        for _ in range(world - GPN):
            ready.get(timeout=300)
        for p in procs[world - GPN:]:
            p.join(timeout=60)
        for pipe, _ in pipes[:world - GPN]:
            pipe.send(NODE_IPS[-1])
        broadcast_rank0_port(pipes[:world - GPN])

    t = threading.Thread(target=agent, daemon=True)
    t.start()
    results = {}
    for _ in range(world):
        r = q.get(timeout=500)
        results[r[0]] = r
    t.join(timeout=60)
    for p in procs:
        p.join(timeout=60)
    assert results[6][1] == "gone" and results[7][1] == "gone"
    for r in range(6):
        assert not isinstance(results[r][1], str), results[r][1]
        assert results[r][2] == STEPS_TOTAL

    # never-failed run: one process, all micro-batches, 2 x the gradient; the sampler restarts after the reconfiguration
    from test_engine_agent_pipe_gloo import never_failed_reference
    flats, ms = never_failed_reference(MARGS, 4, grad_scale=2.0)
    for r in range(6):
        for lid, (lo, shard, m, sharded, opt_step) in results[r][1].items():
            assert sharded and opt_step == STEPS_TOTAL
            n = flats[lid].numel()
            want_p = torch.zeros(max(n, lo + shard.size)); want_p[:n] = flats[lid]
            want_m = torch.zeros_like(want_p); want_m[:n] = ms[lid]
            torch.testing.assert_close(torch.from_numpy(shard), want_p[lo:lo + shard.size], rtol=1e-4, atol=2e-6)
            torch.testing.assert_close(torch.from_numpy(m), want_m[lo:lo + m.size], rtol=1e-3, atol=1e-7)


# ---- the reference's own FSDP test, restated (tests/execution/test_engine.py:805-881) --------------------------------
def worker_fsdp_engine(rank, pipe, q, num_stages):
    torch.set_num_threads(2)
    try:
        from unittest.mock import patch

        from oracle_layer import OracleLayer

        from oobleck_b200.execution.dataloader import SyntheticTokenDataset
        from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine
        num_gpus_per_node, M, mb, steps = 4, 2, 1, 2                     # "Assume all GPUs are in one node" (:814-816)
        patch("socket.gethostbyname", return_value="127.0.0.1").start()
        args = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=mb * M, steps=steps),
                                model=ModelArguments(model_name="gpt2", model_tag="t", model_args=dict(MARGS)))
        ds = SyntheticTokenDataset(num_samples=128, seq_len=32, vocab_size=211, pin_memory=False)
        template = wide_template(4, num_stages, num_gpus_per_node // num_stages, 1, num_gpus_per_node)
        eng = OobleckEngine(rank, 1, num_gpus_per_node, pipe, args, dataset=ds, layer_cls=OracleLayer,
                            templates=[template], backend="gloo", listen=False)
        eng.initialize_distributed()
        eng.instantiate_pipelines(M)
        assert len(eng._reconfiguration._pipelines) == 1                 # :841-842
        assert len(eng._pipeline._ranks) == 4                            # :844-845
        for layer in eng._pipeline.execution._layers:                    # :847-860
            if num_stages == 4:
                assert layer._param_handle._sharding_strategy == "NO_SHARD" and layer._group_size == 1
            else:
                assert layer._param_handle._sharding_strategy == "FULL_SHARD"
                assert layer._group_size == 4 // num_stages
        for _ in range(steps):
            eng._train_step()                                            # :862
        out = {l.layer_id: (l._state.lo, l._state.param_shard.numpy().copy(), l.exp_avg.numpy().copy(), l.sharded)
               for l in eng._pipeline.execution._layers}
        q.put((rank, out, float(eng._pipeline.execution.total_loss) if eng._pipeline.is_last_stage() else None, None))
        dist.barrier()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


@pytest.mark.timeout(300)
@pytest.mark.parametrize("num_stages", [1, 2, 4])
def test_fsdp_engine_train(num_stages):
    """The reference's ``test_fsdp_engine_train``: 4 GPUs of ONE node, ``num_stages`` stages of 4 / num_stages GPUs, the
    engine driven through the agent pipe.  Same assertions (one pipeline, 4 ranks, sharding strategy and group size per
    layer, a train step runs) -- plus the result: two steps against the oracle, the gradient entering 4 / num_stages
    times (every shard column runs the micro-batches; with 2 stages the grid rows are [0, 0, 1, 1] / [2, 2, 3, 3]: four
    columns over two holders, each holding two units)."""
    import threading

    import torch.multiprocessing as mp

    from oobleck_b200.execution.engine import DistributionInfo
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pipes = [ctx.Pipe(duplex=True) for _ in range(4)]
    for pipe, _ in pipes:
        pipe.send(DistributionInfo(["127.0.0.1"], 4))                    # :870-872

    def broadcast_rank0_port():                                          # :650-657
        port = pipes[0][0].recv()
        for pipe, _ in pipes:
            pipe.send(port)
    t = threading.Thread(target=broadcast_rank0_port, daemon=True)
    t.start()
    procs = [ctx.Process(target=worker_fsdp_engine, args=(r, pipes[r][1], q, num_stages)) for r in range(4)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=240) for _ in range(4)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    t.join(timeout=30)
    for r in results:
        assert r[3] is None, r[3]
    k = 4 // num_stages
    flats, ms, totals = weighted_reference([2], [k], 1, 2)
    seen = check_against(results, flats, ms)
    assert sorted(seen) == [0, 1, 2, 3]
    for lid, parts in seen.items():
        assert len(parts) == k and sorted(p[1] for p in parts) == [i * parts[0][2] for i in range(k)] if k > 1 else True
    for _, _, total, _ in results:
        if total is not None:
            assert abs(total - totals[0]) < 1e-5 * abs(totals[0])
