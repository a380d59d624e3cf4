"""2-GPU tests (skipped on a 1-GPU box): 2-stage 1F1B pipeline over the NVLink mailbox transport with the CUDA stage
layers, and 2 replicas with the NCCL gradient all-reduce, both against the oracle."""
import itertools
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs 2 GPUs", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARGS = dict(n_embd=128, n_head=2, num_hidden_layers=4, n_positions=128, vocab_size=1000)


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, mode, M, mb, steps, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    try:
        import torch.distributed as dist

        from oobleck_b200.execution.dataloader import SyntheticTokenDataset
        from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine
        from oobleck_b200.execution.p2p import NvlinkRingTransport
        from oobleck_b200.planning.pipeline_template import even_template
        torch.cuda.set_device(rank)
        args = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=mb * M, steps=steps),
                                model=ModelArguments(model_name="gpt2", model_tag="t", model_args=dict(MARGS)))
        ds = SyntheticTokenDataset(num_samples=256, seq_len=128, vocab_size=1000)
        templates = [even_template(6, 2)] if mode == "pp" else [even_template(6, 1)]
        if mode == "fsdp":    # ONE stage that owns both GPUs of a node: every layer sharded 2-way (layer.py:100-102)
            from oobleck_b200.planning.pipeline_template import PipelineTemplate, StageExecutionResult
            templates = [PipelineTemplate([StageExecutionResult(range(6), 2)], 0.0, 6, 1, 2)]
            eng = OobleckEngine(rank, 1, 2, None, args, dataset=ds, templates=templates,
                                transport_cls=NvlinkRingTransport)
        else:
            eng = OobleckEngine(rank, world, 1, None, args, dataset=ds, templates=templates,
                                transport_cls=NvlinkRingTransport)
        eng.initialize_distributed("nccl")
        eng.instantiate_pipelines(M)
        totals = []
        for _ in range(steps):
            eng._train_step()
            tl = eng._pipeline.execution.total_loss
            totals.append(float(tl) if tl is not None else None)
        torch.cuda.synchronize()
        out = {l.layer_id: l.flat_param.cpu().numpy().copy() for l in eng._pipeline.execution._layers}
        if mode == "fsdp":
            for l in eng._pipeline.execution._layers:
                st = l._state
                assert l.sharded and (st.scatters, st.gathers) == (steps, steps - 1), (st.scatters, st.gathers)
                # the gathered copy and the planes follow the last update once the next forward asks for them
                l.unshard_params()
                out[l.layer_id] = (st.lo, l.flat_param.cpu().numpy().copy(), l.full_param.cpu().numpy().copy(),
                                   l.exp_avg.cpu().numpy().copy())
        q.put((rank, out, totals, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


def reference_run(M, mb, steps, num_pipelines, grad_scale=1.0, return_moments=False):
    sys.path.insert(0, ROOT)
    from oobleck_b200.execution.dataloader import OobleckSampler, SyntheticTokenDataset
    from oobleck_b200.module.model import OobleckModel
    from oracle import gpt2 as og
    from oracle import optim as oo
    model = OobleckModel("gpt2", {"input_ids": None, "attention_mask": None, "labels": None}, None, "t", dict(MARGS))
    d = og.GPT2Dims(n_embd=128, n_head=2, n_layer=4, n_positions=128, vocab_size=1000)
    layers = og.build_layers(d)
    flats = [spec.init_flat() for spec in model.layers]
    for l, f in zip(layers, flats):
        og.load_flat_(l, f)
    ds = SyntheticTokenDataset(num_samples=256, seq_len=128, vocab_size=1000, pin_memory=False)
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
        for i, l in enumerate(layers):
            oo.adamw_step_(flats[i], og.flat_grads(l) * grad_scale, ms[i], vs[i], step + 1, lrs[step])
            og.load_flat_(l, flats[i])
    if return_moments:
        return flats, losses, ms
    return flats, losses


def run(mode, M, mb, steps):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, mode, M, mb, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=400) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert r[3] is None, r[3]
    return sorted(results, key=lambda r: r[0])


def test_two_stage_pipeline_nvlink_p2p():
    M, mb, steps = 6, 2, 4
    results = run("pp", M, mb, steps)
    flats, losses = reference_run(M, mb, steps, 1)
    seen = {}
    for _, out, _, _ in results:
        seen.update(out)
    assert sorted(seen) == list(range(6))
    for lid, f in seen.items():
        err = (torch.from_numpy(f) - flats[lid]).abs().max() / flats[lid].abs().max()
        assert err < 1e-4, (lid, float(err))
    ref_cum = list(itertools.accumulate(sum(t) for t in losses))
    for got, want in zip(results[1][2], ref_cum):
        assert abs(got - want) < 1e-4 * abs(want)


def test_two_replicas_nccl_allreduce():
    M, mb, steps = 4, 2, 3
    results = run("dp", M, mb, steps)
    flats, losses = reference_run(M, mb, steps, 2)
    for rank, out, totals, _ in results:
        for lid, f in out.items():
            err = (torch.from_numpy(f) - flats[lid]).abs().max() / flats[lid].abs().max()
            assert err < 1e-4, (lid, float(err))
        want = sum(t[rank] for t in losses)
        assert abs(totals[-1] - want) < 1e-4 * abs(want)


def test_stage_sharded_over_two_gpus_nccl():
    """one stage on 2 GPUs.  Every layer keeps half of its parameters and moments per GPU, gathers them in
    place once per step and reduce-scatters its gradient once per step (overlapped: the hook of the last micro-batch
    backward starts it on the communication stream).  Both shard columns run the pipeline's micro-batches and the
    reduce-scatter SUMs (reference semantics, layer.py:202-206): the update sees twice the gradient."""
    M, mb, steps = 4, 2, 3
    results = run("fsdp", M, mb, steps)
    flats, losses, ms = reference_run(M, mb, steps, 1, grad_scale=2.0, return_moments=True)
    ref_cum = list(itertools.accumulate(sum(t) for t in losses))
    for rank, out, totals, _ in results:
        assert sorted(out) == list(range(6))
        for lid, (lo, shard, full, m) in out.items():
            n = flats[lid].numel()
            want = torch.zeros(max(n, lo + shard.size)); want[:n] = flats[lid]
            want_m = torch.zeros_like(want); want_m[:n] = ms[lid]
            scale = flats[lid].abs().max()
            assert (torch.from_numpy(shard) - want[lo:lo + shard.size]).abs().max() / scale < 1e-4, lid
            assert (torch.from_numpy(full) - flats[lid]).abs().max() / scale < 1e-4, lid
            assert (torch.from_numpy(m) - want_m[lo:lo + m.size]).abs().max() / ms[lid].abs().max() < 1e-3, lid
        for got, want in zip(totals, ref_cum):
            assert abs(got - want) < 1e-4 * abs(want)
    a, b = results[0][1], results[1][1]
    for lid in a:       # the two ranks hold the two halves and agree on the gathered vector bit for bit
        assert a[lid][0] == 0 and b[lid][0] == a[lid][1].size
        assert (a[lid][2] == b[lid][2]).all()
