"""Cost of sharding a stage over K GPUs: GPT-2-XL width at reduced depth, one stage, K shard columns.

    torchrun --nproc-per-node K --master-addr 127.0.0.1 tools/sharded_stage_timing.py [--depth 12] [--steps 4]

Every column runs the pipeline's micro-batches (reference semantics), so the compute per GPU equals the 1-GPU run: what
is measured is what the sharding adds to a step -- ONE in-place all-gather and ONE reduce-scatter per layer -- and what
it saves in HBM (parameters' moments and the reduce-scattered gradient exist for 1/K of each layer).  K = 1 gives the
unsharded reference time.  Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--microbatches", type=int, default=8)
    a = ap.parse_args()
    from bench import MODELS, VOCAB
    from oobleck_b200.execution.dataloader import SyntheticTokenDataset
    from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine
    from oobleck_b200.execution.p2p import NvlinkRingTransport
    from oobleck_b200.planning.pipeline_template import PipelineTemplate, StageExecutionResult
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.cuda.set_device(rank)
    cfg = MODELS["gpt2-xl"]
    ma = dict(cfg["model_args"], num_hidden_layers=a.depth)
    mb, M, T = cfg["microbatch"], a.microbatches, ma["n_positions"]
    L = a.depth + 2
    oargs = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=mb * M, steps=a.steps),
                             model=ModelArguments(model_name="gpt2", model_tag="shard", model_args=ma))
    ds = SyntheticTokenDataset(num_samples=max(64, mb * M * (a.steps + 2)), seq_len=T, vocab_size=VOCAB)
    tmpl = PipelineTemplate([StageExecutionResult(range(L), world)], 0.0, L, 1, world)
    eng = OobleckEngine(rank, 1, world, None, oargs, dataset=ds, templates=[tmpl], transport_cls=NvlinkRingTransport)
    eng.initialize_distributed("nccl")
    torch.cuda.reset_peak_memory_stats()
    eng.instantiate_pipelines(M, plan=[tmpl])
    layers = eng._pipeline.execution._layers
    times = []
    for s in range(a.steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng._train_step()
        e1.record()
        torch.cuda.synchronize()
        if s:                       # step 0 warms up (and has no gather: initial values are deterministic)
            times.append(e0.elapsed_time(e1))
    t = torch.tensor([sum(times) / len(times)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathers_in_training = layers[0]._state.gathers
    # the collectives alone, back to back on an idle GPU: upper bound of what a step can expose
    comm_ms = None
    if world > 1:
        for l in layers:
            l._state.stale = True
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for l in layers:
            l.unshard_params()
        e1.record()
        for l in layers:
            l._state.grads_scattered = False
            l.prepare_gradient_for_optim()
        e2.record()
        torch.cuda.synchronize()
        comm_ms = {"all_gather_and_plane_split_ms": e0.elapsed_time(e1), "reduce_scatter_ms": e1.elapsed_time(e2)}
    numel = sum(l.numel for l in layers)
    held = sum(l.flat_param.numel() for l in layers)
    if rank == 0:
        print(json.dumps({
            "workload": f"gpt2-xl width, {a.depth} blocks + embedding + head in ONE stage over {world} GPU(s), "
                        f"micro-batch {mb} x {M} per step, T={T}",
            "shard_columns": world, "ms_per_step": float(t), "tokens_per_s_per_column": mb * M * T / float(t) * 1e3,
            "collectives_alone": comm_ms,
            "gathers_per_layer_and_step": gathers_in_training / max(1, a.steps),
            "parameters": numel, "parameters_held_per_gpu": held,
            "optimizer_state_bytes_per_gpu": 8 * held, "gathered_copy_bytes_per_gpu": 4 * numel,
            "peak_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "loss": float(eng._pipeline.execution.total_loss)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
