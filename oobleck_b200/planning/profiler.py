"""Measured per-layer latencies for stage balancing (the hot-path half of oobleck/planning/profiler.py:41-123).

The reference's profiler times ONE forward of every fx layer on rank 0 (``time.time_ns`` around an eager call + device
synchronize), assumes ``backward = 3 x forward`` (:104) and broadcasts the table.  Its numbers feed the C++ template
search (csrc/planning), which is control plane and not part of this package; what this module provides is the
measurement itself for the layers as THIS engine executes them -- forward and the hand-written backward timed
separately with CUDA events, no recompute term -- so that the stage split is balanced on what the kernels really cost
(the lm_head stage layer is ~2.5 transformer blocks on GPT-2-XL, not the 1.3 its FLOP count suggests once the
attention and LayerNorm kernels are accounted for).

All layers of one kind have identical shapes, so one layer per kind is timed (embed, block, head).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def measure_layer_costs(model, microbatch: int, device: torch.device | None = None, warmup: int = 2,
                        iters: int = 5) -> dict[str, dict[str, float]]:
    """{kind: {"forward": ms, "backward": ms}} for kind in embed / block / head, measured on ``device``."""
    from ..execution.layer import HiddenGrad, Layer, StageWorkspace
    device = device or torch.device("cuda", torch.cuda.current_device())
    by_kind = {}
    for l in model.layers:
        by_kind.setdefault(l.kind, l)
    spec0 = model.layers[0]
    T, E = spec0.n_positions, spec0.n_embd
    ws = StageWorkspace(microbatch, T, E, spec0.n_head, device)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, spec0.vocab_size, (microbatch, T), generator=g).to(device)
    hidden = torch.randn(microbatch, T, E, generator=g).to(device)
    out: dict[str, dict[str, float]] = {}
    for kind in ("embed", "block", "head"):
        if kind not in by_kind:
            continue
        layer = Layer(by_kind[kind].index, by_kind[kind], None, None, None, microbatch_size=microbatch,
                      num_pipe_buffers=1, workspace=ws)
        inputs = (ids, torch.ones_like(ids), ids) if kind == "embed" else (hidden, ids)
        grad_scale = layer.loss_scale if layer.bwd_fp16 else 1.0
        dy = HiddenGrad((torch.randn(microbatch, T, E, generator=g) * 1e-4 * grad_scale).to(device))
        fwd, bwd = [], []
        for it in range(warmup + iters):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            layer(inputs, buffer_id=0)
            e[1].record()
            layer.backward(0, None if kind == "head" else dy)
            ws.join()
            e[2].record()
            torch.cuda.synchronize(device)
            if it >= warmup:
                fwd.append(e[0].elapsed_time(e[1]))
                bwd.append(e[1].elapsed_time(e[2]))
        out[kind] = {"forward": sorted(fwd)[len(fwd) // 2], "backward": sorted(bwd)[len(bwd) // 2]}
        layer.remove_tensors()
        del layer
    del ws
    torch.cuda.empty_cache()
    return out


def _measured_table(model, microbatch: int, device: torch.device | None = None):
    """[kind][fwd, bwd] milliseconds, measured on global rank 0 and broadcast when a process group exists
    (profiler.py:108-116), so every rank derives the same pipeline templates from it."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    kinds = ["embed", "block", "head"]
    table = torch.zeros(len(kinds), 2, dtype=torch.float64)
    if rank == 0:
        m = measure_layer_costs(model, microbatch, device)
        for i, k in enumerate(kinds):
            if k in m:
                table[i, 0], table[i, 1] = m[k]["forward"], m[k]["backward"]
    if dist.is_initialized() and dist.get_world_size() > 1:
        dev = device or torch.device("cuda", torch.cuda.current_device())
        t = table.to(dev)
        dist.broadcast(t, 0)
        table = t.cpu()
    return {k: (float(table[i, 0]), float(table[i, 1])) for i, k in enumerate(kinds)}


def measured_layer_costs(model, microbatch: int, device: torch.device | None = None) -> list[float]:
    """fwd + bwd milliseconds per ``model.layers`` entry."""
    t = _measured_table(model, microbatch, device)
    return [t[l.kind][0] + t[l.kind][1] for l in model.layers]


def measured_layer_results(model, microbatch: int, device: torch.device | None = None):
    """The planner's input (``LayerExecutionResults``, what ``get_profile_results`` returns in the reference): measured
    forward / backward per layer, the memory model of ``StageLayerSpec`` for (parameters, activations)."""
    from .pipeline_template import LayerExecutionResult, LayerExecutionResults
    t = _measured_table(model, microbatch, device)
    return LayerExecutionResults([
        LayerExecutionResult(i, t[l.kind][0], t[l.kind][1], {}, {}, (4 * l.num_params, l.activation_bytes(microbatch)))
        for i, l in enumerate(model.layers)])


# ---- the reference's profile cache (planning/profiler.py:246-320 writes it, csrc get_profile_results reads it) -----------
PROFILE_CACHE = "/tmp/oobleck/profiles"       # profiler.py:23


def get_profile_path(model_name: str, model_tag: str, cache: str | None = None):
    """profiler.py:246-247."""
    from pathlib import Path
    return Path(cache or PROFILE_CACHE) / f"{model_name}-{model_tag}"


def save_profile_results(results, model_name: str, model_tag: str, microbatch_size: int, cache: str | None = None):
    """Write a ``LayerExecutionResults`` where the reference's control plane looks for it, in its own file format, so that
    its unchanged ``get_profile_results(model_name, model_tag, microbatch_size)`` (C++, pipeline_template.cpp:26-79; called
    from ``OobleckEngine._initialize_engine``, engine.py:485-489) reads back exactly this profile:

        <cache>/<model_name>-<model_tag>/mb<microbatch_size>.json      [{"forward", "backward", "mem_required": [p, a]}, ...]
        <cache>/<model_name>-<model_tag>/allreduce_in_node.json        [{"<gpus>": ms, ...}, ...]       one dict per layer
        <cache>/<model_name>-<model_tag>/allreduce_across_nodes.json   [{"<nodes>": ms, ...}, ...]

    (what ``profile()`` dumps at profiler.py:296-320).  Returns the directory."""
    import json
    directory = get_profile_path(model_name, model_tag, cache)
    directory.mkdir(parents=True, exist_ok=True)
    layers = results.get()
    with (directory / f"mb{microbatch_size}.json").open("w") as f:
        json.dump([{"forward": l._forward, "backward": l._backward, "mem_required": list(l._mem_required)} for l in layers], f)
    with (directory / "allreduce_in_node.json").open("w") as f:
        json.dump([{str(k): v for k, v in l._allreduce_in_node.items()} for l in layers], f)
    with (directory / "allreduce_across_nodes.json").open("w") as f:
        json.dump([{str(k): v for k, v in l._allreduce_across_nodes.items()} for l in layers], f)
    return directory
