"""Reconfiguration latency after a REAL kill (the second half of BASELINE.json's metric; SURVEY 8d: "wall time from the
lost-node message arriving on the worker pipe to the first completed post-reconfiguration train step").

    python bench.py --reconfig --gpus 8 --replicas 2 [--model gpt2-xl]      # or: python tools/reconfig_bench.py ...

This process plays the agent, with the reference's own fake-agent protocol (tests/execution/test_engine.py:650-657,
1037-1053): it spawns one worker per GPU -- each runs exactly ``worker_main``'s call sequence (elastic/worker.py:23-34)
on ``OobleckEngine(local_rank, num_nodes, 1, pipe, args)`` -- sends ``DistributionInfo`` down every pipe and re-broadcasts
rank 0's TCPStore port.  When every worker has started training step ``--kill-step`` the worker of the last rank dies by
SIGKILL (its pipeline neighbours are left spinning on NVLink flags, everybody else runs into a vote that cannot complete);
the agent notices the death, announces the lost IP to the survivors and re-broadcasts the port once more.

The survivors' listener threads release the GPU (host-mapped abort words for the P2P kernels, ncclCommAbort for the
communicators that contained the victim), the training threads drop the step in flight, re-plan with the reference's
policy (2 x 4 stages -> 4 + 3), rebuild pipelines and links without touching the world group, receive the layers they
now own (parameters + Adam moments) from the surviving replica, and train on.  With ``--replicas 1`` (BASELINE config 5:
one 8-stage pipeline -> the 7-stage template) there is no replica: the stage state comes from the peer shadows
(``PeerShadow``: every stage mirrors its successor's parameters and moments over NVLink after each step).
Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import signal
import sys
import threading
import time

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _ips(world):
    return [f"127.0.0.{r + 1}" for r in range(world)]


def _die_with_parent():
    """Workers must not outlive the agent process (a killed benchmark would leave them spinning on the GPUs)."""
    parent = os.getppid()

    def watch():
        while True:
            time.sleep(1.0)
            if os.getppid() != parent:
                os._exit(1)
    threading.Thread(target=watch, daemon=True).start()


def worker(rank, world, pipe, q, started, model, replicas, steps, kill_step):
    from unittest.mock import patch
    _die_with_parent()
    os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"   # a dead peer is handled by the engine, not by the NCCL watchdog
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    os.environ.setdefault("OOB_P2P_TIMEOUT_S", "60")
    try:
        from bench import MODELS, VOCAB
        from oobleck_b200.execution.dataloader import SyntheticTokenDataset
        from oobleck_b200.execution.engine import (JobArguments, ModelArguments, OobleckArguments, OobleckEngine,
                                                   layer_cost_model)
        from oobleck_b200.execution.p2p import NvlinkRingTransport
        from oobleck_b200.planning.pipeline_template import balanced_template
        torch.cuda.set_device(rank)
        ips = _ips(world)
        patch("socket.gethostbyname", return_value=ips[rank]).start()      # one "node" per GPU (test_engine.py:676)
        real_store = torch.distributed.TCPStore
        patch("torch.distributed.TCPStore", lambda host_name, *a, **kw: real_store("127.0.0.1", *a, **kw)).start()

        cfg = MODELS[model]
        ma = dict(cfg["model_args"])
        mb = cfg["microbatch"]
        stages = world // replicas
        gb = mb * 8 * world                 # 8 micro-batches per GPU and step (tokens_per_step below): a 1F1B steady state
        oargs = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=gb, steps=steps),
                                 model=ModelArguments(model_name="gpt2", model_tag=model, model_args=ma))
        ds = SyntheticTokenDataset(num_samples=max(2334, gb * (steps + 4)), seq_len=ma["n_positions"],
                                   vocab_size=ma.get("vocab_size", VOCAB))
        # worker_main: ctor -> initialize_distributed -> instantiate_pipelines -> train
        eng = OobleckEngine(0, world, 1, pipe, oargs, dataset=ds, transport_cls=NvlinkRingTransport,
                            peer_shadow=(replicas == 1))   # a lone pipeline survives through its neighbours' mirrors
        costs = layer_cost_model(eng._model, mb)
        eng._pipeline_templates = [balanced_template(costs, n, 1) for n in range(1, world + 1)]
        eng._templates_injected = True
        eng.initialize_distributed()
        plan_t = next(t for t in eng._pipeline_templates if t._num_nodes == stages)
        eng.instantiate_pipelines(gb // mb, plan=[plan_t] * replicas)

        orig = eng._guarded_train_step
        count = {"n": 0}

        def hook():
            if count["n"] == kill_step:
                started.put(rank)
                if rank == world - 1:
                    # the victim dies here, with SIGKILL: its peers are inside this step -- pipeline neighbours spinning
                    # on NVLink flags it will never write, everybody else on their way into the vote
                    time.sleep(0.05)
                    os.kill(os.getpid(), signal.SIGKILL)
            count["n"] += 1
            return orig()
        eng._guarded_train_step = hook
        t_train0 = time.perf_counter()
        eng.train()
        torch.cuda.synchronize()
        rc = eng._reconfiguration
        t_note = rc.last_notification_time
        first_after = next((t for t in eng.step_end_times if t_note is not None and t > t_note), None)
        # replicas of a layer must hold identical parameters after the move (bit for bit: same NCCL reduction on both)
        sums = {l.layer_id: float(l.flat_param.double().sum()) for l in eng._pipeline.execution._layers}
        q.put((rank, {
            "pipelines": [p._ranks for p in rc._pipelines],
            "notify_to_rebuilt_s": rc.last_reconfiguration_seconds, "breakdown": rc.last_breakdown,
            "notify_to_first_step_s": (first_after - t_note) if first_after is not None else None,
            "steps_completed": len(eng.step_seconds),
            "step_s_before": eng.step_seconds[:kill_step], "step_s_after": eng.step_seconds[kill_step:],
            "loss": float(eng._pipeline.execution.total_loss) if eng._pipeline.is_last_stage() else None,
            "param_sums": sums, "train_wall_s": time.perf_counter() - t_train0}))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
        raise


def main(args=None):
    if args is None or not hasattr(args, "kill_step"):
        ap = argparse.ArgumentParser()
        ap.add_argument("--gpus", type=int, default=8)
        ap.add_argument("--replicas", type=int, default=2)
        ap.add_argument("--model", default="gpt2-xl")
        ap.add_argument("--steps", type=int, default=5)
        ap.add_argument("--kill-step", type=int, default=2)
        ap.add_argument("--reconfig", action="store_true")
        known, _ = ap.parse_known_args()
        if args is not None:                      # called from bench.py: keep its --gpus / --replicas / --model / --steps
            for k in ("gpus", "replicas", "model"):
                setattr(known, k, getattr(args, k))
            known.steps = max(getattr(args, "steps", 5), known.kill_step + 2)
        args = known
    from oobleck_b200.execution.engine import DistributionInfo
    world = args.gpus
    assert world % args.replicas == 0 and world // args.replicas >= 1
    assert args.replicas >= 2 or world >= 3, "a lone pipeline needs >= 3 stages to lose one and keep a mirror"
    victim = world - 1
    ips = _ips(world)
    ctx = mp.get_context("spawn")
    q, started = ctx.Queue(), ctx.Queue()
    pipes = [ctx.Pipe(duplex=True) for _ in range(world)]
    procs = [ctx.Process(target=worker, args=(r, world, pipes[r][1], q, started, args.model, args.replicas, args.steps,
                                              args.kill_step)) for r in range(world)]
    for p in procs:
        p.start()
    marks = {}

    def rebroadcast(ps):
        port = ps[0][0].recv()
        for pipe, _ in ps:
            pipe.send(port)

    def agent():
        for pipe, _ in pipes:
            pipe.send(DistributionInfo(list(ips), world))
        rebroadcast(pipes)
        for _ in range(world):
            started.get(timeout=1800)
        marks["kill"] = time.perf_counter()
        procs[victim].join(timeout=60)                    # the victim SIGKILLs itself inside this step (see worker)
        if procs[victim].is_alive():
            os.kill(procs[victim].pid, signal.SIGKILL)
            procs[victim].join(timeout=30)
        marks["announce"] = time.perf_counter()
        survivors = [p for i, p in enumerate(pipes) if i != victim]
        for pipe, _ in survivors:
            pipe.send(ips[victim])
        rebroadcast(survivors)

    t = threading.Thread(target=agent, daemon=True)
    t.start()
    results = {}
    for _ in range(world - 1):
        r = q.get(timeout=3000)
        results[r[0]] = r[1]
    t.join(timeout=60)
    for i, p in enumerate(procs):
        p.join(timeout=60)
    errors = {r: v for r, v in results.items() if isinstance(v, str)}
    if errors:
        print(json.dumps({"metric": "reconfiguration_latency_s", "error": errors}), flush=True)
        sys.exit(1)
    worst_first = max(v["notify_to_first_step_s"] for v in results.values())
    worst_rebuilt = max(v["notify_to_rebuilt_s"] for v in results.values())
    # consistency: every replica of a layer ends with the same parameter checksum
    by_layer = {}
    for v in results.values():
        for lid, s in v["param_sums"].items():
            by_layer.setdefault(lid, set()).add(s)
    out = {
        "metric": "reconfiguration_latency_s", "value": worst_first, "unit": "s", "higher_is_better": False,
        "n_gpus": world, "definition": "lost-node message received on the worker pipe -> first completed "
                                       "post-reconfiguration train step, max over the survivors (SURVEY 8d)",
        "config": {"workload": f"{args.model}: {args.replicas} replicas x {world // args.replicas} stages, rank {victim} "
                               f"SIGKILLed inside training step {args.kill_step}", "model": args.model},
        "tokens_per_step": tokens_per_step(args.model, world),
        "notify_to_pipelines_rebuilt_and_states_copied_s": worst_rebuilt,
        "agent_kill_to_announce_s": marks["announce"] - marks["kill"],
        "pipelines_after": next(iter(results.values()))["pipelines"],
        "step_s_before": statistics_of([x for v in results.values() for x in v["step_s_before"][1:]]),
        "step_s_after": statistics_of([x for v in results.values() for x in v["step_s_after"][1:]]),
        "replicas_identical_after": all(len(s) == 1 for s in by_layer.values()),
        "per_rank": {str(r): {k: v[k] for k in ("notify_to_rebuilt_s", "breakdown", "notify_to_first_step_s",
                                                "steps_completed")}
                     for r, v in sorted(results.items())},
    }
    print(json.dumps(out), flush=True)


def tokens_per_step(model: str, world: int) -> int:
    """Tokens one optimizer step of the job consumes (all replicas): the worker's global batch x T."""
    from bench import MODELS
    cfg = MODELS[model]
    return cfg["microbatch"] * 8 * world * cfg["model_args"]["n_positions"]


def statistics_of(xs):
    if not xs:
        return None
    xs = sorted(xs)
    return {"median": xs[len(xs) // 2], "max": xs[-1], "n": len(xs)}


if __name__ == "__main__":
    main()
