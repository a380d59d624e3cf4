"""Reconfiguration latency on real GPUs (the second half of BASELINE.json's metric).

    python tools/reconfig_bench.py --gpus 4 --model gpt2          # 2 replicas x 2 stages, lose the last GPU

Scenario supported by the reference's semantics (SURVEY 7 hard part 5: a lone pipeline cannot survive a loss without a
replica): `gpus/2` -stage replicas x 2; after 2 training steps the last rank leaves; the survivors run the reference's
re-planning policy, rebuild pipelines (no world teardown), copy the moved layers from the surviving replica over NCCL,
and train on.  Reported: seconds from the loss notification to (a) pipelines rebuilt + states copied, (b) first completed
post-reconfiguration train step.
"""
import argparse
import json
import os
import socket
import sys
import time

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import MODELS  # noqa: E402


def log(rank, msg, t0=[None]):
    if t0[0] is None:
        t0[0] = time.perf_counter()
    print(f"[rank {rank} +{time.perf_counter() - t0[0]:7.2f}s] {msg}", file=sys.stderr, flush=True)


def worker(rank, world, port, model, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist

    from oobleck_b200.execution.dataloader import SyntheticTokenDataset
    from oobleck_b200.execution.engine import (JobArguments, ModelArguments, OobleckArguments, OobleckEngine,
                                               layer_cost_model)
    from oobleck_b200.execution.p2p import NvlinkRingTransport
    from oobleck_b200.planning.pipeline_template import balanced_template
    torch.cuda.set_device(rank)
    cfg = MODELS[model]
    ma = cfg["model_args"]
    mb, gb = cfg["microbatch"], 16 * cfg["microbatch"]
    args = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=gb, steps=1),
                            model=ModelArguments(model_name="gpt2", model_tag=model, model_args=dict(ma)))
    ds = SyntheticTokenDataset(num_samples=4096, seq_len=ma["n_positions"], vocab_size=ma.get("vocab_size", 50257))
    eng = OobleckEngine(rank, world, 1, None, args, dataset=ds, transport_cls=NvlinkRingTransport)
    costs = layer_cost_model(eng._model, mb)
    half = world // 2
    eng._pipeline_templates = [balanced_template(costs, n) for n in range(1, half + 1)]
    log(rank, "engine built")
    eng.initialize_distributed("nccl")
    eng.instantiate_pipelines(gb // mb, plan=[eng._pipeline_templates[-1]] * 2)
    log(rank, "pipelines instantiated")
    for i in range(2):
        eng._train_step()
        torch.cuda.synchronize()
        log(rank, f"train step {i} done")
    dist.barrier()
    log(rank, "barrier passed; rank %d leaves now" % (world - 1))
    lost = world - 1
    if rank == lost:
        q.put((rank, None))
        time.sleep(20)   # keep the process (and its CUDA context / IPC exports) out of the way, like a dead node
        return
    t0 = time.perf_counter()
    eng._reconfiguration.on_reconfigure([lost])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    log(rank, f"reconfigured in {t1 - t0:.3f}s: {[p._ranks for p in eng._reconfiguration._pipelines]}")
    eng._train_step()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    log(rank, f"first post-reconfiguration step done at {t2 - t0:.3f}s")
    q.put((rank, {"rebuild_s": t1 - t0, "first_step_done_s": t2 - t0,
                  "new_ranks": [p._ranks for p in eng._reconfiguration._pipelines],
                  "my_layers": len(eng._pipeline.execution._layers)}))
    time.sleep(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=4)
    ap.add_argument("--model", default="gpt2")
    a = ap.parse_args()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, a.gpus, port, a.model, q)) for r in range(a.gpus)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(a.gpus))
    for p in procs:
        p.join(timeout=60)
    alive = {k: v for k, v in res.items() if v}
    print(json.dumps({"metric": "reconfiguration_latency_s", "model": a.model, "gpus": a.gpus,
                      "rebuild_s_max": max(v["rebuild_s"] for v in alive.values()),
                      "first_step_done_s_max": max(v["first_step_done_s"] for v in alive.values()),
                      "new_ranks": next(iter(alive.values()))["new_ranks"], "per_rank": alive}))


if __name__ == "__main__":
    main()
