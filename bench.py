#!/usr/bin/env python
"""Headline benchmark: GPT-2-XL training tokens/s through the Oobleck pipeline-execution hot path.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference path's CPU restatement (oracle port)

Workload (BASELINE.json metric / configs[2], examples/gpt3.yaml): GPT-2-XL shape (48 x 1600 x 25 heads, T=1024,
V=50257), micro-batch 2, global batch 128 sequences (64 micro-batches) per optimizer step, 1F1B over N pipeline
stages (N = --gpus; N=1 runs all 50 stage layers on one GPU), fp32 AdamW.  Synthetic wikitext-2-shaped tokens,
seed-42 HF-style initial weights.  One "step" = pipeline.train() + DP all-reduce + optimizer step, i.e.
``OobleckEngine._train_step`` (oobleck/execution/engine.py:645-649).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODELS = {
    # examples/gpt3.yaml:9-22 (GPT-2-XL shape)
    "gpt2-xl": dict(model_args=dict(n_embd=1600, n_head=25, num_hidden_layers=48, n_positions=1024),
                    microbatch=2, global_batch=128),
    # BASELINE configs[1]: GPT-2 124M, micro-batch 4
    "gpt2": dict(model_args=dict(n_embd=768, n_head=12, num_hidden_layers=12, n_positions=1024),
                 microbatch=4, global_batch=128),
    # tiny shape for smoke runs
    "tiny": dict(model_args=dict(n_embd=128, n_head=2, num_hidden_layers=2, n_positions=128, vocab_size=1000),
                 microbatch=2, global_batch=8),
}
VOCAB = 50257


def flops_per_token(E: int, L: int, T: int, V: int) -> float:
    """fwd+bwd, no recompute, full T x T attention as the reference computes it."""
    return 6.0 * (12 * L * E * E + E * V) + 12.0 * L * T * E


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return {"bf16_tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))),
                    "hbm_gbs": float(d.get("hbm_gbs")), "source": "measured"}
        except Exception:  # noqa: BLE001
            pass
    # /opt/skills/guides/B200_PROFILING.md fallback (sustained figure: the GEMMs are timed inside a long step)
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
class CpuOracleSample:
    """Bounded CPU sample of the workload: fwd+bwd of ONE sequence (1 x T tokens) through the embedding layer,
    ``nblocks`` of the L transformer blocks and the head, timed per layer kind and scaled to the full depth
    (t = t_embed + L * t_block + t_head).  All L blocks are identical, so the scaling is exact up to cache effects;
    running all 48 GPT-2-XL blocks on the host costs minutes per sample."""

    def __init__(self, cfg, nblocks: int = 2):
        from oracle import gpt2 as og
        ma = cfg["model_args"]
        self.d = og.GPT2Dims(n_embd=ma["n_embd"], n_head=ma["n_head"], n_layer=ma["num_hidden_layers"],
                             n_positions=ma["n_positions"], vocab_size=ma.get("vocab_size", VOCAB))
        self.og = og
        self.cores = min(os.cpu_count() or 1, 64)
        torch.set_num_threads(self.cores)
        small = og.GPT2Dims(n_embd=self.d.n_embd, n_head=self.d.n_head, n_layer=min(nblocks, self.d.n_layer),
                            n_positions=self.d.n_positions, vocab_size=self.d.vocab_size)
        self.nblocks = small.n_layer
        self.layers = og.build_layers(small)
        og.init_layers_(self.layers)
        self.sample = (f"fwd+bwd of 1 sequence x {self.d.n_positions} tokens through embedding + {self.nblocks} of "
                       f"{self.d.n_layer} blocks + head, block time scaled to {self.d.n_layer} blocks")

    def run_once(self, index: int) -> float:
        """seconds for the full-depth model, extrapolated from this sample"""
        og, d = self.og, self.d
        batch = og.synthetic_batch(1, d.n_positions, d.vocab_size, index=index)
        x = (batch["input_ids"], batch["attention_mask"], batch["labels"])
        t_fwd = []
        for l in self.layers:
            t0 = time.perf_counter()
            x = l(*x)
            t_fwd.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        x[0].backward()
        t_bwd = time.perf_counter() - t0
        fwd_block = sum(t_fwd[1:-1]) / self.nblocks
        fwd_rest = t_fwd[0] + t_fwd[-1]
        # backward is timed as a whole: attribute it to blocks / rest in proportion to their forward cost
        share_blocks = sum(t_fwd[1:-1]) / sum(t_fwd)
        bwd_block = t_bwd * share_blocks / self.nblocks
        bwd_rest = t_bwd * (1 - share_blocks)
        return fwd_rest + bwd_rest + d.n_layer * (fwd_block + bwd_block)


def run_reference(args, cfg):
    """CPU arm: the oracle port of the reference's per-stage fwd/bwd path on the host cores (the reference itself
    hard-codes cuda/nccl/fused AdamW and cannot be installed here -- deepspeed, accelerate, HF-fx, cppcoro, oneTBB
    are all missing).  One step = one bounded sample (see CpuOracleSample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    smp = CpuOracleSample(cfg)
    times = []
    for it in range(args.warmup + args.steps):
        dt = smp.run_once(it)
        if it >= args.warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    value = smp.d.n_positions / sec
    print(json.dumps({
        "impl": "reference", "metric": "training_tokens_per_s", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} 1F1B train step: {cfg['model_args']['num_hidden_layers'] + 2} stage layers, "
                               f"micro-batch {cfg['microbatch']}, {cfg['global_batch'] // cfg['microbatch']} "
                               f"micro-batches/step, T={cfg['model_args']['n_positions']} -- oracle port of the "
                               "reference's torch path on the host cores, one bounded sample per step",
                   "global_batch": cfg["global_batch"], "seq_len": cfg["model_args"]["n_positions"],
                   "sample": smp.sample},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": smp.cores, "kind": "port", "sample": smp.sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline_sample(cfg, budget_s: float = 20.0):
    """Bounded CPU sample on rank 0 (N=1 only)."""
    smp = CpuOracleSample(cfg)
    smp.run_once(0)   # warm-up (allocator, thread pool)
    t_all, n, wall = 0.0, 0, time.perf_counter()
    while n < 1 or (time.perf_counter() - wall < budget_s and n < 3):
        t_all += smp.run_once(n + 1)
        n += 1
    return {"value": n * smp.d.n_positions / t_all, "unit": "tokens/s", "cores": smp.cores, "kind": "port",
            "sample": f"{n} x ({smp.sample})"}


# ---------------------------------------------------------------------------------------------------------------
def _quiet_nccl():
    # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (printed to stdout at
    # NCCL_DEBUG=VERSION and above) out of it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"


def run_ours(args, cfg):
    import torch.distributed as dist

    from oobleck_b200 import lib as L
    from oobleck_b200.execution.dataloader import SyntheticTokenDataset
    from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"launch with torchrun --nproc-per-node {args.gpus} (WORLD_SIZE={world})"
    _quiet_nccl()
    from oobleck_b200.execution import layer as _layer
    if args.bwd_fp16 is not None:
        _layer.DEFAULT_BWD_FP16 = bool(args.bwd_fp16)
    bwd_fp16 = bool(_layer.DEFAULT_BWD_FP16) and args.nsplit == 3
    from oobleck_b200.execution import pipeline as _pipeline
    if args.fb_overlap is not None:
        _pipeline.FB_OVERLAP = bool(args.fb_overlap)
    torch.cuda.set_device(local_rank)
    L.load()
    ma = cfg["model_args"]
    T, vocab = ma["n_positions"], ma.get("vocab_size", VOCAB)
    mb, gb = cfg["microbatch"], cfg["global_batch"]
    oargs = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=gb, steps=args.steps),
                             model=ModelArguments(model_name="gpt2", model_tag=args.model, model_args=dict(ma)))
    n_samples = max(2334, gb * (args.steps + args.warmup + 2) * 2)
    dataset = SyntheticTokenDataset(num_samples=n_samples, seq_len=T, vocab_size=vocab)
    dataset.to_device(torch.device("cuda", local_rank))

    transport_cls = None
    if world > 1:
        from oobleck_b200.execution.p2p import NvlinkRingTransport
        transport_cls = NvlinkRingTransport
    engine = OobleckEngine(local_rank, world, 1, None, oargs, dataset=dataset, nsplit=args.nsplit,
                           transport_cls=transport_cls, device_resident=True)
    engine.initialize_distributed()
    engine.instantiate_pipelines(gb // mb)
    loader = engine._pipeline._dataloader
    is_last = engine._pipeline.is_last_stage()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps: int, device_resident: bool, read_loss: bool, time_gemms: bool = False):
        loader.device_resident = device_resident
        engine._pipeline.reset_iterator()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        losses = []
        e0.record()
        for i in range(nsteps):
            if time_gemms and i == nsteps - 1:
                # per-launch GEMM durations for the roofline: this one step keeps every kernel on the compute stream
                # (no wgrad side stream), so a launch's event pair times that launch alone
                L.call("oob_side_stream_enable", 0)
                fb_saved, _pipeline.FB_OVERLAP = _pipeline.FB_OVERLAP, False
                L.call("oob_gemm_timing_begin")
            engine._train_step()
            if time_gemms and i == nsteps - 1:
                L.call("oob_side_stream_enable", args.side_stream)
                _pipeline.FB_OVERLAP = fb_saved
            if read_loss and is_last:
                losses.append(float(engine._pipeline.execution.total_loss.item()))   # D2H of the step's result
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, losses

    L.call("oob_side_stream_enable", args.side_stream)
    for _ in range(1):
        timed(args.warmup, True, False)          # W untimed warm-up steps
    sampler = ClockSampler(local_rank)
    launches0 = L.load().oob_launch_count()
    sampler.start()
    ms, _ = timed(args.steps, True, False, time_gemms=True)
    clocks = sampler.stop()
    launches = L.load().oob_launch_count() - launches0
    g_ms, g_fl, g_ex, g_n = C.c_double(), C.c_double(), C.c_double(), C.c_long()
    L.call("oob_gemm_timing_end", C.byref(g_ms), C.byref(g_fl), C.byref(g_ex), C.byref(g_n))
    ms_e2e, losses = timed(args.steps, False, True)

    last_loss = losses[-1] if losses else None
    if world > 1:   # the loss lives on the last stage; rank 0 prints
        t = torch.tensor([last_loss if last_loss is not None else float("-inf")], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        last_loss = float(t.item())
    tokens_per_step = gb * T
    value = tokens_per_step * args.steps / (ms / 1e3)
    e2e = tokens_per_step * args.steps / (ms_e2e / 1e3)
    peaks = measured_peaks()
    E, Lh = ma["n_embd"], ma["num_hidden_layers"]
    fpt = flops_per_token(E, Lh, T, vocab)
    exec_tflops = (g_ex.value / (g_ms.value / 1e3) / 1e12) if g_ms.value > 0 else None
    gemm_tflops = (g_fl.value / (g_ms.value / 1e3) / 1e12) if g_ms.value > 0 else None
    out = {
        "metric": "training_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None,
        "dtype": (("f32 (fp16 x2 planes forward and backward (loss-scaled) on tcgen05, fp32 accumulate + promotion)"
                   if bwd_fp16 else
                   "f32 (fp16 x2 planes forward / bf16 x3 planes backward on tcgen05, fp32 accumulate + promotion)")
                  if args.nsplit == 3 else "split-bf16 x%d on tcgen05, fp32 accumulate" % args.nsplit),
        "data": "synthetic",
        "config": {"workload": f"{args.model} 1F1B train step: {Lh + 2} stage layers over {world} stage(s), "
                               f"micro-batch {mb}, {gb // mb} micro-batches/step, T={T}, AdamW",
                   "global_batch": gb, "seq_len": T, "parallelism": f"pp{world}", "nsplit": args.nsplit,
                   "wgrad_side_stream": bool(args.side_stream), "bwd_fp16": bwd_fp16,
                   "fb_overlap": bool(_pipeline.FB_OVERLAP),
                   "l2": "working set per step (>6 GB of weights) far exceeds the 126 MB L2; no flush needed"},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": 2 * 8 * gb * T,
                "d2h_bytes_per_step": 4, "loss": last_loss},
        "model_flops_fraction_of_bf16_peak": value * fpt / (world * peaks["bf16_tflops"] * 1e12),
        "roofline": {
            "bound": "tensor", "kernel": "gemm_bf16x3_kernel (tcgen05, all GEMM launches of the last timed step)",
            "achieved": gemm_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": (gemm_tflops / peaks["bf16_tflops"]) if gemm_tflops else None,
            "peak_source": peaks["source"] + " (cuBLAS bf16, sustained)",
            "tensor_products_per_algorithmic_flop": (g_ex.value / g_fl.value) if g_fl.value else None,
            "products_note": "GEMMs on fp16 x 2 planes issue 3 products per MAC, on bf16 x 3 planes 6",
            "executed_tensor_tflops": exec_tflops,
            "executed_frac_of_peak": (exec_tflops / peaks["bf16_tflops"]) if exec_tflops else None,
            "launches_timed": int(g_n.value),
            "timing_note": "the last timed step runs without the wgrad side stream and without forward/backward "
                           "stream overlap, so each GEMM launch is timed alone",
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (forward FC GEMM 2048x6400x1600 on fp16 pairs,
            # bias + GELU epilogue writing fp32 + 5 planes) from profiles/r01_ncu_gemm_fwdfc_v5.txt: 60.7 MB read +
            # 130.9 MB written; algorithmic bytes of that launch: 54 MB of operand planes + 183 MB of outputs
            "traffic": 191.6e6 if args.model == "gpt2-xl" and args.nsplit == 3 else None,
            "traffic_note": "bytes/launch, ncu --set full, forward-FC launch; `achieved` aggregates all GEMM shapes",
        },
    }
    if world == 1:
        out["cpu_baseline"] = cpu_baseline_sample(cfg)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="gpt2-xl", choices=sorted(MODELS))
    ap.add_argument("--nsplit", type=int, default=3, choices=[1, 2, 3])
    ap.add_argument("--bwd-fp16", type=int, default=None, choices=[0, 1],
                    help="backward GEMMs on loss-scaled fp16 pairs (3 products) instead of bf16 x 3 (6); default: the "
                         "library default (oobleck_b200.execution.layer.DEFAULT_BWD_FP16)")
    ap.add_argument("--fb-overlap", type=int, default=None, choices=[0, 1],
                    help="forward(i+1) / backward(i) on two streams; default: oobleck_b200.execution.pipeline.FB_OVERLAP")
    ap.add_argument("--side-stream", type=int, default=1, choices=[0, 1],
                    help="0: weight-gradient kernels stay on the compute stream (A/B of the overlap)")
    args = ap.parse_args()
    cfg = MODELS[args.model]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
