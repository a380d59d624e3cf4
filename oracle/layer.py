"""Checker-side stage layer: the oracle's torch modules behind the ``Layer`` interface, so that the pipeline host logic
(schedule interpreter, wire protocol, DP groups, reconfiguration) can run on CPU with gloo -- in the tests, and as the
CPU arm of bench.py (``--impl reference``: P gloo processes running a real 1F1B step over these layers, BASELINE.md
section 4).  Test infrastructure like everything under oracle/: never imported by the product package."""
from __future__ import annotations

import torch

from oobleck_b200.execution.layer import HiddenGrad
from oobleck_b200.execution.optimizer import WarmupLR
from oracle import gpt2 as og
from oracle import optim as oo


class _Handle:
    def __init__(self, flat):
        self.flat_param = flat
        self.process_group = None
        self._sharding_strategy = "NO_SHARD"


class OracleAdamW:
    def __init__(self, layers, lr, betas, eps, weight_decay=0.01):
        self.layers = list(layers)
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        self.state = {}
        self._step = 0

    def step(self):
        self._step += 1
        g = self.param_groups[0]
        for l in self.layers:
            l.sync_grads()
            l.opt_step += 1
            oo.adamw_step_(l.flat_param, l.flat_param.grad, l.exp_avg, l.exp_avg_sq, l.opt_step, g["lr"], g["betas"][0],
                           g["betas"][1], g["eps"], g["weight_decay"])
            og.load_flat_(l.module, l.flat_param)

    def zero_grad(self):
        for l in self.layers:
            l.module.zero_grad()
            l.flat_param.grad.zero_()


class OracleLayer:
    device_type = "cpu"

    @staticmethod
    def optimizer_factory():
        return OracleAdamW, WarmupLR

    fast_init = False      # bench.py's CPU arm: constant weights (timing does not depend on the values; drawing 1.5 G
                           # normal deviates on the host costs more than the sample itself)

    def __init__(self, layer_id, spec, process_group=None, pre_stream=None, post_stream=None, *, microbatch_size,
                 num_pipe_buffers, workspace=None, nsplit=3):
        self.layer_id = layer_id
        self.spec = spec
        self.num_pipe_buffers = num_pipe_buffers
        d = og.GPT2Dims(n_embd=spec.n_embd, n_head=spec.n_head, n_layer=spec.n_layer, n_positions=spec.n_positions,
                        vocab_size=spec.vocab_size, layer_norm_epsilon=spec.layer_norm_epsilon)
        self.module = {"embed": og.EmbeddingLayer, "block": og.BlockLayer, "head": og.HeadLayer}[spec.kind](d)
        flat = torch.full((spec.num_params,), 0.01) if self.fast_init else spec.init_flat()
        og.load_flat_(self.module, flat)
        flat.grad = torch.zeros_like(flat)
        self._param_handle = _Handle(flat)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
        self.saved = [None] * num_pipe_buffers
        self.opt_step = 0

    @classmethod
    def create_layer_from_layer(cls, existing, pg, num_pipe_buffers=None):
        if num_pipe_buffers is not None and num_pipe_buffers > existing.num_pipe_buffers:
            existing.saved += [None] * (num_pipe_buffers - existing.num_pipe_buffers)
            existing.num_pipe_buffers = num_pipe_buffers
        return existing

    def state_tensors(self):
        return [self.flat_param, self.exp_avg, self.exp_avg_sq]

    @property
    def flat_param(self):
        return self._param_handle.flat_param

    @property
    def flat_grad(self):
        self.sync_grads()
        return self._param_handle.flat_param.grad

    def sync_grads(self):
        self._param_handle.flat_param.grad.copy_(og.flat_grads(self.module))

    def load_flat_(self, flat):
        self.flat_param.copy_(flat)
        og.load_flat_(self.module, flat)

    def refresh_planes(self):
        og.load_flat_(self.module, self.flat_param)

    def remove_tensors(self):
        pass

    def __call__(self, inputs, buffer_id=0, total_loss=None):
        ins = tuple(t.detach().requires_grad_(t.is_floating_point()) for t in inputs)
        outs = self.module(*ins)
        self.saved[buffer_id] = (ins, outs)
        if self.spec.kind == "head":
            if total_loss is not None:
                total_loss += outs[0].detach()
            return outs[0].detach(), outs[1].detach()
        return tuple(o.detach().requires_grad_(o.is_floating_point()) for o in outs)

    def backward(self, buffer_id, grad):
        ins, outs = self.saved[buffer_id]
        if self.spec.kind == "head":
            outs[0].backward()
        else:
            torch.autograd.backward(outs[0], grad.grad.view_as(outs[0]))
        self.saved[buffer_id] = None
        x = ins[0]
        return HiddenGrad(x.grad) if x.is_floating_point() else None

    def reduce_gradients(self, process_groups):
        self.sync_grads()
        for _, pg in process_groups.items():
            torch.distributed.all_reduce(self._param_handle.flat_param.grad, group=getattr(pg, "group", pg))
        # write the reduced gradient back into the module
        off = 0
        for p in self.module.parameters():
            n = p.numel()
            p.grad = self._param_handle.flat_param.grad[off:off + n].view_as(p).clone()
            off += n
