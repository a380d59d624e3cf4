"""Oracle: the GPT-2 stage layers the reference runs, as plain torch fp32 (CPU or CUDA eager).

The reference shards a HF ``GPT2LMHeadModel`` with torch.fx at the split points of
oobleck/module/sharding.py:15-18 (``transformer.h.<i>`` ... ``transformer.ln_f``), which gives
``L + 2`` stage layers (tests/module/test_model.py:22):

    layer 0      : wte[input_ids] + wpe[arange(T)]  (+ dropout)           -> EmbeddingLayer
    layer 1..L   : one ``GPT2Block``                                       -> BlockLayer
    layer L+1    : ln_f, lm_head (no bias), shift, CrossEntropyLoss(mean)  -> HeadLayer

The arithmetic itself lives in the third-party ``transformers`` package (>=4.29, environment.yml:28);
it is restated here operation by operation and pinned against the installed transformers 5.5
``GPT2LMHeadModel`` in tests/test_oracle_gpt2.py (loss and all gradients).

Each layer maps a tuple to a tuple like the fx shards do (sharding.py:86-96 threads every
later-used value through each boundary).  The wire tuple is ``(hidden[mb,T,E] f32, labels[mb,T] i64)``:
``attention_mask`` is all ones by construction of ``group_texts`` (execution/dataset.py:183-206), so the
additive mask term HF derives from it is identically zero and is not carried.

Parameter order inside each layer == ``layer.parameters()`` order of the HF module (that is the order
FSDP's FlatParamHandle flattens them in, layer.py:96-111), so flat parameter vectors are
interchangeable between the oracle and the CUDA engine.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass(frozen=True)
class GPT2Dims:
    n_embd: int = 768
    n_head: int = 12
    n_layer: int = 12
    n_positions: int = 1024
    vocab_size: int = 50257
    layer_norm_epsilon: float = 1e-5

    @property
    def head_dim(self) -> int:
        return self.n_embd // self.n_head


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    # transformers.activations.NewGELUActivation
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class EmbeddingLayer(nn.Module):
    """fx shard 0: everything before ``transformer_h_0``."""

    def __init__(self, d: GPT2Dims):
        super().__init__()
        self.d = d
        self.wte = nn.Parameter(torch.empty(d.vocab_size, d.n_embd))
        self.wpe = nn.Parameter(torch.empty(d.n_positions, d.n_embd))

    def forward(self, input_ids, attention_mask, labels):
        T = input_ids.shape[-1]
        pos = torch.arange(T, device=input_ids.device)
        hidden = self.wte[input_ids] + self.wpe[pos].unsqueeze(0)
        return hidden, labels


class BlockLayer(nn.Module):
    """fx shard i (1..L): one HF GPT2Block, eager attention."""

    def __init__(self, d: GPT2Dims):
        super().__init__()
        E = d.n_embd
        self.d = d
        # order == GPT2Block.parameters()
        self.ln_1_w = nn.Parameter(torch.empty(E))
        self.ln_1_b = nn.Parameter(torch.empty(E))
        self.c_attn_w = nn.Parameter(torch.empty(E, 3 * E))  # Conv1D: [in, out]
        self.c_attn_b = nn.Parameter(torch.empty(3 * E))
        self.c_proj_w = nn.Parameter(torch.empty(E, E))
        self.c_proj_b = nn.Parameter(torch.empty(E))
        self.ln_2_w = nn.Parameter(torch.empty(E))
        self.ln_2_b = nn.Parameter(torch.empty(E))
        self.c_fc_w = nn.Parameter(torch.empty(E, 4 * E))
        self.c_fc_b = nn.Parameter(torch.empty(4 * E))
        self.mlp_proj_w = nn.Parameter(torch.empty(4 * E, E))
        self.mlp_proj_b = nn.Parameter(torch.empty(E))

    def forward(self, hidden, labels):
        d = self.d
        B, T, E = hidden.shape
        H, hd = d.n_head, d.head_dim
        x = hidden
        h = F.layer_norm(x, (E,), self.ln_1_w, self.ln_1_b, d.layer_norm_epsilon)
        qkv = torch.addmm(self.c_attn_b, h.view(-1, E), self.c_attn_w).view(B, T, 3 * E)
        q, k, v = qkv.split(E, dim=2)
        q = q.view(B, T, H, hd).transpose(1, 2)
        k = k.view(B, T, H, hd).transpose(1, 2)
        v = v.view(B, T, H, hd).transpose(1, 2)
        att = torch.matmul(q, k.transpose(-1, -2)) / (hd ** 0.5)
        causal = torch.tril(torch.ones(T, T, dtype=torch.bool, device=x.device))
        att = torch.where(causal, att, torch.full([], torch.finfo(att.dtype).min, dtype=att.dtype, device=x.device))
        att = F.softmax(att, dim=-1)
        o = torch.matmul(att, v).transpose(1, 2).contiguous().view(B, T, E)
        a = torch.addmm(self.c_proj_b, o.view(-1, E), self.c_proj_w).view(B, T, E)
        x = x + a
        h = F.layer_norm(x, (E,), self.ln_2_w, self.ln_2_b, d.layer_norm_epsilon)
        f = torch.addmm(self.c_fc_b, h.view(-1, E), self.c_fc_w)
        f = gelu_new(f)
        m = torch.addmm(self.mlp_proj_b, f, self.mlp_proj_w).view(B, T, E)
        x = x + m
        return x, labels


class HeadLayer(nn.Module):
    """fx shard L+1: ln_f + lm_head + shifted mean cross entropy.  Returns ``(loss,)``-first tuple
    (pipeline.py:191: ``self._loss = outputs[0]``)."""

    def __init__(self, d: GPT2Dims):
        super().__init__()
        self.d = d
        self.ln_f_w = nn.Parameter(torch.empty(d.n_embd))
        self.ln_f_b = nn.Parameter(torch.empty(d.n_embd))
        self.lm_head_w = nn.Parameter(torch.empty(d.vocab_size, d.n_embd))  # untied copy (README.md:99)

    def forward(self, hidden, labels):
        d = self.d
        h = F.layer_norm(hidden, (d.n_embd,), self.ln_f_w, self.ln_f_b, d.layer_norm_epsilon)
        logits = F.linear(h, self.lm_head_w)
        shift_logits = logits[..., :-1, :].contiguous()
        shift_labels = labels[..., 1:].contiguous()
        loss = F.cross_entropy(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.view(-1))
        return loss, logits


def build_layers(d: GPT2Dims) -> list[nn.Module]:
    return [EmbeddingLayer(d)] + [BlockLayer(d) for _ in range(d.n_layer)] + [HeadLayer(d)]


def init_layers_(layers: list[nn.Module], seed: int = 42) -> None:
    """Deterministic HF-style init (model.py:56-57 seeds 42; SURVEY 8(d): N(0,0.02), LN = (1,0),
    biases 0, residual projections N(0, 0.02/sqrt(2L))).  Each layer draws from its own generator
    keyed by (seed, layer index) so that any subset of layers can be materialised independently --
    a stage does not need the whole model to reproduce its weights."""
    n_layer = sum(isinstance(l, BlockLayer) for l in layers)
    for idx, layer in enumerate(layers):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        for name, p in layer.named_parameters():
            with torch.no_grad():
                if name.startswith("ln_") and name.endswith("_w"):
                    p.fill_(1.0)
                elif name.endswith("_b"):
                    p.zero_()
                elif name in ("c_proj_w", "mlp_proj_w"):
                    p.copy_(torch.randn(p.shape, generator=g) * (0.02 / math.sqrt(2 * n_layer)))
                else:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)


def flat_params(layer: nn.Module) -> torch.Tensor:
    return torch.cat([p.detach().reshape(-1) for p in layer.parameters()])


def flat_grads(layer: nn.Module) -> torch.Tensor:
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                      for p in layer.parameters()])


def load_flat_(layer: nn.Module, flat: torch.Tensor) -> None:
    off = 0
    with torch.no_grad():
        for p in layer.parameters():
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n
    assert off == flat.numel()


def synthetic_batch(mb: int, T: int, vocab: int, seed: int = 0, index: int = 0):
    """wikitext-2-shaped synthetic tokens; labels = input_ids (dataset.py:183-202)."""
    g = torch.Generator().manual_seed(seed * 7919 + index)
    ids = torch.randint(0, vocab, (mb, T), generator=g, dtype=torch.int64)
    return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ids.clone()}
