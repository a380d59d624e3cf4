"""Tiny driver for ncu: a few launches of the dominant GEMM shapes (GPT-2-XL, micro-batch 2)."""
import sys

import torch

sys.path.insert(0, ".")
from oobleck_b200 import ops  # noqa: E402

M, E = 2048, 1600
x = ops.split(torch.randn(M, E, device="cuda"))
w = ops.split(torch.randn(E, 4 * E, device="cuda") * 0.02)
d = torch.empty(M, 4 * E, device="cuda")
dy = ops.split(torch.randn(M, 4 * E, device="cuda"))
dw = torch.zeros(E, 4 * E, device="cuda")
for _ in range(3):
    ops.gemm(x, False, w, True, M, 4 * E, E, d=d)                       # forward  FC
    ops.gemm(x, True, dy, True, E, 4 * E, M, d=dw, accumulate=True)    # wgrad    FC
torch.cuda.synchronize()
