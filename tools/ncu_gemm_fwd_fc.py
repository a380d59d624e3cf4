"""ncu driver: the forward FC GEMM of GPT-2-XL exactly as oob_block_forward launches it (fp16 pair operands, bias +
GELU epilogue writing fp32 pre-activation + 5 planes)."""
import sys

import torch

sys.path.insert(0, ".")
from oobleck_b200 import lib as L  # noqa: E402
from oobleck_b200 import ops  # noqa: E402

M, E = 2048, 1600
x = ops.split(torch.randn(M, E, device="cuda"), nplanes=5)
w = ops.split(torch.randn(E, 4 * E, device="cuda") * 0.02, nplanes=5)
bias = torch.randn(4 * E, device="cuda")
d = torch.empty(M, 4 * E, device="cuda")
gp = ops.new_planes(M, 4 * E, 5)
for _ in range(4):
    ops.gemm(x, False, w, True, M, 4 * E, E, nsplit=2, d=d, bias=bias, act=L.ACT_GELU, planes_out=gp, a_fp16=True,
             b_fp16=True)
torch.cuda.synchronize()
