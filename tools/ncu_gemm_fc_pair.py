"""One GEMM launch for ncu: forward FC of GPT-2-XL in the default build (fp16 pairs, pair-only buffers):
2048 x 6400 x 1600, bias + GELU epilogue writing the fp32 pre-activation and the two planes of GELU(x)."""
import sys

import torch

sys.path.insert(0, ".")
from oobleck_b200 import lib as L  # noqa: E402
from oobleck_b200 import ops  # noqa: E402

H2 = ops.PLANES_FP16_PAIR
M, E = 2048, 1600
X = ops.split(torch.randn(M, E, device="cuda"), nplanes=H2)
W = ops.split(torch.randn(E, 4 * E, device="cuda") * 0.02, nplanes=H2)
b = torch.randn(4 * E, device="cuda")
d = torch.empty(M, 4 * E, device="cuda")
out = ops.new_planes(M, 4 * E, 2)
for _ in range(4):
    ops.gemm(X, False, W, True, M, 4 * E, E, nsplit=2, a_pair0=True, b_pair0=True, d=d, bias=b, act=L.ACT_GELU,
             planes_out=out, planes_code=H2)
torch.cuda.synchronize()
