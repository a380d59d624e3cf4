"""Is `tcgen05.mma.kind::f16` legal with A = fp16 and B = bf16 (independent a_format / b_format fields)?

Run in its own process (an illegal instruction would kill the CUDA context).  Single-plane GEMM, A from the fp16 pair
of a 5-plane buffer, B from the bf16 planes; compared with the fp64 product of the rounded operands."""
import sys

import torch

sys.path.insert(0, ".")
from oobleck_b200 import ops  # noqa: E402

M, N, K = 256, 256, 512
A = torch.randn(M, K, device="cuda")
B = torch.randn(N, K, device="cuda")
ap, bp = ops.split(A, nplanes=5), ops.split(B, nplanes=5)
a16 = ap[3].view(torch.float16).double()
b16 = bp[0].double()
for name, kw, ref in [("A fp16 x B bf16", dict(a_fp16=True), a16 @ b16.t()),
                      ("A bf16 x B fp16", dict(b_fp16=True), ap[0].double() @ bp[3].view(torch.float16).double().t())]:
    d = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(ap, False, bp, False, M, N, K, nsplit=1, d=d, **kw)
    torch.cuda.synchronize()
    err = ((d.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"{name}: max rel err vs product of the rounded operands = {err:.3e}  ->", "LEGAL" if err < 1e-5 else "WRONG")
