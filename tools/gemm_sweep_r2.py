"""The GEMM calls of one GPT-2-XL block in the DEFAULT build (fp16 pairs everywhere, pair-only plane buffers), each with
the fused epilogue the stage really attaches (csrc/stage.cu), timed alone with CUDA events; M = 2048 tokens."""
import sys

import torch

sys.path.insert(0, ".")
from oobleck_b200 import lib as L  # noqa: E402
from oobleck_b200 import ops  # noqa: E402

H2 = ops.PLANES_FP16_PAIR


def t(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


M, E = 2048, 1600
pair = lambda r, c: ops.split(torch.randn(r, c, device="cuda") * 0.05, nplanes=H2)  # noqa: E731
f32 = lambda r, c: torch.randn(r, c, device="cuda")  # noqa: E731
out2 = lambda r, c: ops.new_planes(r, c, 2)  # noqa: E731
kw = dict(nsplit=2, a_pair0=True, b_pair0=True)
total = 0.0
cases = []
X, Wqkv, Wp, Wfc, Wp2 = pair(M, E), pair(E, 3 * E), pair(E, E), pair(E, 4 * E), pair(4 * E, E)
X4 = pair(M, 4 * E)
b3, b1, b4 = torch.randn(3 * E, device="cuda"), torch.randn(E, device="cuda"), torch.randn(4 * E, device="cuda")
R, aux = f32(M, E), f32(M, 4 * E)
dY, dY4, dY3 = pair(M, E), pair(M, 4 * E), pair(M, 3 * E)
d1, d4, d3 = f32(M, E), f32(M, 4 * E), f32(M, 3 * E)
gq, gp, gf, gp2 = f32(E, 3 * E), f32(E, E), f32(E, 4 * E), f32(4 * E, E)
cases = [
    ("fwd qkv   bias -> pair planes", M, 3 * E, E, lambda: ops.gemm(X, False, Wqkv, True, M, 3 * E, E, bias=b3, planes_out=out2(M, 3 * E), planes_code=H2, **kw)),
    ("fwd proj  bias + resid -> f32", M, E, E, lambda: ops.gemm(X, False, Wp, True, M, E, E, d=d1, bias=b1, resid=R, **kw)),
    ("fwd fc    bias+GELU -> f32+pair", M, 4 * E, E, lambda: ops.gemm(X, False, Wfc, True, M, 4 * E, E, d=d4, bias=b4, act=L.ACT_GELU, planes_out=out2(M, 4 * E), planes_code=H2, **kw)),
    ("fwd proj2 bias + resid -> f32", M, E, 4 * E, lambda: ops.gemm(X4, False, Wp2, True, M, E, 4 * E, d=d1, bias=b1, resid=R, **kw)),
    ("dgrad proj2 dGELU -> f32+pair", M, 4 * E, E, lambda: ops.gemm(dY, False, Wp2, False, M, 4 * E, E, d=d4, act=L.ACT_DGELU, aux=aux, planes_out=out2(M, 4 * E), planes_code=H2, **kw)),
    ("dgrad fc    -> f32", M, E, 4 * E, lambda: ops.gemm(dY4, False, Wfc, False, M, E, 4 * E, d=d1, **kw)),
    ("dgrad proj  -> f32+pair", M, E, E, lambda: ops.gemm(dY, False, Wp, False, M, E, E, d=d1, planes_out=out2(M, E), planes_code=H2, **kw)),
    ("dgrad qkv   -> f32", M, E, 3 * E, lambda: ops.gemm(dY3, False, Wqkv, False, M, E, 3 * E, d=d1, **kw)),
    ("wgrad proj2 accumulate", 4 * E, E, M, lambda: ops.gemm(X4, True, dY, True, 4 * E, E, M, d=gp2, accumulate=True, alpha=0.5, **kw)),
    ("wgrad fc    accumulate", E, 4 * E, M, lambda: ops.gemm(X, True, dY4, True, E, 4 * E, M, d=gf, accumulate=True, alpha=0.5, **kw)),
    ("wgrad proj  accumulate", E, E, M, lambda: ops.gemm(X, True, dY, True, E, E, M, d=gp, accumulate=True, alpha=0.5, **kw)),
    ("wgrad qkv   accumulate", E, 3 * E, M, lambda: ops.gemm(X, True, dY3, True, E, 3 * E, M, d=gq, accumulate=True, alpha=0.5, **kw)),
]
flops = 0.0
for name, m, n, k, fn in cases:
    ms = t(fn)
    total += ms
    flops += 2.0 * m * n * k
    print(f"  {name:32s} {m:5d} x {n:5d} x {k:5d}: {ms * 1e3:7.1f} us  {2.0 * m * n * k / ms / 1e9:6.1f} TF alg", flush=True)
print(f"  block total: {total * 1e3:.1f} us, {flops / total / 1e9:.1f} TF algorithmic ({3 * flops / total / 1e9:.1f} TF executed)")
