"""Where does the 3-product (fp16 pair) GEMM spend its time?  FC / FC2 / QKV shapes, epilogue variants, debug bits
(OOB_GEMM_DEBUG: 1 skip output stores, 2 skip TMA, 4 skip MMAs); one subprocess per env setting."""
import os
import subprocess
import sys

CODE = r'''
import sys, torch
sys.path.insert(0, ".")
from oobleck_b200 import ops, lib as L
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(2048, 6400, 1600), (2048, 1600, 6400), (2048, 4800, 1600)]:
    A, B = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda") * 0.02
    ap, bp = ops.split(A, nplanes=5), ops.split(B, nplanes=5)
    d = torch.empty(M, N, device="cuda"); bias = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda")
    p5, p3 = ops.new_planes(M, N, 5), ops.new_planes(M, N, 3)
    h = dict(nsplit=2, a_fp16=True, b_fp16=True)
    cases = [("bf16x3 d", dict(nsplit=3, d=d)), ("fp16x2 d", dict(d=d, **h)),
             ("fp16x2 d+bias+resid", dict(d=d, bias=bias, resid=R, **h)),
             ("fp16x2 d+bias+GELU+5planes", dict(d=d, bias=bias, act=L.ACT_GELU, planes_out=p5, **h)),
             ("fp16x2 bias+3planes (no d)", dict(bias=bias, planes_out=p3, **h)),
             ("bf16x3 d+bias+GELU+3planes", dict(nsplit=3, d=d, bias=bias, act=L.ACT_GELU, planes_out=p3))]
    for name, kw in cases:
        ms = t(lambda: ops.gemm(ap, False, bp, True, M, N, K, **kw))
        print(f"  M{M} N{N} K{K} {name:32s}: {ms*1e3:7.1f} us {2*M*N*K/ms/1e9:6.1f} TF alg", flush=True)
'''
for env in sys.argv[1:] or ["OOB_GEMM_DEBUG=0"]:
    e = dict(os.environ)
    for kv in env.split(","):
        k, v = kv.split("=")
        e[k] = v
    print("==", env, flush=True)
    subprocess.run([sys.executable, "-c", CODE], env=e, check=False)
