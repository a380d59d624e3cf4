"""A/B sweep of GEMM knobs (env vars read by the library) on the GPT-2-XL shapes; one subprocess per setting."""
import os
import subprocess
import sys

CODE = r'''
import sys, torch
sys.path.insert(0, ".")
from oobleck_b200 import ops
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(2048, 6400, 1600), (2048, 1600, 6400), (2048, 4800, 1600)]:
    A, B = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    ap, bp = ops.split(A), ops.split(B)
    d = torch.empty(M, N, device="cuda")
    for ns in (3, 1):
        ms = t(lambda: ops.gemm(ap, False, bp, True, M, N, K, nsplit=ns, d=d))
        print(f"  M{M} N{N} K{K} nsplit={ns}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.1f} TF", flush=True)
torch.backends.cuda.matmul.allow_tf32 = False
A, B = torch.randn(2048, 1600, device="cuda"), torch.randn(1600, 6400, device="cuda")
ms = t(lambda: A @ B)
print(f"  torch fp32 2048x6400x1600: {ms*1e3:.1f} us {2*2048*6400*1600/ms/1e9:.1f} TF", flush=True)
a16, b16 = A.bfloat16(), B.bfloat16()
ms = t(lambda: a16 @ b16)
print(f"  torch bf16 2048x6400x1600: {ms*1e3:.1f} us {2*2048*6400*1600/ms/1e9:.1f} TF", flush=True)
'''
for env in sys.argv[1:] or ["OOB_GEMM_CHUNK_KB=0"]:
    e = dict(os.environ)
    for kv in env.split(","):
        k, v = kv.split("=")
        e[k] = v
    print("==", env, flush=True)
    subprocess.run([sys.executable, "-c", CODE], env=e, check=False)
