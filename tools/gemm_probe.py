"""GPU probe: run the tcgen05 GEMM in every operand-major combination and print error statistics.
Usage (on the GPU box): python tools/gemm_probe.py [--perf]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from oobleck_b200 import ops  # noqa: E402


def run(M, N, K, a_mn, b_mn, nsplit, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(K, N, device="cuda", generator=g)
    a_st = A.t().contiguous() if a_mn else A            # [K,M] if M-major else [M,K]
    b_st = B.contiguous() if b_mn else B.t().contiguous()  # [K,N] if N-major else [N,K]
    ap, bp = ops.split(a_st), ops.split(b_st)
    d = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(ap, a_mn, bp, b_mn, M, N, K, nsplit=nsplit, d=d)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double())
    err = (d.double() - ref).abs()
    scale = ref.abs().mean().item()
    nan = torch.isnan(d).sum().item()
    return err.max().item() / scale, err.mean().item() / scale, nan


def main():
    torch.manual_seed(0)
    print("device", torch.cuda.get_device_name(0))
    for (M, N, K) in [(128, 128, 64), (128, 128, 128), (256, 256, 512), (304, 200, 136), (2048, 1600, 1600), (256, 256, 6400)]:
        for a_mn in (False, True):
            for b_mn in (False, True):
                for ns in (1, 2, 3):
                    try:
                        mx, mean, nan = run(M, N, K, a_mn, b_mn, ns)
                        print(f"M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} nsplit={ns}: "
                              f"max_rel={mx:.3e} mean_rel={mean:.3e} nan={nan}", flush=True)
                    except Exception as e:  # noqa: BLE001
                        print(f"M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} nsplit={ns}: EXC {e}", flush=True)
    if "--perf" in sys.argv:
        for (M, N, K) in [(2048, 4800, 1600), (2048, 1600, 6400), (2048, 6400, 1600), (8192, 4800, 1600)]:
            for (a_mn, b_mn) in [(False, True), (False, False), (True, True)]:
                A = torch.randn(K, M, device="cuda") if a_mn else torch.randn(M, K, device="cuda")
                B = torch.randn(K, N, device="cuda") if b_mn else torch.randn(N, K, device="cuda")
                ap, bp = ops.split(A), ops.split(B)
                d = torch.empty(M, N, device="cuda")
                for ns in (1, 2, 3):
                    for _ in range(3):
                        ops.gemm(ap, a_mn, bp, b_mn, M, N, K, nsplit=ns, d=d)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        ops.gemm(ap, a_mn, bp, b_mn, M, N, K, nsplit=ns, d=d)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 10
                    print(f"perf M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} nsplit={ns}: {ms*1e3:.1f} us "
                          f"{2*M*N*K/ms/1e9:.1f} TFLOP/s algorithmic", flush=True)
        # torch fp32 (no TF32) baseline
        torch.backends.cuda.matmul.allow_tf32 = False
        for (M, N, K) in [(2048, 4800, 1600)]:
            A, B = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
            for _ in range(3):
                A @ B
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                A @ B
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"torch fp32 matmul M{M} N{N} K{K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
