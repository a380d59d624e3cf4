"""Driver for ncu / timing of the attention kernels at the GPT-2-XL micro-batch shape (B=2, T=1024, H=25, D=64)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from oobleck_b200 import lib as L  # noqa: E402
from oobleck_b200 import ops  # noqa: E402

B, T, H, D = 2, 1024, 25, 64
E = H * D
P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
qkv = torch.randn(B * T, 3 * E, device="cuda")
out = torch.empty(B * T, E, device="cuda")
outp = ops.new_planes(B * T, E)
lse = torch.empty(B, H, T, device="cuda")
dout = torch.randn(B * T, E, device="cuda")
dqkv = torch.empty(B * T, 3 * E, device="cuda")
dqp = ops.new_planes(B * T, 3 * E)
delta = torch.empty(B, H, T, device="cuda")


H2 = int(sys.argv[1]) if len(sys.argv) > 1 else 1      # 1: fp16-pair operands (3 products), 0: bf16 x 3 (6)
CODE = ops.PLANES_FP16_PAIR if H2 else 3
qp, dop = ops.split(qkv, nplanes=CODE), ops.split(dout, nplanes=CODE)


def fwd():
    L.call("oob_attention_fwd", P(qp), qp.stride(0), H2, P(out), P(outp), outp.stride(0), 3, P(lse), B, T, H, D, S())


def bwd():
    L.call("oob_attention_bwd", P(qp), qp.stride(0), H2, P(out), P(dout), P(dop), dop.stride(0), P(lse), P(delta),
           P(dqkv), P(dqp), dqp.stride(0), CODE, B, T, H, D, S())


for fn, name, flops in [(fwd, "fwd", 4 * B * H * T * T * D / 2), (bwd, "bwd", 10 * B * H * T * T * D / 2)]:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"attention (fp16 pair={H2}) {name}: {ms*1e3:.1f} us, {flops/ms/1e9:.1f} TFLOP/s (causal algorithmic)", flush=True)
