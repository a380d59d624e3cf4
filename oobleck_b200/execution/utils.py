"""Wire enum + small helpers (oobleck/execution/utils.py:4-40)."""
from __future__ import annotations

from typing import Iterable, Union

import torch

# order is the wire format of the P2P meta handshake (utils.py:4-18); index == dtype id
ID_TO_DTYPE = [
    torch.float32,
    torch.float64,
    torch.complex64,
    torch.complex128,
    torch.float16,
    torch.bfloat16,
    torch.uint8,
    torch.int8,
    torch.int16,
    torch.int32,
    torch.int64,
    torch.bool,
]
DTYPE_TO_ID = {dtype: id_ for id_, dtype in enumerate(ID_TO_DTYPE)}


def zero_grads(inputs: Union[torch.Tensor, Iterable[torch.Tensor]]) -> None:
    """utils.py:33-40."""
    tensors = [inputs] if isinstance(inputs, torch.Tensor) else inputs
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.grad is not None:
            t.grad.data.zero_()
