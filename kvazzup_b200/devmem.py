"""Device-memory plumbing for tests and bench: torch is used only to own HBM buffers and streams."""
from __future__ import annotations

import numpy as np


def torch_mod():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA device required: kvazzup_b200 has no CPU fallback")
    return torch


def to_device(a: np.ndarray, device: str = "cuda"):
    torch = torch_mod()
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def empty_u8(n: int, device: str = "cuda"):
    torch = torch_mod()
    return torch.empty(int(n), dtype=torch.uint8, device=device)


def current_stream_ptr() -> int:
    torch = torch_mod()
    return int(torch.cuda.current_stream().cuda_stream)
