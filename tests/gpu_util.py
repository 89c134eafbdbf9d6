"""Helpers for the -m gpu parity tests: move numpy uint arrays to torch CUDA storage and back."""
import numpy as np
import torch


def to_dev(a: np.ndarray, bits: int) -> torch.Tensor:
    if bits == 64:
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.uint32).view(np.int32)).cuda()


def to_host(t: torch.Tensor, bits: int) -> np.ndarray:
    a = t.cpu().numpy()
    return a.view(np.uint64) if bits == 64 else a.view(np.uint32).astype(np.uint64)


def to_host_signed(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().astype(np.int64)
