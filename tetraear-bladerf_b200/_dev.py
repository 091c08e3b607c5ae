"""Device-buffer helpers: PyTorch tensors are only the carrier of device memory and streams."""
import numpy as np
import torch


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def to_dev(a, device, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype))
    return t.to(f"cuda:{device}", non_blocking=False)


def empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=f"cuda:{device}")


def zeros(shape, dtype, device):
    return torch.zeros(shape, dtype=dtype, device=f"cuda:{device}")
