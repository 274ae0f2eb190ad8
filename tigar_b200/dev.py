"""
Device-memory plumbing: torch tensors own HBM allocations and the CUDA stream;
every computation goes through the C-ABI in libtigar_b200.so.  There is no CPU
path: anything that needs device memory raises if CUDA is unavailable.
"""
import numpy as np
import torch

F64 = torch.float64
I32 = torch.int32
I64 = torch.int64
U8 = torch.uint8


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("tigar_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def empty(n, dtype=F64):
    return torch.empty(int(n), dtype=dtype, device=device())


def zeros(n, dtype=F64):
    return torch.zeros(int(n), dtype=dtype, device=device())


def from_np(a, dtype=None):
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device())


def to_np(t):
    return t.detach().cpu().numpy()


def ptr(t):
    """Raw device pointer (None for absent / empty)."""
    if t is None:
        return None
    return t.data_ptr() or None


def stream():
    return torch.cuda.current_stream().cuda_stream or None


def sync():
    torch.cuda.current_stream().synchronize()
