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


_cuda_ok = [False]


def require_cuda():
    if _cuda_ok[0]:                     # torch.cuda.is_available() costs up to 1 ms per call
        return
    if not torch.cuda.is_available():
        raise RuntimeError("tigar_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    _cuda_ok[0] = True


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
    """Device -> host numpy.  Large transfers go through a pinned staging tensor (torch's
    caching host allocator recycles the block): ~25 GB/s instead of the ~3 GB/s of a pageable
    copy -- the 139 MB solution vector of the 256^3 patch costs 6 ms instead of 45."""
    t = t.detach()
    if t.is_cuda and t.numel() * t.element_size() >= (1 << 20):
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()
    return t.cpu().numpy()


def ptr(t):
    """Raw device pointer (None for absent / empty)."""
    if t is None:
        return None
    return t.data_ptr() or None


def stream():
    return torch.cuda.current_stream().cuda_stream or None


def sync():
    torch.cuda.current_stream().synchronize()


class _Prof(object):
    """Live per-kernel timing for bench.py's rooflines: CUDA events on the stream the kernels
    are launched on (torch's current stream), recorded around every launch of a named kernel
    class together with its algorithmic bytes / flops.  Off unless ``start()`` was called."""

    def __init__(self):
        self.on = False
        self.rec = {}

    def start(self):
        self.on = True
        self.rec = {}

    class _Range(object):
        def __init__(self, prof, name, nbytes, flops):
            self.p, self.name, self.nbytes, self.flops = prof, name, nbytes, flops

        def __enter__(self):
            if self.p.on:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
            return self

        def __exit__(self, *exc):
            if self.p.on:
                self.e1.record()
                self.p.rec.setdefault(self.name, []).append((self.e0, self.e1, self.nbytes,
                                                             self.flops))
            return False

    def range(self, name, nbytes=0, flops=0):
        return self._Range(self, name, nbytes, flops)

    def stop(self):
        """-> {name: dict(launches, ms, bytes, flops)} (synchronises)."""
        self.on = False
        torch.cuda.synchronize()
        out = {}
        for name, lst in self.rec.items():
            ms = sum(e0.elapsed_time(e1) for e0, e1, _, _ in lst)
            out[name] = dict(launches=len(lst), ms=ms, bytes=sum(x[2] for x in lst),
                             flops=sum(x[3] for x in lst))
        self.rec = {}
        return out


PROF = _Prof()
