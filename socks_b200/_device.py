"""Device plumbing: torch is used ONLY to own device memory and streams (no torch math on the path)."""
import numpy as np
import torch

from socks_b200 import _lib

F64 = torch.float64


def device():
    if not torch.cuda.is_available():
        raise _lib.SocksError("socks_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def to_dev(a, dtype=F64):
    """Host array (list / int / float32 / float64 ndarray) or tensor -> contiguous device tensor."""
    dev = device()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    h = torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float64 if dtype == F64 else None))
    return h.to(dev, non_blocking=False)


def as2d(a):
    """np.array(X) as the reference does (kernel_sr.py:282-284), promoted to float64, 2-D."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        a = a.reshape(a.shape[0], -1)
    return np.ascontiguousarray(a)


def empty(*shape):
    return torch.empty(*shape, dtype=F64, device=device())


def zeros(*shape):
    return torch.zeros(*shape, dtype=F64, device=device())


def work(nbytes):
    return torch.empty((int(nbytes) + 7) // 8 + 2, dtype=F64, device=device())


def padded(n):
    return int(_lib.load().socks_padded_dim(int(n)))


def even(n):
    return int(n) + (int(n) & 1)
