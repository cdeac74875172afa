"""Device plumbing: torch is used ONLY to own device memory and streams (no torch math on the path)."""
import weakref

import numpy as np
import torch

from socks_b200 import _lib

F64 = torch.float64


# Optional stage timing (bench.py, tools/): when `stage_marks` is a list, mark(name) records a CUDA event on the
# current stream after the work enqueued so far; stage_times() turns consecutive marks into seconds.
stage_marks = None


def mark(name):
    if stage_marks is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        stage_marks.append((name, e))


def stage_times():
    """{name: seconds since the previous mark} (call after a synchronize)."""
    out = {}
    for (_, e0), (name, e1) in zip(stage_marks[:-1], stage_marks[1:]):
        out[name] = out.get(name, 0.0) + e0.elapsed_time(e1) * 1e-3
    return out


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


class _Lease:
    """Owner object of one result array: numpy keeps it (as `.base`) for as long as the array or any view of it
    is alive, so a dead weak reference to the lease means the caller has dropped the result."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.__array_interface__ = {"shape": tuple(tensor.shape), "typestr": "<f8",
                                    "data": (tensor.data_ptr(), False), "version": 3}


class HostPool:
    """Pinned host result buffers, reused only when the caller has dropped the previous result.

    `predict()` hands out `numpy.asarray(lease)`: a zero-copy ndarray on the pinned buffer whose `.base` is a
    lease object.  A buffer is free again once its lease has been garbage collected (every array and view the
    caller held is gone); otherwise a new buffer is allocated.  This keeps the reference's "fresh ndarray per
    call" semantics without a 30 MB page-faulting allocation and a pageable device-to-host copy per call."""

    def __init__(self, max_per_shape=3, max_shapes=8):
        self._bufs, self._max, self._max_shapes = {}, max_per_shape, max_shapes

    def get(self, rows, cols):
        """-> (pinned tensor, finish) where finish() returns the ndarray to hand to the caller."""
        key = (int(rows), int(cols))
        lst = self._bufs.get(key)
        if lst is None:
            if len(self._bufs) >= self._max_shapes:
                self._bufs.pop(next(iter(self._bufs)))
            lst = self._bufs[key] = []
        entry = None
        for e in lst:
            if e["lease"] is None or e["lease"]() is None:
                entry = e
                break
        if entry is None:
            entry = {"tensor": torch.empty(key, dtype=F64, pin_memory=True), "lease": None}
            if len(lst) < self._max:
                lst.append(entry)
        tensor = entry["tensor"]

        def finish():
            lease = _Lease(tensor)
            entry["lease"] = weakref.ref(lease)
            return np.asarray(lease)

        return tensor, finish


host_pool = HostPool()
_copy_streams = {}


def copy_stream():
    d = torch.cuda.current_device()
    if d not in _copy_streams:
        _copy_streams[d] = torch.cuda.Stream(device=d)
    return _copy_streams[d]
