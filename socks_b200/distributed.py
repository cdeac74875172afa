"""Single-node multi-GPU sharding of the hot path (north_star item 4).

One process per GPU (torchrun).  The factorisation and the fit recursion run on ONE rank; the fitted state
(`X`, `w = diag(W)`, value functions; plus `CUA` for KernelMaximalSR) is broadcast once with
`torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests), then test points are split in
contiguous blocks and every rank evaluates its block independently.  There is no data-path collective
(test points are independent columns, `gym_socks/utils/__init__.py:22`), and because the predict kernels sum
the sample axis in fixed 2048-sample splits, a sharded run is bit-identical to the single-GPU run.

The functions take/return torch tensors on whatever device the process group uses, so the same code is
exercised by the world_size-2 gloo tests on CPU.
"""
import os

import torch
import torch.distributed as dist

__all__ = ["init_from_env", "shard_bounds", "broadcast_state", "gather_columns", "predict_sharded", "SocksComm"]


class SocksComm:
    """The library's own NCCL communicator (csrc/compose.cu: socks_ctx_* / socks_comm_init / socks_bcast /
    socks_allgather): what a non-torch binder of the C ABI uses for the path's two collectives.  torch.distributed is
    only the rendezvous here (it carries the 128-byte NCCL unique id from rank 0 to the others)."""

    def __init__(self):
        import ctypes

        from socks_b200 import _lib

        self._lib, self._ct = _lib, ctypes
        self.rank, self.world = _rank(), _world()
        self.ctx = ctypes.c_void_p()
        _lib.call("socks_ctx_create", torch.cuda.current_device(), ctypes.byref(self.ctx))
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (ctypes.c_char * 128)()
            _lib.call("socks_comm_unique_id", buf)
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if self.world > 1:
            dev_id = ident.cuda() if dist.get_backend() == "nccl" else ident
            dist.broadcast(dev_id, src=0)
            ident = dev_id.cpu()
        raw = (ctypes.c_char * 128).from_buffer_copy(bytes(ident.numpy().tobytes()))
        _lib.call("socks_comm_init", self.ctx, self.rank, self.world, raw)

    def broadcast_state(self, tensors, src=0):
        """In-place broadcast of every tensor of the dict on the context stream; returns the bytes sent."""
        torch.cuda.current_stream().synchronize()  # the tensors were produced on torch's stream
        nbytes = 0
        for key in sorted(tensors):
            t = tensors[key]
            assert t.is_contiguous()
            self._lib.call("socks_bcast", self.ctx, t.data_ptr(), t.numel() * t.element_size(), src)
            nbytes += t.numel() * t.element_size()
        self._lib.call("socks_ctx_synchronize", self.ctx)
        return nbytes

    def allgather(self, local):
        """(world * local.numel(),) tensor of every rank's contiguous buffer, rank-major."""
        torch.cuda.current_stream().synchronize()
        out = torch.empty(self.world * local.numel(), dtype=local.dtype, device=local.device)
        self._lib.call("socks_allgather", self.ctx, local.data_ptr(), out.data_ptr(), local.numel() * local.element_size())
        self._lib.call("socks_ctx_synchronize", self.ctx)
        return out

    def close(self):
        if self.ctx:
            self._lib.call("socks_ctx_destroy", self.ctx)
            self.ctx = None


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_bounds(m, world=None, rank=None):
    """Contiguous block [start, end) of `m` test points owned by `rank`; sizes differ by at most one."""
    world = _world() if world is None else world
    rank = _rank() if rank is None else rank
    base, rem = divmod(int(m), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_state(tensors, src=0):
    """Broadcast a dict of equally-shaped-on-all-ranks tensors from `src` (in place). Returns bytes sent."""
    nbytes = 0
    if _world() > 1:
        for key in sorted(tensors):
            dist.broadcast(tensors[key], src=src)
            nbytes += tensors[key].numel() * tensors[key].element_size()
    return nbytes


def gather_columns(local, m_total):
    """All-gather column blocks (T, m_rank) -> (T, m_total) on every rank (blocks as in shard_bounds)."""
    world = _world()
    if world == 1:
        return local
    T = local.shape[0]
    width = -(-int(m_total) // world)
    pad = torch.zeros(T, width, dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.empty(T, m_total, dtype=local.dtype, device=local.device)
    for r in range(world):
        s, e = shard_bounds(m_total, world, r)
        out[:, s:e] = parts[r][:, : e - s]
    return out


def predict_sharded(predict_fn, pts, gather=True):
    """Evaluate `predict_fn` (points block -> (T, block) tensor) on this rank's block of `pts`.

    Returns the gathered (T, m) tensor if `gather` else (local block, (start, end))."""
    m = pts.shape[0]
    s, e = shard_bounds(m)
    local = predict_fn(pts[s:e])
    if not gather:
        return local, (s, e)
    return gather_columns(local, m)
