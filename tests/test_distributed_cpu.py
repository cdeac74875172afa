"""World-size-2 gloo test (CPU) of the sharding logic in socks_b200/distributed.py: rank 0 "fits" (oracle),
the state is broadcast, each rank evaluates its block of test points, the gathered result equals the
single-process result (to 1e-13: numpy's BLAS is not bit-stable across column blockings; the CUDA path is, which the
GPU tests check).  The per-block evaluator is the CPU oracle (test infrastructure)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from socks_b200.distributed import shard_bounds


def test_shard_bounds_cover_everything():
    for m in (0, 1, 7, 250000, 250001):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(m, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import socks_oracle as orc
    from socks_b200 import synthetic as syn
    from socks_b200.distributed import broadcast_state, gather_columns, init_from_env, predict_sharded

    r, w, _ = init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    n, T, m = 120, 5, 301
    ct, tt = syn.integrator_tubes(T)
    X, U, Y = syn.integrator_sample(n, seed=3)
    pts = syn.uniform_points(m, seed=4)
    alg = orc.KernelSROracle(T, ct, tt, "FHT", 0.2, 1e-2)
    if rank == 0:
        alg.fit(X, U, Y)
        state = {"X": torch.from_numpy(alg.X.copy()), "w": torch.from_numpy(alg.w_diag.copy()),
                 "vf": torch.from_numpy(alg.value_functions.copy())}
    else:
        state = {"X": torch.empty(n, 2, dtype=torch.float64), "w": torch.empty(n, dtype=torch.float64),
                 "vf": torch.empty(T, n, dtype=torch.float64)}
    nbytes = broadcast_state(state, src=0)
    assert nbytes == 8 * (n * 2 + n + T * n)
    alg.X, alg.w_diag, alg.value_functions = state["X"].numpy(), state["w"].numpy(), state["vf"].numpy()

    def predict_fn(block):
        return torch.from_numpy(alg.predict(block.numpy()))

    full = predict_sharded(predict_fn, torch.from_numpy(pts))
    local, (s, e) = predict_sharded(predict_fn, torch.from_numpy(pts), gather=False)
    assert local.shape == (T, e - s)
    np.save(os.path.join(tmp, f"out{rank}.npy"), full.numpy())
    if rank == 0:
        np.save(os.path.join(tmp, "ref.npy"), alg.predict(pts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_predict_matches_single(tmp_path):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = np.load(tmp_path / "ref.npy")
    for r in range(2):
        got = np.load(tmp_path / f"out{r}.npy")
        assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= 1e-13
