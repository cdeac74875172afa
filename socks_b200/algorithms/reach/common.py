"""Shared host glue for the reachability algorithms (tube packing, problem codes).

The recursion steps themselves (`_fht_step`, `_tht_step`, `gym_socks/algorithms/reach/common.py:8-69`)
are fused into the CUDA kernels (`sr_step` in csrc/common.cuh).
"""
import numpy as np

from socks_b200 import _device as dv

PROBLEM_CODE = {"THT": 0, "FHT": 1}


def pack_tubes(constraint_tube, target_tube, time_horizon, dim):
    """-> device tensor [T][4][d] = (constraint.low, constraint.high, target.low, target.high).

    The bounds are taken as the spaces store them (float32 for gym Boxes) and promoted to float64,
    which is what numpy does in `points >= space.low` (`gym_socks/utils/__init__.py:42-47`)."""
    if len(constraint_tube) < time_horizon or len(target_tube) < time_horizon:
        raise IndexError("constraint_tube and target_tube must have at least time_horizon entries")
    out = np.empty((time_horizon, 4, dim), dtype=np.float64)
    for t in range(time_horizon):
        for k, space in ((0, constraint_tube[t]), (2, target_tube[t])):
            if not (hasattr(space, "low") and hasattr(space, "high")):
                raise TypeError(
                    "tubes must be Box-like objects with .low/.high (the reference's indicator_fn reads "
                    "space.low/space.high, gym_socks/utils/__init__.py:42-43)"
                )
            out[t, k] = np.broadcast_to(np.asarray(space.low, dtype=np.float64).reshape(-1), (dim,))
            out[t, k + 1] = np.broadcast_to(np.asarray(space.high, dtype=np.float64).reshape(-1), (dim,))
    return dv.to_dev(out.reshape(-1))
