"""Randomised GPU-vs-oracle parity over shapes the golden fixtures do not cover: ragged sizes around the
128/256/2048 tile and split boundaries, dimensions 1..8, horizons around the 15-step launch boundary, both
problems, time-varying tubes, regularisation from well- to ill-conditioned.  Seeds are fixed (deterministic)."""
from functools import partial

import numpy as np
import pytest

from oracle import socks_oracle as orc
from socks_b200.spaces import Box

pytestmark = pytest.mark.gpu
TOL = 1e-8

CASES = [
    # n,   m,   d, T,  problem, sigma, lam
    (1, 3, 1, 2, "THT", 0.5, 1.0),
    (2, 1, 2, 1, "FHT", 0.3, 0.5),
    (63, 65, 1, 3, "FHT", 0.2, 1e-2),
    (127, 129, 2, 15, "THT", 0.15, 1e-3),
    (128, 256, 3, 16, "FHT", 0.3, 1e-1),
    (129, 255, 4, 17, "THT", 0.4, 1.0),
    (257, 511, 5, 5, "FHT", 0.6, 1e-2),
    (640, 1000, 6, 4, "THT", 0.8, 1e-1),
    (1000, 333, 7, 6, "FHT", 1.0, 1.0),
    (2049, 300, 2, 31, "FHT", 0.1, 1e-4),   # crosses the 2048-sample split and needs three predict launches
    (2200, 257, 8, 3, "THT", 1.2, 1e-3),
    (150, 90, 11, 4, "FHT", 1.5, 1e-2),    # d > 8: runtime-dimension kernels
]


def _tubes(T, d, rng):
    lo = -1.0 + 0.05 * rng.random((T, d))
    hi = 1.0 - 0.05 * rng.random((T, d))
    ct = [Box(lo[t].astype(np.float32), hi[t].astype(np.float32)) for t in range(T)]
    tt = [Box((0.5 * lo[t]).astype(np.float32), (0.5 * hi[t]).astype(np.float32)) for t in range(T)]
    return ct, tt


@pytest.mark.parametrize("n,m,d,T,problem,sigma,lam", CASES)
def test_kernel_sr_random_shapes(n, m, d, T, problem, sigma, lam):
    import torch

    assert torch.cuda.is_available()
    from socks_b200.algorithms.reach.kernel_sr import KernelSR
    from socks_b200.kernel.metrics import rbf_kernel

    rng = np.random.default_rng(1000 * n + m)
    X = rng.uniform(-1.1, 1.1, (n, d))
    Y = X + 0.1 * rng.standard_normal((n, d))
    U = rng.uniform(-1, 1, (n, 1))
    pts = rng.uniform(-1.05, 1.05, (m, d))
    pts[0] = 1.0 - 0.05 * 0  # a corner of the nominal box
    ct, tt = _tubes(T, d, rng)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem=problem,
                   kernel_fn=partial(rbf_kernel, sigma=sigma), regularization_param=lam)
    alg.fit(list(zip(X, U, Y)))
    o = orc.KernelSROracle(T, ct, tt, problem, sigma, lam).fit(X, U, Y)
    assert np.max(np.abs(alg.value_functions - o.value_functions)) <= TOL
    got, want = alg.predict(pts), o.predict(pts)
    assert got.shape == (T, m)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.max(np.abs(got[ok] - want[ok]), initial=0.0) <= TOL
    alg.predict_variant = 1  # FMA-pipe contraction with CUDA's exp: an independent second implementation
    got1 = alg.predict(pts)
    assert np.max(np.abs(got1[ok] - want[ok]), initial=0.0) <= TOL


@pytest.mark.parametrize("n,m,P,T,problem", [(130, 140, 1, 3, "THT"), (300, 129, 33, 4, "FHT"), (513, 700, 128, 3, "FHT"),
                                             (257, 64, 129, 2, "THT")])
def test_kernel_sr_max_random_shapes(n, m, P, T, problem):
    from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR
    from socks_b200.kernel.metrics import rbf_kernel

    rng = np.random.default_rng(7 * n + P)
    d = 2
    X = rng.uniform(-1.1, 1.1, (n, d))
    U = rng.uniform(-1, 1, (n, 1))
    Y = X + 0.1 * U + 0.02 * rng.standard_normal((n, d))
    A = np.linspace(-1, 1, P).reshape(-1, 1)
    pts = rng.uniform(-1.05, 1.05, (m, d))
    ct, tt = _tubes(T, d, rng)
    alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem=problem,
                          regularization_param=1e-1, kernel_fn=partial(rbf_kernel, sigma=0.3))
    alg.fit(list(zip(X, U, Y)), A)
    o = orc.KernelMaximalSROracle(T, ct, tt, problem, 0.3, 1e-1).fit(X, U, Y, A)
    assert np.max(np.abs(alg.value_functions - o.value_functions)) <= TOL
    assert np.max(np.abs(alg.predict(pts) - o.predict(pts))) <= TOL
