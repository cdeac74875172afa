"""Seeded synthetic samples for the benchmark configurations (host-side numpy, float64).

The reference draws its samples by stepping gym environments
(`benchmarks/kernel_sr/baseline.py:36-72`, `examples/control/tracking.py:55-88`); gym is not
part of this build, so the same distributions are written in closed form
(SURVEY.md §8(d)).  All arrays are float64 so that the reference (when it is used to make
golden vectors) and this package see identical inputs.

Nothing here is on the timed path: the reference excludes sample generation from its timed
region too (`docs/benchmarks/index.rst:91-94`).
"""

import numpy as np

from socks_b200.spaces import Box

__all__ = [
    "integrator_sample",
    "uniform_points",
    "grid_points",
    "integrator_tubes",
    "unicycle_sample",
    "unicycle_actions",
    "tracking_cost_fn",
]


def integrator_sample(n, dim=2, seed=0, sampling_time=0.1, noise=1e-2):
    """(X, U, Y) for the `dim`-D chain of integrators driven by a scalar action.

    X ~ U[-1.1, 1.1]^dim (`benchmarks/kernel_sr/baseline.py:42-47`), U ~ U[-1, 1]
    (`gym_socks/envs/integrator.py:133-137`), disturbance 1e-2 * N(0, I)
    (`integrator.py:143-145`), one exact step of x' = A x + B u + w over `sampling_time`
    (the closed form of what `solve_ivp` integrates at `integrator.py:147-149`).
    """
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1.1, 1.1, size=(n, dim))
    U = rng.uniform(-1.0, 1.0, size=(n, 1))
    W = noise * rng.standard_normal(size=(n, dim))
    ts = float(sampling_time)
    # xdot_i = x_{i+1} + w_i (i < dim-1), xdot_{dim-1} = u + w_{dim-1}; exact Taylor step.
    Y = np.empty_like(X)
    drive = np.concatenate([X, U], axis=1)  # x_1..x_dim, u
    for i in range(dim):
        acc = X[:, i].copy()
        fact = 1.0
        for k in range(1, dim - i + 1):
            fact *= ts / k
            # k-th derivative of x_i is x_{i+k} plus the disturbance entering at level i+k-1
            acc += fact * (drive[:, i + k] + W[:, i + k - 1])
        Y[:, i] = acc
    return X, U, Y


def uniform_points(m, dim=2, seed=1, low=-1.0, high=1.0):
    """Test points ~ U[low, high]^dim (`benchmarks/kernel_sr/baseline.py:60-72`)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(low, high, size=(m, dim))


def grid_points(per_axis=50, dim=2, low=-1.0, high=1.0):
    """Regular grid, rows in the order of `make_grid_from_ranges`
    (`gym_socks/utils/grid.py:9-34`; used at `examples/reach/stoch_reach.py:92-94`).
    Includes points exactly on the constraint-box boundary."""
    axes = [np.linspace(low, high, per_axis)] * dim
    mesh = np.meshgrid(*axes, indexing="ij")
    return np.stack([m.reshape(-1) for m in mesh], axis=1)


def integrator_tubes(time_horizon, dim=2, constraint=1.0, target=0.5):
    """Constant constraint/target tubes (`benchmarks/kernel_sr/baseline.py:83-110`)."""
    c = [Box(low=-constraint, high=constraint, shape=(dim,), dtype=np.float32)] * time_horizon
    t = [Box(low=-target, high=target, shape=(dim,), dtype=np.float32)] * time_horizon
    return c, t


def unicycle_sample(n, seed=12345, sampling_time=0.1, noise=1e-2):
    """(X, U, Y) for the nonholonomic vehicle of `examples/control/tracking.py:61-88`,
    dynamics `gym_socks/envs/nonholonomic.py:84-94`; exact step for constant inputs."""
    rng = np.random.default_rng(seed)
    X = rng.uniform([-1.2, -1.2, -2 * np.pi], [1.2, 1.2, 2 * np.pi], size=(n, 3))
    U = rng.uniform([0.1, -10.1], [1.1, 10.1], size=(n, 2))
    W = noise * rng.standard_normal(size=(n, 3))
    ts = float(sampling_time)
    th0 = X[:, 2]
    om = U[:, 1] + W[:, 2]
    th1 = th0 + ts * om
    v = U[:, 0]
    safe = np.where(np.abs(om) < 1e-9, 1.0, om)
    dx = np.where(np.abs(om) < 1e-9, v * np.sin(th0) * ts, v * (np.cos(th0) - np.cos(th1)) / safe)
    dy = np.where(np.abs(om) < 1e-9, v * np.cos(th0) * ts, v * (np.sin(th1) - np.sin(th0)) / safe)
    Y = np.stack([X[:, 0] + dx + ts * W[:, 0], X[:, 1] + dy + ts * W[:, 1], th1], axis=1)
    return X, U, Y


def unicycle_actions(n_speed=25, n_turn=40):
    """Admissible-action grid over the action box (`examples/control/tracking.py:90`)."""
    a, b = np.meshgrid(np.linspace(0.1, 1.1, n_speed), np.linspace(-10.1, 10.1, n_turn), indexing="ij")
    return np.stack([a.reshape(-1), b.reshape(-1)], axis=1)


def tracking_cost_fn(time_horizon=20, amplitude=0.5, period=2.0):
    """Squared distance to the v-shaped path (`examples/control/tracking.py:94-123`)."""
    a, p = amplitude, period
    traj = [
        [
            (x * 0.1) - 1.0,
            4 * a / p * np.abs((((((x * 0.1) - 1.0) - p / 2) % p) + p) % p - p / 2) - a,
        ]
        for x in range(time_horizon)
    ]

    def _cost(time=0, state=None):
        dist = state[:, :2] - np.array([traj[time]])
        return np.power(np.linalg.norm(dist, ord=2, axis=1), 2)

    return _cost


def as_sample(X, U, Y):
    """List-of-tuples form the reference API takes (`sampling/transform.py:4-26`)."""
    return list(zip(X, U, Y))
