"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (test infrastructure).

Run in the authoring container only (needs /root/reference; the GPU box does not have it):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference is imported from /root/reference with the test-only gym shim in
oracle/gym_shim (gym is not installed).  Inputs are float64 and come from
`socks_b200.synthetic` with fixed seeds; inputs AND reference outputs are stored so the
fixtures stay valid even if the generators change.
"""
import os
import sys
import warnings
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "gym_shim"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import gym  # noqa: E402  (the shim)
from gym_socks.kernel.metrics import rbf_kernel, regularized_inverse  # noqa: E402
from gym_socks.algorithms.reach.kernel_sr import KernelSR  # noqa: E402
from gym_socks.algorithms.reach.kernel_sr_max import KernelMaximalSR  # noqa: E402
from gym_socks.algorithms.control.kernel_control_fwd import KernelControlFwd  # noqa: E402
from gym_socks.algorithms.control.kernel_control_bwd import KernelControlBwd  # noqa: E402
from gym_socks.utils import normalize, indicator_fn  # noqa: E402

from socks_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def box(lo, hi, d):
    return gym.spaces.Box(low=np.asarray(lo, dtype=np.float32), high=np.asarray(hi, dtype=np.float32),
                          shape=(d,), dtype=np.float32)


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name), **kw)
    print("wrote", name, {k: np.shape(v) for k, v in kw.items()})


def kernel_cases():
    rng = np.random.default_rng(7)
    X = rng.standard_normal((37, 3))
    Y = rng.standard_normal((23, 3))
    U = rng.standard_normal((37, 2))
    X2 = np.arange(4).reshape(2, 2)
    X3 = np.arange(6).reshape(3, 2)
    save("kernel_metrics.npz",
         X=X, Y=Y, U=U,
         K_xy_s07=rbf_kernel(X, Y, sigma=0.7),
         K_xx_s01=rbf_kernel(X, sigma=0.1),
         K_xx_median=rbf_kernel(X),
         K_arange22_median=rbf_kernel(X2),
         K_arange32_median=rbf_kernel(X3),
         K_arange22_32_median=rbf_kernel(X2, X3),
         W_arange22_default=regularized_inverse(X2),
         W_x_default=regularized_inverse(X),
         W_x_s07_lam=regularized_inverse(X, kernel_fn=partial(rbf_kernel, sigma=0.7), regularization_param=1e-3),
         W_xu_s07_lam=regularized_inverse(X, U=U, kernel_fn=partial(rbf_kernel, sigma=0.7), regularization_param=1e-3),
         normalize_in=np.abs(X[:6, :]) + 0.1,
         normalize_out=normalize(np.abs(X[:6, :]) + 0.1),
         indicator_pts=np.array([[0.5, -0.5], [1.0, -1.0], [1.0000001, 0.0], [0.0, -1.5], [-1.0, 1.0]]),
         indicator_out=indicator_fn(np.array([[0.5, -0.5], [1.0, -1.0], [1.0000001, 0.0], [0.0, -1.5], [-1.0, 1.0]]),
                                    box(-1, 1, 2)))


def sr_case(name, n, m_pts, T, d, sigma, lam, problem, seed, grid=None, time_varying=False):
    X, U, Y = syn.integrator_sample(n, dim=d, seed=seed)
    pts = syn.grid_points(grid, d) if grid else syn.uniform_points(m_pts, d, seed=seed + 1, low=-1.05, high=1.05)
    if time_varying:
        lo = [(-1.0 + 0.02 * t) for t in range(T)]
        ctube = [box([l] * d, [-l] * d, d) for l in lo]
        ttube = [box([-0.5 + 0.01 * t] * d, [0.5 - 0.01 * t] * d, d) for t in range(T)]
    else:
        ctube = [box([-1] * d, [1] * d, d)] * T
        ttube = [box([-0.5] * d, [0.5] * d, d)] * T
    alg = KernelSR(time_horizon=T, constraint_tube=ctube, target_tube=ttube, problem=problem,
                   kernel_fn=partial(rbf_kernel, sigma=sigma), regularization_param=lam)
    alg.fit(list(zip(X, U, Y)))
    out = alg.predict(pts)
    save(name, X=X, U=U, Y=Y, pts=pts, T=T, sigma=sigma, lam=lam, problem=problem,
         ctube=np.array([[b.low, b.high] for b in ctube], dtype=np.float64),
         ttube=np.array([[b.low, b.high] for b in ttube], dtype=np.float64),
         w_diag=np.diag(alg.W).copy(), value_functions=alg.value_functions, out=out)


def srmax_case(name, n, m_pts, p, T, sigma, lam, problem, seed, batch_size=None):
    d = 2
    X, U, Y = syn.integrator_sample(n, dim=d, seed=seed)
    pts = syn.uniform_points(m_pts, d, seed=seed + 1, low=-1.05, high=1.05)
    A = np.linspace(-1, 1, p).reshape(-1, 1)
    ctube = [box([-1] * d, [1] * d, d)] * T
    ttube = [box([-0.5] * d, [0.5] * d, d)] * T
    alg = KernelMaximalSR(time_horizon=T, constraint_tube=ctube, target_tube=ttube, problem=problem,
                          kernel_fn=partial(rbf_kernel, sigma=sigma), regularization_param=lam,
                          batch_size=batch_size)
    alg.fit(list(zip(X, U, Y)), A)
    out = alg.predict(pts)
    save(name, X=X, U=U, Y=Y, A=A, pts=pts, T=T, sigma=sigma, lam=lam, problem=problem,
         ctube=np.array([[b.low, b.high] for b in ctube], dtype=np.float64),
         ttube=np.array([[b.low, b.high] for b in ttube], dtype=np.float64),
         w_diag=np.diag(alg.W).copy(), CUA=alg.CUA, value_functions=alg.value_functions, out=out)


def fwd_case(name, n, sigma, lam, T, seed, constrained):
    X, U, Y = syn.unicycle_sample(n, seed=seed)
    A = syn.unicycle_actions(6, 9)
    cost = syn.tracking_cost_fn(T)

    def constraint(time=0, state=None):  # keep |y| below a moving bound (<= 0 is feasible)
        return np.abs(state[:, 1]) - (0.45 + 0.01 * time)

    pol = KernelControlFwd(cost_fn=cost, constraint_fn=constraint if constrained else None, heuristic=True,
                           regularization_param=lam, kernel_fn=partial(rbf_kernel, sigma=sigma), verbose=False)
    pol.train(list(zip(X, U, Y)), A)
    rng = np.random.default_rng(seed + 5)
    states = rng.uniform([-1, -0.6, -1], [1, 0.6, 1], size=(T, 3))
    Cs, Ds, acts = [], [], []
    for t in range(T):
        s = states[t:t + 1]
        CXT = rbf_kernel(pol.X, s, sigma=sigma)
        beta = pol.W @ (CXT * pol.CUA)
        Cs.append(np.matmul(np.array(pol.cost_fn(time=t), dtype=np.float32), beta))
        if constrained:
            Ds.append(np.matmul(np.array(pol.constraint_fn(time=t), dtype=np.float32), beta))
        acts.append(pol(time=t, state=s))
    save(name, X=X, U=U, Y=Y, A=A, T=T, sigma=sigma, lam=lam, states=states, constrained=constrained,
         C=np.array(Cs), D=np.array(Ds) if constrained else np.zeros((0,)), actions=np.array(acts),
         cond=np.linalg.cond(np.linalg.inv(pol.W)))


def bwd_case(name, n, sigma, lam, T, seed, constrained):
    X, U, Y = syn.unicycle_sample(n, seed=seed)
    A = syn.unicycle_actions(6, 9)
    cost = syn.tracking_cost_fn(T)

    def constraint(time=0, state=None):
        return np.abs(state[:, 1]) - (0.45 + 0.01 * time)

    pol = KernelControlBwd(time_horizon=T, cost_fn=cost, constraint_fn=constraint if constrained else None,
                           heuristic=True, regularization_param=lam, kernel_fn=partial(rbf_kernel, sigma=sigma),
                           verbose=False)
    pol.train(list(zip(X, U, Y)), A)
    rng = np.random.default_rng(seed + 5)
    states = rng.uniform([-1, -0.6, -1], [1, 0.6, 1], size=(T, 3))
    Cs, acts = [], []
    for t in range(T):
        s = states[t:t + 1]
        CXT = rbf_kernel(pol.X, s, sigma=sigma)
        beta = pol.W @ (CXT * pol.CUA)
        Cs.append(np.matmul(np.array(pol.value_functions[t], dtype=np.float32), beta))
        acts.append(pol(time=t, state=s))
    save(name, X=X, U=U, Y=Y, A=A, T=T, sigma=sigma, lam=lam, states=states, constrained=constrained,
         value_functions=pol.value_functions, C=np.array(Cs), actions=np.array(acts),
         cond=np.linalg.cond(np.linalg.inv(pol.W)))


def separating_case(name, n, m, sigma, lam, seed):
    from gym_socks.kernel.metrics import abel_kernel, euclidean_distance
    from gym_socks.algorithms.reach.separating_kernel import SeparatingKernelClassifier

    rng = np.random.default_rng(seed)
    ang = rng.uniform(0, 2 * np.pi, n)
    rad = 0.6 + 0.05 * rng.standard_normal(n)
    X = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1)  # a ring: the support to be learned
    T = rng.uniform(-1, 1, (m, 2))
    clf = SeparatingKernelClassifier(kernel_fn=partial(abel_kernel, sigma=sigma), regularization_param=lam)
    clf.fit(X)
    K = abel_kernel(clf.X, T, sigma=sigma)
    C = np.diagonal((1 / len(X)) * K.T @ clf.W @ K)
    save(name, X=X, T=T, sigma=sigma, lam=lam, tau=clf.tau, C=C, labels=clf.predict(T),
         K_abel=abel_kernel(X[:40], T[:30], sigma=sigma), K_abel_median=abel_kernel(X[:40]),
         D=euclidean_distance(X[:40], T[:30]), D2=euclidean_distance(X[:40], T[:30], squared=True),
         cond=np.linalg.cond(np.linalg.inv(clf.W)))


def embedding_case(name, seed):
    from gym_socks.algorithms.kernel import ConditionalEmbedding
    from gym_socks.kernel.probability import maximum_mean_discrepancy, witness_function

    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (150, 2))
    y1 = np.sin(3 * X[:, 0]) + X[:, 1] ** 2
    y2 = np.stack([y1, np.cos(2 * X[:, 1])], axis=1)
    Xt = rng.uniform(-1, 1, (70, 2))
    G = rbf_kernel(X, sigma=0.4)
    K = rbf_kernel(X, Xt, sigma=0.4)
    ce = ConditionalEmbedding(regularization_param=1e-4).fit(G)
    P = rng.normal(0.0, 1.0, (120, 2))
    Q = rng.normal(0.3, 1.2, (90, 2))
    t = rng.uniform(-2, 2, (40, 2))
    kf = partial(rbf_kernel, sigma=0.8)
    save(name, X=X, y1=y1, y2=y2, Xt=Xt, G=G, K=K, lam=1e-4, pred1=ce.predict(y1, K), pred2=ce.predict(y2, K),
         cond=np.linalg.cond(ce._W), P=P, Q=Q, t=t,
         mmd_unbiased=maximum_mean_discrepancy(P, Q, kernel_fn=kf), mmd_biased_sq=maximum_mean_discrepancy(P, Q, kernel_fn=kf, biased=True, squared=True),
         mmd_default=maximum_mean_discrepancy(P, Q, squared=True), witness=witness_function(P, Q, t, kernel_fn=kf))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mid":
        # mid-size cases in the regimes the reference's benchmarks/examples use
        srmax_case("srmax_n1200_p100_fht.npz", 1200, 400, 100, 5, 0.1, 1.0, "FHT", seed=8)   # kernel_sr_max defaults
        fwd_case("fwd_tracking_regime.npz", 1500, 3.0, 1e-7, 6, seed=12345, constrained=False)  # tracking.py regime
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "emb":
        embedding_case("embedding_mmd.npz", seed=41)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sep":
        separating_case("separating_ring.npz", 300, 400, 0.25, 1e-3, seed=31)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bwd":
        bwd_case("bwd_unconstrained.npz", 300, 1.0, 1e-3, 6, seed=21, constrained=False)
        bwd_case("bwd_constrained.npz", 260, 1.5, 1e-3, 5, seed=22, constrained=True)
        sys.exit(0)
    kernel_cases()
    # BASELINE.json configs[0]: N=1000, T=15, 50x50 grid, sigma=0.1, lambda=1/N^2
    sr_case("sr_c1_fht.npz", 1000, None, 15, 2, 0.1, 1 / 1000 ** 2, "FHT", seed=0, grid=50)
    sr_case("sr_c1_tht.npz", 1000, None, 15, 2, 0.1, 1 / 1000 ** 2, "THT", seed=0, grid=50)
    sr_case("sr_d3_timevarying_fht.npz", 300, 257, 6, 3, 0.25, 1.0, "FHT", seed=3, time_varying=True)
    sr_case("sr_d4_tht.npz", 260, 130, 4, 4, 0.3, 1e-4, "THT", seed=4)
    srmax_case("srmax_fht.npz", 260, 301, 7, 5, 0.1, 1.0, "FHT", seed=5)
    srmax_case("srmax_tht_batch.npz", 200, 150, 12, 4, 0.15, 1 / 200 ** 2, "THT", seed=6, batch_size=64)
    fwd_case("fwd_unconstrained.npz", 400, 3.0, 1e-5, 8, seed=12345, constrained=False)
    fwd_case("fwd_constrained.npz", 300, 1.0, 1e-3, 6, seed=77, constrained=True)
