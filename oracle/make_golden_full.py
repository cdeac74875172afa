"""Full-size goldens for BASELINE.json configs[1] (KernelMaximalSR), [2], [3], [4] (test infrastructure).

Run in the authoring container only (needs /root/reference, ~45 GB RAM, ~45 min on 8 cores):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_full.py [c3] [srmax] [c4] [c5]

  c3    configs[2]: the UNMODIFIED REFERENCE `KernelControlFwd` (examples/control/tracking.py:41-156 regime:
        N=5000, P=25x40=1000, sigma=3, lambda=1e-7, T=20, seed 12345), scores + actions for 8 (time, state)
        pairs, cond(G), and an extended-precision truth of the same scores (LU + iterative refinement with
        np.longdouble residuals) so the GPU test can be condition-aware in the way SURVEY.md §8(c) asks.
  srmax configs[1] read as KernelMaximalSR: N=20000, P=100, T=15.  The reference cannot hold the N x N x P tensor of
        kernel_sr_max.py:45-47 (32 TB), so this fixture comes from the ORACLE PORT (oracle/socks_oracle.py, pinned
        against the reference on the smaller srmax_* goldens): kind = "port".
  c4    configs[3]: N=40000 2-D integrator, sigma=0.1, lambda=1/N^2: 64 columns of W = inv(G) from
        numpy.linalg.solve on the oracle Gram (rows subsampled every 16th to keep the fixture small).
  c5    configs[4]: N=40000, d=4, T=15 KernelSR, oracle port (numpy LU inverse), first 1000 of the 4M test points.

Inputs come from socks_b200.synthetic with fixed seeds and are NOT stored for the N >= 20000 cases (regenerated in
the test from the same seeds; the generators are covered by a checksum stored here).
"""
import hashlib
import os
import sys
import time
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

from oracle import socks_oracle as orc  # noqa: E402
from socks_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def checksum(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()[:16]


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name), **kw)
    print("wrote", name, {k: np.shape(v) for k, v in kw.items()}, flush=True)


def refine_solve(G, lu_piv, b, iters=4):
    """Solve G z = b to ~longdouble accuracy: double LU + residuals in np.longdouble."""
    from scipy.linalg import lu_solve

    Gl = None
    z = lu_solve(lu_piv, b).astype(np.longdouble)
    for _ in range(iters):
        if Gl is None:
            Gl = G.astype(np.longdouble)
        r = b.astype(np.longdouble) - Gl @ z
        z = z + lu_solve(lu_piv, r.astype(np.float64)).astype(np.longdouble)
    return z


def case_c3():
    import gym  # noqa: F401  (the shim)
    from gym_socks.algorithms.control.kernel_control_fwd import KernelControlFwd
    from gym_socks.kernel.metrics import rbf_kernel
    from scipy.linalg import lu_factor

    n, sigma, lam, T, seed = 5000, 3.0, 1e-7, 20, 12345
    X, U, Y = syn.unicycle_sample(n, seed=seed)
    A = syn.unicycle_actions(25, 40)
    cost = syn.tracking_cost_fn(T)
    pol = KernelControlFwd(cost_fn=cost, heuristic=True, regularization_param=lam,
                           kernel_fn=partial(rbf_kernel, sigma=sigma), verbose=False, time_horizon=T)
    t0 = time.time()
    pol.train(list(zip(X, U, Y)), A)
    print(f"c3: reference train {time.time() - t0:.1f} s", flush=True)
    rng = np.random.default_rng(seed + 5)
    times = np.array([0, 3, 6, 9, 12, 15, 17, 19])
    states = rng.uniform([-1, -0.6, -1], [1, 0.6, 1], size=(len(times), 3))
    G = rbf_kernel(X, sigma=sigma) * rbf_kernel(U, sigma=sigma)
    G[np.diag_indices(n)] += lam * n
    cond = float(np.linalg.cond(G))
    print(f"c3: cond(G) = {cond:.3e}", flush=True)
    lu_piv = lu_factor(G)
    Cs, Ct, acts, gaps = [], [], [], []
    for t, s in zip(times, states):
        s = s.reshape(1, -1)
        CXT = rbf_kernel(pol.X, s, sigma=sigma)
        beta = pol.W @ (CXT * pol.CUA)  # kernel_control_fwd.py:222
        c32 = np.array(pol.cost_fn(time=int(t)), dtype=np.float32)
        C = np.matmul(c32, beta)  # :225-226
        Cs.append(C)
        z = refine_solve(G, lu_piv, c32.astype(np.float64))
        q = z * CXT[:, 0].astype(np.longdouble)
        Ct.append(np.asarray(q @ pol.CUA.astype(np.longdouble), dtype=np.float64))
        acts.append(pol(time=int(t), state=s))
        srt = np.sort(C)
        gaps.append((srt[1] - srt[0]) / max(abs(srt[0]), 1e-300))
        print(f"c3: t={t} ref-vs-truth rel {np.max(np.abs(C - Ct[-1])) / np.max(np.abs(Ct[-1])):.2e} "
              f"argmin gap {gaps[-1]:.2e}", flush=True)
    save("fwd_c3_n5000_p1000.npz", n=n, sigma=sigma, lam=lam, T=T, seed=seed, n_speed=25, n_turn=40, times=times,
         states=states, C=np.array(Cs), C_truth=np.array(Ct), actions=np.array(acts), argmin_rel_gap=np.array(gaps),
         cond=cond, inputs_sha=checksum(X, U, Y, A), kind="reference")


def case_srmax():
    n, P, T, sigma, lam, m_sub = 20000, 100, 15, 0.1, 1.0, 512
    X, U, Y = syn.integrator_sample(n, dim=2, seed=0)
    A = np.linspace(-1, 1, P).reshape(-1, 1)
    pts = syn.uniform_points(250000, 2, seed=1)[:m_sub]
    ct, tt = syn.integrator_tubes(T, 2)
    t0 = time.time()
    alg = orc.KernelMaximalSROracle(T, ct, tt, "FHT", sigma, lam, chunk=512).fit(X, U, Y, A)
    t1 = time.time()
    out = alg.predict(pts)
    print(f"srmax: oracle fit {t1 - t0:.1f} s predict {time.time() - t1:.1f} s", flush=True)
    save("srmax_c2_n20000_p100_fht.npz", n=n, P=P, T=T, sigma=sigma, lam=lam, m_sub=m_sub, vf_stride=10,
         w_diag_sub=alg.w_diag[::10].copy(), value_functions_sub=alg.value_functions[:, ::10].copy(), out=out,
         inputs_sha=checksum(X, U, Y, A, pts), kind="port")


def case_c4():
    n, sigma = 40000, 0.1
    lam = 1 / n ** 2
    X, _, _ = syn.integrator_sample(n, dim=2, seed=0)
    t0 = time.time()
    G = orc.rbf_kernel(X, X, sigma)
    G[np.diag_indices(n)] += lam * n
    cols = np.arange(0, n, n // 64)[:64]
    rhs = np.zeros((n, len(cols)))
    rhs[cols, np.arange(len(cols))] = 1.0
    Wc = np.linalg.solve(G, rhs)
    print(f"c4: gram + solve {time.time() - t0:.1f} s", flush=True)
    # cond estimate: lambda_min >= lam*n, lambda_max <= max row sum (Gershgorin)
    lam_max = float(np.max(np.sum(G, axis=1)))
    res = float(np.max(np.abs(G @ Wc - rhs)))
    save("inv_c4_n40000.npz", n=n, sigma=sigma, lam=lam, cols=cols, row_stride=16, W_cols_sub=Wc[::16].copy(),
         W_diag_cols=Wc[cols, np.arange(len(cols))].copy(), cond_upper=lam_max / (lam * n), solve_residual=res,
         inputs_sha=checksum(X), kind="port")


def case_c5():
    n, d, T, sigma, m = 40000, 4, 15, 0.1, 1000
    lam = 1 / n ** 2
    X, U, Y = syn.integrator_sample(n, dim=d, seed=0)
    pts = syn.uniform_points(4000000, d, seed=1)[:m]
    ct, tt = syn.integrator_tubes(T, d)
    t0 = time.time()
    W = orc.regularized_inverse(X, regularization_param=lam, sigma=sigma)
    w = np.diag(W).copy()
    del W
    print(f"c5: oracle inverse {time.time() - t0:.1f} s", flush=True)
    alg = orc.KernelSROracle(T, ct, tt, "FHT", sigma, lam, chunk=500)
    alg.X, alg.w_diag = X, w
    CXY = orc.rbf_kernel(X, Y, sigma)
    # sr_recursion's beta = normalize(w[:, None] * C), in place to fit host RAM (same arithmetic, kernel_sr.py:45)
    CXY *= w[:, None]
    CXY /= np.sum(CXY, axis=0)
    vf = np.empty((T, n))
    vf[T - 1] = orc.indicator(Y, tt[T - 1])
    for t in range(T - 2, -1, -1):
        vf[t] = orc.fht_step(Y, vf[t + 1] @ CXY, ct[t], tt[t])
    del CXY
    alg.value_functions = vf
    out = alg.predict(pts)
    print(f"c5: total {time.time() - t0:.1f} s", flush=True)
    save("sr_c5_n40000_d4_fht.npz", n=n, d=d, T=T, sigma=sigma, lam=lam, m_sub=m, vf_stride=20,
         w_diag_sub=w[::20].copy(), value_functions_sub=vf[:, ::20].copy(), out=out,
         inputs_sha=checksum(X, U, Y, pts), kind="port")


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "srmax", "c4", "c5"]
    for name in which:
        {"c3": case_c3, "srmax": case_srmax, "c4": case_c4, "c5": case_c5}[name]()
