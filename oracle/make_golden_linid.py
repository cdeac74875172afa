"""Golden vectors for KernelLinearId, made by EXECUTING THE UNMODIFIED REFERENCE (test infrastructure).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_linid.py

Mirrors gym_socks/algorithms/identification/tests/test_kernel_linear_id.py:33-77 (CWH4D / CWH6D, 1000 samples,
lambda = 1e-9, tolerance 1e-2 against the true matrices): the true state / input matrices come from the reference's own
CWH environments (gym_socks/envs/cwh.py:116-207, :264-376); states and actions are drawn with numpy (gym's Box.sample of
an unbounded box is a standard normal; gym itself is not installed), next states by the reference's closed-form
dynamics A x + B u + w with its disturbance scale (cwh.py:199-202)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "gym_shim"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import gym  # noqa: E402,F401  (the shim)
from gym_socks.algorithms.identification.kernel_linear_id import KernelLinearId  # noqa: E402
from gym_socks.envs.cwh import CWH4DEnv, CWH6DEnv  # noqa: E402

out = {}
for name, Env, wscale in (("cwh4d", CWH4DEnv, [1e-4, 1e-4, 5e-8, 5e-8]),
                          ("cwh6d", CWH6DEnv, [1e-4, 1e-4, 1e-4, 5e-8, 5e-8, 5e-8])):
    env = Env()
    A, B = np.array(env.state_matrix, dtype=np.float64), np.array(env.input_matrix, dtype=np.float64)
    rng = np.random.default_rng(0)
    n = 1000
    X = rng.standard_normal((n, A.shape[0]))
    U = rng.standard_normal((n, B.shape[1]))
    W = rng.standard_normal((n, A.shape[0])) * np.array(wscale)
    Y = X @ A.T + U @ B.T + W
    alg = KernelLinearId(regularization_param=1e-9, verbose=False)
    alg.fit(list(zip(X, U, Y)))
    d = A.shape[0]
    T = rng.standard_normal((d, d))      # len(T) == x_dim: the only shape the reference's predict broadcasts for
    Ut = rng.standard_normal((d, B.shape[1]))
    Z = np.concatenate((X, U, np.ones((n, 1))), axis=1)
    out.update({f"{name}_A_true": A, f"{name}_B_true": B, f"{name}_X": X, f"{name}_U": U, f"{name}_Y": Y,
                f"{name}_state_matrix": alg.state_matrix, f"{name}_input_matrix": alg.input_matrix,
                f"{name}_disturbance_mean": alg.disturbance_mean, f"{name}_T": T, f"{name}_Ut": Ut,
                f"{name}_pred_T": alg.predict(T), f"{name}_pred_TU": alg.predict(T, Ut),
                f"{name}_cond": np.linalg.cond(Z.T @ Z + 1e-9 * n * np.identity(Z.shape[1]))})
    print(name, "max |A - A_true|", np.max(np.abs(alg.state_matrix - A)), "cond", out[f"{name}_cond"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "linear_id_cwh.npz"), lam=1e-9, **out)
print("wrote linear_id_cwh.npz")
