"""KernelLinearId on the GPU: the reference's own test (identification/tests/test_kernel_linear_id.py:33-77, atol 1e-2
against the true CWH matrices) and a golden made by executing the unmodified reference (oracle/make_golden_linid.py)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cwh4d", "cwh6d"])
def test_linear_identification_cwh(name):
    from socks_b200.algorithms.identification.kernel_linear_id import KernelLinearId, kernel_linear_id

    g = load_golden("linear_id_cwh.npz")
    X, U, Y = g[f"{name}_X"], g[f"{name}_U"], g[f"{name}_Y"]
    alg = KernelLinearId(regularization_param=float(g["lam"]), verbose=False)
    alg.fit(S=list(zip(X, U, Y)))
    # the reference's assertions
    assert np.allclose(alg.state_matrix, g[f"{name}_A_true"], atol=1e-2)
    assert np.allclose(alg.input_matrix, g[f"{name}_B_true"], atol=1e-2)
    # parity with the reference's estimate (cond(C_ZZ) ~ 1.3: both solvers are accurate to a few ulp of the norm)
    tol = 1e-11
    for key, got in (("state_matrix", alg.state_matrix), ("input_matrix", alg.input_matrix),
                     ("disturbance_mean", alg.disturbance_mean)):
        want = g[f"{name}_{key}"]
        assert got.shape == want.shape and np.max(np.abs(got - want)) <= tol * max(1.0, np.max(np.abs(want))), key
    assert np.max(np.abs(alg.predict(g[f"{name}_T"]) - g[f"{name}_pred_T"])) <= 1e-10
    assert np.max(np.abs(alg.predict(g[f"{name}_T"], g[f"{name}_Ut"]) - g[f"{name}_pred_TU"])) <= 1e-10
    # shapes for which the reference's `result += np.squeeze(disturbance_mean.T)` raises, raise here as well
    with pytest.raises(ValueError):
        alg.predict(np.zeros((alg.state_matrix.shape[0] + 3, alg.state_matrix.shape[0])))
    assert np.array_equal(kernel_linear_id(list(zip(X, U, Y)), regularization_param=float(g["lam"])).state_matrix,
                          alg.state_matrix)


def test_linear_identification_large_sample_and_errors():
    """N = 40 000 (BASELINE.json configs[3] names kernel_linear_id at that size): estimates against numpy's solve of the
    same normal equations; default lambda = 1 (kernel_linear_id.py:99-100); None sample raises ValueError."""
    from socks_b200.algorithms.identification.kernel_linear_id import KernelLinearId

    rng = np.random.default_rng(3)
    n, d, m = 40000, 4, 2
    A, B, c = rng.standard_normal((d, d)), rng.standard_normal((d, m)), rng.standard_normal(d)
    X, U = rng.standard_normal((n, d)), rng.standard_normal((n, m))
    Y = X @ A.T + U @ B.T + c + 1e-3 * rng.standard_normal((n, d))
    alg = KernelLinearId(regularization_param=1e-9).fit_arrays(X, U, Y)
    Z = np.concatenate((X, U, np.ones((n, 1))), axis=1)
    H = np.linalg.solve((Z.T @ Z + 1e-9 * n * np.identity(d + m + 1)).T, (Y.T @ Z).T).T
    assert np.max(np.abs(alg.state_matrix - H[:d, :d])) <= 1e-11
    assert np.max(np.abs(alg.input_matrix - H[:d, d:d + m])) <= 1e-11
    assert np.max(np.abs(alg.disturbance_mean - H[:d, -1:])) <= 1e-11
    assert np.max(np.abs(alg.state_matrix - A)) < 1e-3
    assert KernelLinearId().fit_arrays(X[:100], U[:100], Y[:100]).regularization_param == 1
    with pytest.raises(ValueError):
        KernelLinearId().fit(None)
