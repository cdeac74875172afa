"""Pins the CPU oracle (oracle/socks_oracle.py) to the reference.

(1) the reference's own known-answer tests for this path
    (gym_socks/kernel/tests/test_kernel.py:46-130, gym_socks/utils/tests/test_utils.py:10-54);
(2) golden vectors produced by executing the unmodified reference (oracle/make_golden.py).
CPU only.
"""
import numpy as np
import pytest

from conftest import load_golden, tubes_from
from oracle import socks_oracle as orc


def test_rbf_known_answers():
    # test_kernel.py:46-66 (sigma=None -> median of squared distances)
    Y = np.arange(4).reshape((2, 2))
    gt = np.array([[1.0, 0.77880078], [0.77880078, 1.0]])
    assert np.allclose(orc.rbf_kernel(Y), gt)
    assert np.allclose(orc.rbf_kernel(Y, Y), gt)
    # test_kernel.py:68-100
    Y3 = np.arange(6).reshape((3, 2))
    gt3 = np.array([[1.0, 0.93941306, 0.77880078], [0.93941306, 1.0, 0.93941306], [0.77880078, 0.93941306, 1.0]])
    assert np.allclose(orc.rbf_kernel(Y3), gt3)


def test_regularized_inverse_known_answer():
    # test_kernel.py:118-130
    gt = np.array([[0.66676608, -0.0081415], [-0.0081415, 0.66676608]])
    assert np.allclose(orc.regularized_inverse(np.arange(4).reshape((2, 2))), gt)


def test_normalize_indicator_known_answers():
    # test_utils.py:10-29
    assert np.allclose(orc.normalize(np.array([1.0, 1.0])), [0.5, 0.5])
    assert np.allclose(orc.normalize(np.ones((2, 2))), 0.5 * np.ones((2, 2)))
    g = load_golden("kernel_metrics.npz")
    assert np.array_equal(orc.normalize(g["normalize_in"]), g["normalize_out"])
    from socks_b200.spaces import Box

    assert np.array_equal(orc.indicator(g["indicator_pts"], Box(-1, 1, (2,))), g["indicator_out"])


def test_kernel_golden():
    g = load_golden("kernel_metrics.npz")
    X, Y, U = g["X"], g["Y"], g["U"]
    assert np.array_equal(orc.rbf_kernel(X, Y, 0.7), g["K_xy_s07"])
    assert np.array_equal(orc.rbf_kernel(X, None, 0.1), g["K_xx_s01"])
    assert np.array_equal(orc.rbf_kernel(X), g["K_xx_median"])
    assert np.allclose(orc.regularized_inverse(X), g["W_x_default"], rtol=1e-12, atol=0)
    assert np.allclose(orc.regularized_inverse(X, regularization_param=1e-3, sigma=0.7), g["W_x_s07_lam"], rtol=1e-11)
    assert np.allclose(orc.regularized_inverse(X, U=U, regularization_param=1e-3, sigma=0.7), g["W_xu_s07_lam"], rtol=1e-11)


@pytest.mark.parametrize("name", ["sr_c1_fht.npz", "sr_c1_tht.npz", "sr_d3_timevarying_fht.npz", "sr_d4_tht.npz"])
def test_kernel_sr_golden(name):
    g = load_golden(name)
    alg = orc.KernelSROracle(int(g["T"]), tubes_from(g["ctube"]), tubes_from(g["ttube"]), str(g["problem"]),
                             float(g["sigma"]), float(g["lam"]))
    alg.fit(g["X"], g["U"], g["Y"])
    assert np.allclose(alg.w_diag, g["w_diag"], rtol=1e-9)
    assert np.max(np.abs(alg.value_functions - g["value_functions"])) < 1e-11
    assert np.max(np.abs(alg.predict(g["pts"]) - g["out"])) < 1e-11


@pytest.mark.parametrize("name", ["srmax_fht.npz", "srmax_tht_batch.npz", "srmax_n1200_p100_fht.npz"])
def test_kernel_sr_max_golden(name):
    g = load_golden(name)
    alg = orc.KernelMaximalSROracle(int(g["T"]), tubes_from(g["ctube"]), tubes_from(g["ttube"]),
                                    str(g["problem"]), float(g["sigma"]), float(g["lam"]))
    alg.fit(g["X"], g["U"], g["Y"], g["A"])
    assert np.array_equal(alg.CUA, g["CUA"])
    assert np.max(np.abs(alg.value_functions - g["value_functions"])) < 1e-11
    assert np.max(np.abs(alg.predict(g["pts"]) - g["out"])) < 1e-11


@pytest.mark.parametrize("name", ["fwd_unconstrained.npz", "fwd_constrained.npz", "fwd_tracking_regime.npz"])
def test_control_fwd_golden(name):
    from socks_b200 import synthetic as syn

    g = load_golden(name)
    T = int(g["T"])
    cons = None
    if bool(g["constrained"]):
        def cons(time=0, state=None):
            return np.abs(state[:, 1]) - (0.45 + 0.01 * time)
    pol = orc.KernelControlFwdOracle(syn.tracking_cost_fn(T), cons, float(g["sigma"]), float(g["lam"]))
    pol.train(g["X"], g["U"], g["Y"], g["A"])
    tol = max(1e-8, 100 * float(g["cond"]) * 2.0 ** -53)
    for t in range(T):
        C, D = pol.scores(t, g["states"][t])
        assert np.max(np.abs(C - g["C"][t])) <= tol * np.max(np.abs(g["C"][t]))
        if cons is not None:
            assert np.max(np.abs(D - g["D"][t])) <= tol * np.max(np.abs(g["D"][t]))
        assert np.array_equal(pol(time=t, state=g["states"][t]), g["actions"][t])


@pytest.mark.parametrize("name", ["bwd_unconstrained.npz", "bwd_constrained.npz"])
def test_control_bwd_golden(name):
    from socks_b200 import synthetic as syn

    g = load_golden(name)
    T = int(g["T"])
    cons = None
    if bool(g["constrained"]):
        def cons(time=0, state=None):
            return np.abs(state[:, 1]) - (0.45 + 0.01 * time)
    pol = orc.KernelControlBwdOracle(T, syn.tracking_cost_fn(T), cons, float(g["sigma"]), float(g["lam"]))
    pol.train(g["X"], g["U"], g["Y"], g["A"])
    tol = max(1e-8, 100 * float(g["cond"]) * 2.0 ** -53)
    scale = np.max(np.abs(g["value_functions"]))
    assert np.max(np.abs(pol.value_functions - g["value_functions"])) <= tol * scale
    for t in range(T):
        C, _ = pol.scores(t, g["states"][t])
        assert np.max(np.abs(C - g["C"][t])) <= tol * np.max(np.abs(g["C"][t]))
        assert np.array_equal(pol(time=t, state=g["states"][t]), g["actions"][t])


def test_abel_and_separating_golden():
    g = load_golden("separating_ring.npz")
    X, T = g["X"], g["T"]
    assert np.array_equal(orc.abel_kernel(X[:40], T[:30], float(g["sigma"])), g["K_abel"])
    assert np.array_equal(orc.abel_kernel(X[:40]), g["K_abel_median"])
    clf = orc.SeparatingKernelOracle(float(g["sigma"]), float(g["lam"])).fit(X)
    tol = 100 * float(g["cond"]) * 2.0 ** -53
    assert abs(clf.tau - float(g["tau"])) <= tol
    assert np.max(np.abs(clf.scores(T) - g["C"])) <= tol
    assert np.array_equal(clf.predict(T), g["labels"])


def test_embedding_and_mmd_golden():
    g = load_golden("embedding_mmd.npz")
    tol = 100 * float(g["cond"]) * 2.0 ** -53
    for y, want in ((g["y1"], g["pred1"]), (g["y2"], g["pred2"])):
        got = orc.conditional_embedding_predict(g["G"], float(g["lam"]), y, g["K"])
        assert got.shape == want.shape and np.max(np.abs(got - want)) <= tol * np.max(np.abs(want))
    assert abs(orc.maximum_mean_discrepancy(g["P"], g["Q"], 0.8) - float(g["mmd_unbiased"])) <= 1e-13
    assert abs(orc.maximum_mean_discrepancy(g["P"], g["Q"], 0.8, biased=True, squared=True) - float(g["mmd_biased_sq"])) <= 1e-14
    assert abs(orc.maximum_mean_discrepancy(g["P"], g["Q"], 1.0, squared=True) - float(g["mmd_default"])) <= 1e-14
    assert np.max(np.abs(orc.witness_function(g["P"], g["Q"], g["t"], 0.8) - g["witness"])) <= 1e-14
