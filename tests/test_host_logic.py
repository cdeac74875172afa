"""CPU-only tests of the host-side mirror of the reference interface (no kernels run)."""
from functools import partial

import numpy as np
import pytest

from conftest import load_golden
from socks_b200.algorithms.control.common import compute_solution
from socks_b200.algorithms.reach.kernel_sr import KernelSR
from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR
from socks_b200.algorithms.control.kernel_control_fwd import KernelControlFwd
from socks_b200.kernel.metrics import RBFSpec, rbf_kernel, regularized_inverse, resolve_kernel
from socks_b200.sampling.transform import transpose_sample
from socks_b200.spaces import Box
from socks_b200.utils import generate_batches, indicator_fn, normalize
from socks_b200 import synthetic as syn


def test_resolve_kernel_forms():
    assert resolve_kernel(None, 0.1).sigma == 0.1
    s = resolve_kernel(partial(rbf_kernel, sigma=0.5), 1)
    assert s.style == "sigma" and s.sigma == 0.5
    from sklearn.metrics.pairwise import rbf_kernel as sk_rbf

    g = resolve_kernel(partial(sk_rbf, gamma=50.0), 1)
    assert g.style == "gamma" and g.gamma == 50.0
    assert g.params_for(np.zeros((2, 2)), None) == (-50.0, 0)
    assert resolve_kernel(sk_rbf, 1).params_for(np.zeros((2, 4)), None) == (-0.25, 0)  # sklearn default 1/n_features
    assert RBFSpec(sigma=0.1).params_for(None, None) == (-2 * 0.1 ** 2, 1)  # the reference divides by -2 sigma^2
    assert RBFSpec(sigma=0.1).gamma_for(None, None) == pytest.approx(50.0)
    # the sklearn kernels the reference's own test feeds to regularized_inverse (test_kernel.py:18,132-147)
    from sklearn.metrics.pairwise import laplacian_kernel, linear_kernel, polynomial_kernel, sigmoid_kernel

    lap = resolve_kernel(partial(laplacian_kernel, gamma=2.0), 1)
    assert lap.kind == 4 and lap.params_for(np.zeros((2, 3)), None) == (-2.0, 0)
    assert resolve_kernel(laplacian_kernel, 1).params_for(np.zeros((2, 4)), None) == (-0.25, 0)
    pol = resolve_kernel(partial(polynomial_kernel, degree=2, coef0=0.5), 1)
    assert pol.kind == 3 and pol.extra == (0.5, 2.0) and pol.params_for(np.zeros((2, 5)), None) == (0.2, 0)
    assert resolve_kernel(linear_kernel, 1).kind == 2
    with pytest.raises(TypeError):
        resolve_kernel(sigmoid_kernel, 1)
    with pytest.raises(TypeError):
        resolve_kernel(partial(polynomial_kernel, bogus=1), 1)
    with pytest.raises(TypeError):
        resolve_kernel(lambda X, Y=None: X, 1)
    with pytest.raises(AssertionError):
        resolve_kernel(partial(rbf_kernel, sigma=-1.0), 1)


def test_assertions_match_reference():
    # metrics.py:133, :254-256, :261-262, :269-270 raise AssertionError before any device work
    with pytest.raises(AssertionError):
        rbf_kernel(np.zeros((2, 2)), sigma=0)
    with pytest.raises(AssertionError):
        regularized_inverse(np.zeros((3, 2)), regularization_param=0)
    with pytest.raises(AssertionError):
        regularized_inverse(np.zeros((3, 2)), Y=np.zeros((4, 2)))
    with pytest.raises(AssertionError):
        regularized_inverse(np.zeros((3, 2)), U=np.zeros((4, 1)))


def test_algorithm_validation_errors():
    tube = [Box(-1, 1, (2,))] * 3
    S = syn.as_sample(*syn.integrator_sample(8))
    with pytest.raises(ValueError):
        KernelSR(time_horizon=3, constraint_tube=tube, target_tube=tube).fit(None)
    with pytest.raises(ValueError):
        KernelSR(constraint_tube=tube, target_tube=tube).fit(S)
    with pytest.raises(AssertionError):
        KernelSR(time_horizon=3.0, constraint_tube=tube, target_tube=tube).fit(S)
    with pytest.raises(ValueError):
        KernelSR(time_horizon=3, target_tube=tube).fit(S)
    with pytest.raises(ValueError):
        KernelSR(time_horizon=3, constraint_tube=tube).fit(S)
    with pytest.raises(ValueError):
        KernelSR(time_horizon=3, constraint_tube=tube, target_tube=tube, problem="XHT").fit(S)
    with pytest.raises(ValueError):
        KernelMaximalSR(time_horizon=3, constraint_tube=tube, target_tube=tube, problem="XHT").fit(S, np.zeros((2, 1)))
    with pytest.raises(ValueError):
        KernelControlFwd().train(S, np.zeros((2, 1)))
    with pytest.raises(ValueError):
        KernelControlFwd(cost_fn=lambda time, state: 0).train(None, np.zeros((2, 1)))
    assert KernelControlFwd(cost_fn=lambda time, state: 0)(state=None) is None


def test_utils_known_answers():
    # gym_socks/utils/tests/test_utils.py:10-54 and tests/test_batch.py:10-41
    assert np.allclose(normalize(np.array([1.0, 1.0])), [0.5, 0.5])
    assert np.allclose(normalize(np.ones((2, 2))), 0.5)
    box = Box(-1, 1, (2,))
    assert indicator_fn(np.array([[0.5, 0.5]]), box)[0]
    assert not indicator_fn(np.array([[1.5, 0.5]]), box)[0]
    assert list(generate_batches(10, 5)) == [slice(0, 5), slice(5, 10)]
    assert list(generate_batches(10, 3)) == [slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 10)]
    assert list(generate_batches(3, 5)) == [slice(0, 3)]
    assert transpose_sample([(1, 2, 3), (4, 5, 6)]) == ([1, 4], [2, 5], [3, 6])


@pytest.mark.parametrize("name", ["fwd_unconstrained.npz", "fwd_constrained.npz"])
def test_compute_solution_matches_reference_actions(name):
    g = load_golden(name)
    for t in range(int(g["T"])):
        D = g["D"][t] if bool(g["constrained"]) else None
        for heuristic in (True, False):
            sol = compute_solution(g["C"][t], D, heuristic=heuristic)
            if heuristic:
                assert np.array_equal(np.array(g["A"][int(np.argmax(sol))], dtype=np.float32), g["actions"][t])
            else:
                assert abs(sol.sum() - 1) < 1e-8


def test_synthetic_is_seeded_and_closed_form():
    X, U, Y = syn.integrator_sample(5, seed=0)
    X2, U2, Y2 = syn.integrator_sample(5, seed=0)
    assert np.array_equal(Y, Y2)
    g = load_golden("sr_c1_fht.npz")
    Xg, Ug, Yg = syn.integrator_sample(1000, dim=2, seed=0)
    assert np.array_equal(Xg, g["X"]) and np.array_equal(Yg, g["Y"])
    assert np.array_equal(syn.grid_points(50, 2), g["pts"])


def test_host_pool_never_reuses_a_live_result(monkeypatch):
    """The pinned result pool of predict(): a buffer is reused only after every array/view handed out is gone."""
    import gc

    import torch

    from socks_b200 import _device as dv

    orig = torch.empty

    def unpinned_empty(*a, **k):  # no CUDA here: drop pin_memory
        k.pop("pin_memory", None)
        return orig(*a, **k)

    monkeypatch.setattr(torch, "empty", unpinned_empty)
    pool = dv.HostPool()
    t1, f1 = pool.get(3, 4)
    t1.fill_(1.0)
    a = f1()
    t2, f2 = pool.get(3, 4)
    assert t2 is not t1
    t2.fill_(2.0)
    b = f2()
    assert a[0, 0] == 1.0 and b[0, 0] == 2.0
    view = a[1:]
    del a
    gc.collect()
    t3, _ = pool.get(3, 4)
    assert t3 is not t1 and t3 is not t2  # the view still leases buffer 1
    del view
    gc.collect()
    t4, _ = pool.get(3, 4)
    assert t4 is t1


def test_sweep_repeat_rule_follows_the_reference():
    """tests/measure/sweeps.py repeats like benchmarks/experiment_matrix.py:44-62: one warm-up, >= 3 runs, <= 32 runs, and
    it stops as soon as `sample_mean` (:224-255) reports the last run outside the 95 % interval of the runs so far."""
    import importlib.util
    import os

    from conftest import ROOT

    spec = importlib.util.spec_from_file_location("sweeps", os.path.join(ROOT, "tests", "measure", "sweeps.py"))
    sw = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sw)
    calls = []

    def const():
        calls.append(1)
        return 1.0 + 1e-6 * (len(calls) % 2)  # always inside the interval: runs to max_iter

    assert len(sw.reference_repeats(const)) == 32 and len(calls) == 33  # + the discarded warm-up
    seq = iter([1.0, 1.02, 0.98, 1.01, 1.0, 0.99, 5.0, 1.0, 1.0])
    out = sw.reference_repeats(lambda: next(seq), warmup=False)
    assert out[-1] == 5.0 and len(out) == 7  # the outlier ends the repeats, as the reference codes it
    assert sw.sample_mean([1.0, 1.01, 0.99, 1.0]) is True and sw.sample_mean([1.0, 1.01, 0.99, 1.0, 1.0, 1.0, 3.0]) is False
    assert len(sw.reference_repeats(lambda: 1.0, min_iter=3, max_iter=5, warmup=False, callback=None)) == 3


def test_factorization_engine_argument_is_validated():
    from socks_b200.kernel import metrics as mt

    with pytest.raises(ValueError):
        mt.set_factorization_engine("fp32")  # before anything touches the device
