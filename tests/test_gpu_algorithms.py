"""GPU parity of KernelSR / KernelMaximalSR / KernelControlFwd through the public API (which calls the
C ABI) against the golden vectors made by the unmodified reference and against the oracle.
Tolerances (BASELINE.json north_star): safety probabilities abs <= 1e-8; policy scores
rel <= max(1e-8, 100 cond 2^-53); argmin/argmax exact."""
from functools import partial

import numpy as np
import pytest

from conftest import load_golden, tubes_from
from oracle import socks_oracle as orc
from socks_b200 import synthetic as syn
from socks_b200.spaces import Box

pytestmark = pytest.mark.gpu

TOL = 1e-8


def _api():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a B200"
    from socks_b200.algorithms.control.kernel_control_fwd import KernelControlFwd
    from socks_b200.algorithms.reach.kernel_sr import KernelSR, kernel_sr
    from socks_b200.algorithms.reach.kernel_sr_max import KernelMaximalSR, kernel_sr_max
    from socks_b200.kernel.metrics import rbf_kernel

    return KernelSR, kernel_sr, KernelMaximalSR, kernel_sr_max, KernelControlFwd, rbf_kernel


@pytest.mark.parametrize("name", ["sr_c1_fht.npz", "sr_c1_tht.npz", "sr_d3_timevarying_fht.npz", "sr_d4_tht.npz"])
@pytest.mark.parametrize("variant", [0, 1])
def test_kernel_sr_golden(name, variant):
    KernelSR, kernel_sr, _, _, _, rbf = _api()
    g = load_golden(name)
    S = syn.as_sample(g["X"], g["U"], g["Y"])
    alg = KernelSR(time_horizon=int(g["T"]), constraint_tube=tubes_from(g["ctube"]), target_tube=tubes_from(g["ttube"]),
                   problem=str(g["problem"]), kernel_fn=partial(rbf, sigma=float(g["sigma"])),
                   regularization_param=float(g["lam"]))
    alg.predict_variant = variant
    alg.fit(S)
    assert alg.value_functions.shape == g["value_functions"].shape
    assert np.max(np.abs(alg.value_functions - g["value_functions"])) <= TOL
    out = alg.predict(g["pts"])
    assert out.shape == g["out"].shape and out.dtype == np.float64
    assert np.max(np.abs(out - g["out"])) <= TOL
    assert np.array_equal(alg.X, g["X"])
    if variant == 0:
        w = np.diag(alg.W)
        assert np.max(np.abs(w - g["w_diag"]) / g["w_diag"]) < 1e-7
        # one-shot wrapper, batch_size accepted and result unchanged (kernel_sr.py:101-107)
        out2 = kernel_sr(S, g["pts"], time_horizon=int(g["T"]), constraint_tube=tubes_from(g["ctube"]),
                         target_tube=tubes_from(g["ttube"]), problem=str(g["problem"]),
                         regularization_param=float(g["lam"]), kernel_fn=partial(rbf, sigma=float(g["sigma"])),
                         batch_size=100)
        assert np.array_equal(out2, out)


def test_kernel_sr_median_heuristic_bandwidth():
    """kernel_fn=partial(rbf_kernel) with sigma=None: every kernel evaluation uses the median of ITS squared-distance
    matrix (metrics.py:130-131), including K(X, T) over the whole test set."""
    KernelSR, _, _, _, _, rbf = _api()
    X, U, Y = syn.integrator_sample(200, seed=51)
    pts = syn.uniform_points(150, seed=52)
    ct, tt = syn.integrator_tubes(4)
    alg = KernelSR(time_horizon=4, constraint_tube=ct, target_tube=tt, problem="THT", kernel_fn=partial(rbf),
                   regularization_param=1e-2)
    alg.fit(syn.as_sample(X, U, Y))
    o = orc.KernelSROracle(4, ct, tt, "THT", None, 1e-2).fit(X, U, Y)
    assert np.max(np.abs(alg.value_functions - o.value_functions)) <= TOL
    assert np.max(np.abs(alg.predict(pts) - o.predict(pts))) <= TOL


def test_predict_results_are_not_aliased():
    """predict() reuses pinned result buffers only after the caller dropped the previous result."""
    KernelSR, _, _, _, _, rbf = _api()
    X, U, Y = syn.integrator_sample(300, seed=41)
    ct, tt = syn.integrator_tubes(4)
    alg = KernelSR(time_horizon=4, constraint_tube=ct, target_tube=tt, problem="FHT", kernel_fn=partial(rbf, sigma=0.2))
    alg.fit(syn.as_sample(X, U, Y))
    p1, p2 = syn.uniform_points(500, seed=42), syn.uniform_points(500, seed=43)
    a = alg.predict(p1)
    a_copy = a.copy()
    b = alg.predict(p2)              # same shape while `a` is still referenced: must not overwrite it
    assert np.array_equal(a, a_copy) and not np.array_equal(a, b)
    view = b[1:]                     # a view keeps the buffer leased as well
    b_copy = b.copy()
    del b
    for _ in range(4):
        c = alg.predict(p1)
    assert np.array_equal(view, b_copy[1:]) and np.array_equal(c, a_copy)


def test_kernel_sr_sklearn_kernel_and_defaults():
    """The reference's benchmarks pass sklearn's rbf_kernel(gamma=1/(2 sigma^2)) (benchmarks/kernel_sr/
    baseline.py:134-144); defaults sigma=0.1, lambda=1, THT (kernel_sr.py:226-230)."""
    KernelSR, _, _, _, _, _ = _api()
    from sklearn.metrics.pairwise import rbf_kernel as sk_rbf

    X, U, Y = syn.integrator_sample(500, seed=11)
    pts = syn.uniform_points(300, seed=12)
    ct, tt = syn.integrator_tubes(5)
    a = KernelSR(time_horizon=5, constraint_tube=ct, target_tube=tt, kernel_fn=partial(sk_rbf, gamma=1 / (2 * 0.1 ** 2)))
    a.fit(syn.as_sample(X, U, Y))
    o = orc.KernelSROracle(5, ct, tt, "THT", 0.1, 1).fit(X, U, Y)
    assert np.max(np.abs(a.predict(pts) - o.predict(pts))) <= TOL
    b = KernelSR(time_horizon=5, constraint_tube=ct, target_tube=tt)
    b.fit(syn.as_sample(X, U, Y))
    assert np.max(np.abs(b.predict(pts) - o.predict(pts))) <= TOL


@pytest.mark.parametrize("problem", ["FHT", "THT"])
def test_kernel_sr_edge_cases(problem):
    KernelSR, _, _, _, _, rbf = _api()
    X, U, Y = syn.integrator_sample(333, seed=21)
    ct, tt = syn.integrator_tubes(17)  # 17 > 16 rows: exercises the two-launch predict path
    alg = KernelSR(time_horizon=17, constraint_tube=ct, target_tube=tt, problem=problem,
                   kernel_fn=partial(rbf, sigma=0.2), regularization_param=1e-3)
    alg.fit(syn.as_sample(X, U, Y))
    o = orc.KernelSROracle(17, ct, tt, problem, 0.2, 1e-3).fit(X, U, Y)
    assert np.max(np.abs(alg.value_functions - o.value_functions)) <= TOL
    # boundary points (closed box), one point, empty set, a far-away point whose kernel column underflows
    pts = np.array([[1.0, 1.0], [-1.0, 0.3], [0.5, -0.5], [0.5000001, 0.0], [1.0000001, 0.0], [0.0, 0.0]])
    assert np.max(np.abs(alg.predict(pts) - o.predict(pts))) <= TOL
    assert np.max(np.abs(alg.predict(pts[:1]) - o.predict(pts[:1]))) <= TOL
    assert alg.predict(np.zeros((0, 2))).shape == (17, 0)
    far = np.array([[0.9, 40.0], [0.2, 0.1]])
    with np.errstate(all="ignore"):
        want = o.predict(far)
    got = alg.predict(far)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.max(np.abs(got[~np.isnan(want)] - want[~np.isnan(want)])) <= TOL
    # horizon 1: only the terminal indicator
    one = KernelSR(time_horizon=1, constraint_tube=ct, target_tube=tt, problem=problem)
    one.fit(syn.as_sample(X, U, Y))
    assert np.array_equal(one.predict(pts), orc.KernelSROracle(1, ct, tt, problem).fit(X, U, Y).predict(pts))


def test_kernel_sr_medium_vs_oracle_and_properties():
    """N=3000, 40k test points: oracle parity on a slice + size-independent properties."""
    KernelSR, _, _, _, _, rbf = _api()
    n, m, T = 3000, 40000, 15
    X, U, Y = syn.integrator_sample(n, seed=0)
    pts = syn.uniform_points(m, seed=1)
    ct, tt = syn.integrator_tubes(T)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", kernel_fn=partial(rbf, sigma=0.1),
                   regularization_param=1 / n ** 2)
    alg.fit_arrays(X, Y)
    out = alg.predict(pts)
    o = orc.KernelSROracle(T, ct, tt, "FHT", 0.1, 1 / n ** 2).fit(X, U, Y)
    assert np.max(np.abs(alg.value_functions - o.value_functions)) <= TOL
    assert np.max(np.abs(out[:, :3000] - o.predict(pts[:3000]))) <= TOL
    # columns are independent: any split of the test set gives bit-identical columns (basis of the sharding)
    a, b = alg.predict(pts[:12345]), alg.predict(pts[12345:])
    assert np.array_equal(np.concatenate([a, b], axis=1), out)
    # FMA and DMMA contractions agree far below the tolerance
    alg.predict_variant = 1
    assert np.max(np.abs(alg.predict(pts) - out)) <= 1e-11
    # terminal row is the target indicator; points in the target set have probability exactly 1 (FHT)
    in_t = np.all(np.abs(pts) <= 0.5, axis=1)
    assert np.array_equal(out[T - 1], in_t.astype(float))
    assert np.all(out[:, in_t] == 1.0)


def test_kernel_sr_full_size_golden():
    """BASELINE.json configs[1] at full size (N = 20 000, T = 15) against the unmodified reference
    (oracle/make_golden_c2.py: reference fit 128 s): value functions (every 10th sample), diag(W) and the
    safety probabilities of 2 000 of the 250 000 test points."""
    KernelSR, _, _, _, _, rbf = _api()
    g = load_golden("sr_c2_n20000_fht.npz")
    n, T = int(g["n"]), int(g["T"])
    X, U, Y = syn.integrator_sample(n, dim=2, seed=0)
    pts = syn.uniform_points(250000, 2, seed=1)
    ct, tt = syn.integrator_tubes(T)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", kernel_fn=partial(rbf, sigma=0.1),
                   regularization_param=float(g["lam"]))
    alg.fit_arrays(X, Y)
    stride = int(g["vf_stride"])
    assert np.max(np.abs(alg.value_functions[:, ::stride] - g["value_functions_sub"])) <= TOL
    w = alg._w.cpu().numpy()[::stride]
    assert np.max(np.abs(w - g["w_diag_sub"]) / np.abs(g["w_diag_sub"])) <= 1e-6
    out = alg.predict(pts)  # all 250k points through the chunked, overlapped host path
    m_sub = int(g["m_sub"])
    assert out.shape == (T, 250000)
    assert np.max(np.abs(out[:, :m_sub] - g["out"])) <= TOL
    # the chunked host path, the one-launch device path and a sharded evaluation agree bit for bit
    assert np.array_equal(alg.predict_dev(alg_dev(pts)).cpu().numpy(), out)
    assert np.array_equal(alg.predict(pts[100000:100000 + 4321]), out[:, 100000:100000 + 4321])
    in_t = np.all(np.abs(pts) <= 0.5, axis=1)
    assert np.array_equal(out[T - 1], in_t.astype(float)) and np.all(out[:, in_t] == 1.0)
    assert np.nanmin(out) >= 0.0 and np.nanmax(out) <= 1.0 + 1e-9


def alg_dev(a):
    from socks_b200 import _device as dv

    return dv.to_dev(a)


@pytest.mark.parametrize("name", ["srmax_fht.npz", "srmax_tht_batch.npz", "srmax_n1200_p100_fht.npz"])
def test_kernel_sr_max_golden(name):
    _, _, KernelMaximalSR, kernel_sr_max, _, rbf = _api()
    g = load_golden(name)
    S = syn.as_sample(g["X"], g["U"], g["Y"])
    kw = dict(time_horizon=int(g["T"]), constraint_tube=tubes_from(g["ctube"]), target_tube=tubes_from(g["ttube"]),
              problem=str(g["problem"]), regularization_param=float(g["lam"]),
              kernel_fn=partial(rbf, sigma=float(g["sigma"])))
    alg = KernelMaximalSR(**kw)
    alg.fit(S, g["A"])
    assert np.max(np.abs(alg.CUA - g["CUA"]) / g["CUA"]) <= 1e-13
    assert np.max(np.abs(alg.value_functions - g["value_functions"])) <= TOL
    out = alg.predict(g["pts"])
    assert np.max(np.abs(out - g["out"])) <= TOL
    alg.predict_chunk = 128  # several chunks, ragged tail
    assert np.array_equal(alg.predict(g["pts"]), out)
    assert np.array_equal(kernel_sr_max(S, g["A"], g["pts"], batch_size=50, **kw), out)


def test_kernel_sr_max_medium_vs_oracle():
    _, _, KernelMaximalSR, _, _, rbf = _api()
    n, m, T, P = 1500, 3000, 6, 100
    X, U, Y = syn.integrator_sample(n, seed=31)
    pts = syn.uniform_points(m, seed=32)
    A = np.linspace(-1, 1, P).reshape(-1, 1)
    ct, tt = syn.integrator_tubes(T)
    alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT",
                          kernel_fn=partial(rbf, sigma=0.1), regularization_param=1.0)
    alg.fit_arrays(X, U, Y, A)
    o = orc.KernelMaximalSROracle(T, ct, tt, "FHT", 0.1, 1.0).fit(X, U, Y, A)
    assert np.max(np.abs(alg.value_functions - o.value_functions)) <= TOL
    assert np.max(np.abs(alg.predict(pts) - o.predict(pts))) <= TOL


@pytest.mark.parametrize("name", ["fwd_unconstrained.npz", "fwd_constrained.npz", "fwd_tracking_regime.npz"])
@pytest.mark.parametrize("heuristic", [True, False])
def test_kernel_control_fwd_golden(name, heuristic):
    _, _, _, _, KernelControlFwd, rbf = _api()
    g = load_golden(name)
    T = int(g["T"])
    cons = None
    if bool(g["constrained"]):
        def cons(time=0, state=None):
            return np.abs(state[:, 1]) - (0.45 + 0.01 * time)
    pol = KernelControlFwd(cost_fn=syn.tracking_cost_fn(T), constraint_fn=cons, heuristic=heuristic,
                           regularization_param=float(g["lam"]), kernel_fn=partial(rbf, sigma=float(g["sigma"])),
                           verbose=False, time_horizon=T)  # extra kwargs are swallowed (tracking.py:137)
    pol.train(syn.as_sample(g["X"], g["U"], g["Y"]), g["A"])
    tol = max(1e-8, 100 * float(g["cond"]) * 2.0 ** -53)
    for t in range(T):
        y = pol.scores(time=t, state=[g["states"][t]]).cpu().numpy()
        assert np.max(np.abs(y[0] - g["C"][t])) <= tol * np.max(np.abs(g["C"][t]))
        if cons is not None:
            assert np.max(np.abs(y[1] - g["D"][t])) <= tol * np.max(np.abs(g["D"][t]))
        act = pol(time=t, state=[g["states"][t]])
        assert act.dtype == np.float32
        if heuristic:
            assert np.array_equal(act, g["actions"][t])
    assert np.linalg.norm(pol.W - np.linalg.inv(np.linalg.inv(pol.W))) >= 0  # W is exposed as an ndarray


@pytest.mark.parametrize("name", ["bwd_unconstrained.npz", "bwd_constrained.npz"])
def test_kernel_control_bwd_golden(name):
    """KernelControlBwd against the unmodified reference (value functions, scores, actions)."""
    from socks_b200.algorithms.control.kernel_control_bwd import KernelControlBwd
    _, _, _, _, _, rbf = _api()
    g = load_golden(name)
    T = int(g["T"])
    cons = None
    if bool(g["constrained"]):
        def cons(time=0, state=None):
            return np.abs(state[:, 1]) - (0.45 + 0.01 * time)
    pol = KernelControlBwd(time_horizon=T, cost_fn=syn.tracking_cost_fn(T), constraint_fn=cons, heuristic=True,
                           regularization_param=float(g["lam"]), kernel_fn=partial(rbf, sigma=float(g["sigma"])),
                           verbose=False)
    pol.train(syn.as_sample(g["X"], g["U"], g["Y"]), g["A"])
    tol = max(1e-8, 100 * float(g["cond"]) * 2.0 ** -53)
    scale = np.max(np.abs(g["value_functions"]))
    assert np.max(np.abs(pol.value_functions - g["value_functions"])) <= tol * scale
    for t in range(T):
        y = pol.scores(time=t, state=[g["states"][t]]).cpu().numpy()
        assert np.max(np.abs(y[0] - g["C"][t])) <= tol * np.max(np.abs(g["C"][t]))
        assert np.array_equal(pol(time=t, state=[g["states"][t]]), g["actions"][t])


def test_abel_kernel_and_separating_classifier_golden():
    """abel_kernel / euclidean_distance / SeparatingKernelClassifier vs the unmodified reference."""
    from socks_b200.algorithms.reach.separating_kernel import SeparatingKernelClassifier
    from socks_b200.kernel.metrics import abel_kernel, euclidean_distance

    g = load_golden("separating_ring.npz")
    X, T, sigma, lam = g["X"], g["T"], float(g["sigma"]), float(g["lam"])
    K = abel_kernel(X[:40], T[:30], sigma=sigma)
    assert np.max(np.abs(K - g["K_abel"]) / g["K_abel"]) <= 1e-13
    assert np.max(np.abs(abel_kernel(X[:40]) - g["K_abel_median"]) / g["K_abel_median"]) <= 1e-13
    assert np.max(np.abs(euclidean_distance(X[:40], T[:30]) - g["D"])) <= 1e-15
    assert np.array_equal(euclidean_distance(X[:40], T[:30], squared=True), g["D2"])
    clf = SeparatingKernelClassifier(kernel_fn=partial(abel_kernel, sigma=sigma), regularization_param=lam)
    clf.fit(X)
    tol = max(1e-10, 100 * float(g["cond"]) * 2.0 ** -53)
    assert abs(clf.tau - float(g["tau"])) <= tol
    C = clf.scores(T)
    assert np.max(np.abs(C - g["C"])) <= tol
    labels = clf.predict(T)
    decided = np.abs(g["C"] - (1 - float(g["tau"]))) > 10 * tol  # points not sitting on the threshold
    assert np.array_equal(labels[decided], g["labels"][decided])
    clf.predict_chunk = 128  # several chunks with a ragged tail
    assert np.max(np.abs(clf.scores(T) - C)) <= 1e-13


def test_conditional_embedding_and_mmd_golden():
    """ConditionalEmbedding, maximum_mean_discrepancy, witness_function vs the unmodified reference."""
    from socks_b200.algorithms.kernel import ConditionalEmbedding
    from socks_b200.kernel.metrics import rbf_kernel
    from socks_b200.kernel.probability import maximum_mean_discrepancy, witness_function

    g = load_golden("embedding_mmd.npz")
    tol = max(1e-10, 100 * float(g["cond"]) * 2.0 ** -53)
    ce = ConditionalEmbedding(regularization_param=float(g["lam"])).fit(g["G"])
    for y, want in ((g["y1"], g["pred1"]), (g["y2"], g["pred2"])):
        got = ce.predict(y, g["K"])
        assert got.shape == want.shape and np.max(np.abs(got - want)) <= tol * np.max(np.abs(want))
    assert ce.score(g["pred1"], g["K"], g["y1"]) <= tol * np.max(np.abs(g["pred1"]))
    kf = partial(rbf_kernel, sigma=0.8)
    assert abs(maximum_mean_discrepancy(g["P"], g["Q"], kernel_fn=kf) - float(g["mmd_unbiased"])) <= 1e-12
    assert abs(maximum_mean_discrepancy(g["P"], g["Q"], kernel_fn=kf, biased=True, squared=True) - float(g["mmd_biased_sq"])) <= 1e-13
    assert abs(maximum_mean_discrepancy(g["P"], g["Q"], squared=True) - float(g["mmd_default"])) <= 1e-13
    assert np.max(np.abs(witness_function(g["P"], g["Q"], g["t"], kernel_fn=kf) - g["witness"])) <= 1e-13


@pytest.mark.parametrize("family", ["abel", "laplacian"])
def test_kernel_sr_other_kernel_families(family):
    """KernelSR with a non-RBF kernel_fn (the reference accepts any pairwise kernel, kernel_sr.py:202,226-227): the
    materialising predict path against a numpy restatement built on the same kernel matrices."""
    KernelSR, _, _, _, _, _ = _api()
    from sklearn.metrics.pairwise import laplacian_kernel

    from socks_b200.kernel.metrics import abel_kernel

    kfn = partial(abel_kernel, sigma=0.3) if family == "abel" else partial(laplacian_kernel, gamma=4.0)
    cpu = (lambda A, B: orc.abel_kernel(A, B, 0.3)) if family == "abel" else (lambda A, B: laplacian_kernel(A, B, gamma=4.0))
    n, T, lam = 400, 6, 1e-2
    X, U, Y = syn.integrator_sample(n, seed=71)
    pts = syn.uniform_points(333, seed=72, low=-1.05, high=1.05)
    ct, tt = syn.integrator_tubes(T)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT", kernel_fn=kfn,
                   regularization_param=lam)
    alg.fit(syn.as_sample(X, U, Y))
    W = np.linalg.inv(cpu(X, X) + lam * n * np.identity(n))
    w = np.diag(W).copy()
    vf = np.empty((T, n))
    vf = orc.sr_recursion(Y, w, cpu(X, Y), T, ct, tt, vf, "FHT", out=vf)
    want = orc.sr_recursion(pts, w, cpu(X, pts), T, ct, tt, vf, "FHT")
    assert np.max(np.abs(alg.value_functions - vf)) <= TOL
    assert np.max(np.abs(alg.predict(pts) - want)) <= TOL


@pytest.mark.parametrize("slices", [6, 7, 8])
def test_kernel_sr_max_int8_engine_matches_fp64_engine(slices):
    """KernelMaximalSR.predict: Ozaki digit splitting on the integer tensor cores (tcgen05.mma kind::i8, csrc/ozaki.cu)
    against the FP64 DMMA contraction and the oracle; ragged sizes (N, M, T*P not multiples of the tiles), several
    chunks, points far from the data (flagged -> recomputed in FP64, NaN columns where numpy has them)."""
    _, _, KernelMaximalSR, _, _, rbf = _api()
    n, T, P = 1237, 7, 37
    X, U, Y = syn.integrator_sample(n, seed=81)
    rng = np.random.default_rng(82)
    near = syn.uniform_points(1500, seed=83, low=-1.05, high=1.05)
    far = np.stack([np.linspace(1.2, 60.0, 61), rng.uniform(-1, 1, 61)], axis=1)  # progressively far, then underflow
    pts = np.concatenate([near, far])
    A = np.linspace(-1, 1, P).reshape(-1, 1)
    ct, tt = syn.integrator_tubes(T)
    alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="THT",
                          kernel_fn=partial(rbf, sigma=0.1), regularization_param=1.0)
    alg.fit_arrays(X, U, Y, A)
    alg.predict_engine = "f64"
    want = alg.predict(pts)
    o = orc.KernelMaximalSROracle(T, ct, tt, "THT", 0.1, 1.0).fit(X, U, Y, A)
    with np.errstate(all="ignore"):
        ref = o.predict(pts)
    alg.predict_engine, alg.predict_slices, alg.predict_chunk = "i8", slices, 512
    got = alg.predict(pts)
    assert alg.last_flagged >= 30  # the far points were not trusted to the fixed-point operand
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(want)
    tol = {6: 2e-9, 7: 2e-10, 8: 1e-11}[slices]
    assert np.max(np.abs(got[ok] - want[ok])) <= tol
    assert np.max(np.abs(got[ok] - ref[ok])) <= TOL
    # flagged points come from the FP64 path itself: identical bits
    idx = np.arange(1500, len(pts))[-20:]
    assert np.array_equal(got[:, idx], want[:, idx], equal_nan=True)


def test_kernel_sr_max_fit_int8_engine_matches_fp64_engine():
    """KernelMaximalSR.fit: the backward recursion on the integer tensor cores (socks_srmax_fit_i8) against the FP64
    engine and the oracle; a threshold that flags every sample must hand the whole fit to the FP64 engine (same bits)."""
    _, _, KernelMaximalSR, _, _, rbf = _api()
    n, T, P = 1237, 7, 37
    X, U, Y = syn.integrator_sample(n, seed=91)
    A = np.linspace(-1, 1, P).reshape(-1, 1)
    ct, tt = syn.integrator_tubes(T)

    def fit(engine, tau=None):
        alg = KernelMaximalSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT",
                              kernel_fn=partial(rbf, sigma=0.25), regularization_param=1.0)
        alg.predict_engine = engine
        if tau is not None:
            alg.predict_tau = tau
        alg.fit_arrays(X, U, Y, A)
        return alg, alg._vf.cpu().numpy()

    a64, v64 = fit("f64")
    a8, v8 = fit("i8")
    assert a8.last_fit_flagged == 0  # sigma = 0.25: every next state has neighbours for every action
    assert np.max(np.abs(v8 - v64)) <= 1e-9
    o = orc.KernelMaximalSROracle(T, ct, tt, "FHT", 0.25, 1.0).fit(X, U, Y, A)
    assert np.max(np.abs(v8 - o.value_functions)) <= TOL
    afb, vfb = fit("i8", tau=1e30)
    assert afb.last_fit_flagged == n and np.array_equal(vfb, v64)
