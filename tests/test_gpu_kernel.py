"""GPU parity: rbf_kernel / regularized_inverse / Cholesky pieces through the C ABI vs the oracle,
the reference's known-answer tests and the golden vectors.  Tolerances:
rbf_kernel rel <= 1e-13; W rel-Frobenius <= 100 * cond * 2^-53."""
from functools import partial

import numpy as np
import pytest

from conftest import load_golden
from oracle import socks_oracle as orc

pytestmark = pytest.mark.gpu


def _mods():
    import torch
    from socks_b200 import _device as dv, _lib
    from socks_b200.kernel import metrics

    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch, dv, _lib, metrics


def test_known_answers_from_reference_tests():
    _, _, _, m = _mods()
    # gym_socks/kernel/tests/test_kernel.py:46-66
    Y = np.arange(4).reshape((2, 2))
    gt = np.array([[1.0, 0.77880078], [0.77880078, 1.0]], dtype=np.float32)
    assert np.allclose(m.rbf_kernel(Y), gt)
    assert np.allclose(m.rbf_kernel(Y, Y), gt)
    # :68-100
    Y3 = np.arange(6).reshape((3, 2))
    gt3 = np.array([[1.0, 0.93941306, 0.77880078], [0.93941306, 1.0, 0.93941306], [0.77880078, 0.93941306, 1.0]])
    assert np.allclose(m.rbf_kernel(Y3), gt3)
    assert np.allclose(m.rbf_kernel(Y3, Y3, distance_fn=lambda *a, **k: None), gt3)  # distance_fn ignored
    assert m.rbf_kernel(Y, Y3).shape == (2, 3)
    # :118-130
    gtw = np.array([[0.66676608, -0.0081415], [-0.0081415, 0.66676608]])
    assert np.allclose(m.regularized_inverse(np.arange(4).reshape((2, 2))), gtw)


def test_rbf_golden_and_median():
    _, _, _, m = _mods()
    g = load_golden("kernel_metrics.npz")
    X, Y = g["X"], g["Y"]
    for got, want in [
        (m.rbf_kernel(X, Y, sigma=0.7), g["K_xy_s07"]),
        (m.rbf_kernel(X, sigma=0.1), g["K_xx_s01"]),
        (m.rbf_kernel(X), g["K_xx_median"]),
        (m.rbf_kernel(np.arange(4).reshape(2, 2), np.arange(6).reshape(3, 2)), g["K_arange22_32_median"]),
    ]:
        assert got.dtype == np.float64 and got.shape == want.shape
        assert np.max(np.abs(got - want) / np.maximum(want, 1e-300)) <= 1e-13


@pytest.mark.parametrize("n,m,d,sigma", [(1, 1, 1, 1.0), (37, 23, 3, 0.7), (130, 257, 2, 0.1), (300, 129, 8, 2.0),
                                         (65, 33, 11, 1.5), (33, 1000, 5, 0.9), (1029, 7, 2, 0.05)])
def test_rbf_random_vs_oracle(n, m, d, sigma):
    _, _, _, mt = _mods()
    rng = np.random.default_rng(n * 1000 + m)
    X, Y = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    K, Ko = mt.rbf_kernel(X, Y, sigma=sigma), orc.rbf_kernel(X, Y, sigma)
    big = Ko > 1e-290
    assert np.max(np.abs(K[big] - Ko[big]) / Ko[big]) <= 1e-13
    assert np.max(np.abs(K[~big] - Ko[~big]), initial=0.0) <= 1e-290  # underflow region: both (sub)zero
    # inputs may be lists / ints / float32 (kernel_sr.py:282-284 does np.array on them)
    Ki = mt.rbf_kernel(X.astype(np.float32).tolist(), Y.astype(np.float32), sigma=sigma)
    Kio = orc.rbf_kernel(X.astype(np.float32).astype(np.float64), Y.astype(np.float32), sigma)
    big = Kio > 1e-290
    assert np.max(np.abs(Ki[big] - Kio[big]) / Kio[big]) <= 1e-13
    assert np.max(np.abs(Ki[~big] - Kio[~big]), initial=0.0) <= 1e-290


def test_rbf_empty_and_underflow():
    _, _, _, mt = _mods()
    assert mt.rbf_kernel(np.zeros((0, 2)), np.zeros((5, 2)), sigma=1.0).shape == (0, 5)
    far = mt.rbf_kernel(np.zeros((1, 2)), np.array([[1e3, 0.0], [3.0, 0.0]]), sigma=0.1)
    want = orc.rbf_kernel(np.zeros((1, 2)), np.array([[3.0, 0.0]]), 0.1)[0, 0]  # exp(9 / (-2 * 0.1**2))
    assert far[0, 0] == 0.0 and abs(far[0, 1] / want - 1) < 1e-15


def test_sklearn_partial_accepted():
    _, _, _, mt = _mods()
    from sklearn.metrics.pairwise import rbf_kernel as sk_rbf

    rng = np.random.default_rng(3)
    X = rng.standard_normal((50, 2))
    W = mt.regularized_inverse(X, kernel_fn=partial(sk_rbf, gamma=0.8), regularization_param=1e-2)
    Wo = np.linalg.inv(sk_rbf(X, X, gamma=0.8) + 1e-2 * 50 * np.eye(50))
    assert np.linalg.norm(W - Wo) / np.linalg.norm(Wo) < 1e-11


@pytest.mark.parametrize("key,kw", [
    ("W_x_default", {}),
    ("W_x_s07_lam", dict(sigma=0.7, lam=1e-3)),
    ("W_xu_s07_lam", dict(sigma=0.7, lam=1e-3, U=True)),
])
def test_regularized_inverse_golden(key, kw):
    _, _, _, mt = _mods()
    g = load_golden("kernel_metrics.npz")
    X = g["X"]
    args = {}
    if "sigma" in kw:
        args["kernel_fn"] = partial(mt.rbf_kernel, sigma=kw["sigma"])
        args["regularization_param"] = kw["lam"]
    if kw.get("U"):
        args["U"] = g["U"]
    W = mt.regularized_inverse(X, **args)
    want = g[key]
    cond = np.linalg.cond(want)
    assert W.shape == want.shape and np.allclose(W, W.T, rtol=0, atol=1e-14 * np.abs(W).max())
    assert np.linalg.norm(W - want) / np.linalg.norm(want) <= 100 * cond * 2.0 ** -53


@pytest.mark.parametrize("n,d,sigma,lam,with_u", [(128, 2, 0.3, 1e-2, False), (257, 3, 0.5, 1e-3, True),
                                                  (1000, 2, 0.1, 1e-6, False), (1500, 3, 3.0, 1e-7, True)])
@pytest.mark.parametrize("two_pass", [False, True])
def test_factorization_pieces(n, d, sigma, lam, with_u, two_pass):
    """potrf residual, M L = I, diag(W) and full W against float64 numpy (cond-aware), through the fused
    socks_potrf_inv recursion and through the separate socks_potrf + socks_trtri entry points."""
    torch, dv, _lib, mt = _mods()
    rng = np.random.default_rng(n)
    X = rng.uniform(-1.1, 1.1, (n, d))
    U = rng.uniform(-1, 1, (n, 2)) if with_u else None
    spec = mt.RBFSpec(sigma=sigma)
    fac = mt.factorize(X, U, spec, lam, keep_l=True, two_pass=two_pass)
    G = orc.rbf_kernel(X, X, sigma)
    if with_u:
        G = G * orc.rbf_kernel(U, U, sigma)
    G[np.diag_indices(n)] += lam * n
    L = np.tril(fac.L[:n, :n].cpu().numpy())
    fac.check()
    assert np.linalg.norm(L @ L.T - G) / np.linalg.norm(G) < 1e-14
    M = np.tril(fac.inverse_factor()[:n, :n].cpu().numpy())
    cond = np.linalg.cond(G)
    assert np.linalg.norm(M @ L - np.eye(n)) / np.sqrt(n) < 50 * np.sqrt(cond) * 2.0 ** -53 * n ** 0.5
    Wref = np.linalg.inv(G)
    w = fac.diag().cpu().numpy()
    assert np.max(np.abs(w - np.diag(Wref)) / np.diag(Wref)) <= 100 * cond * 2.0 ** -53
    W = fac.full()[:n, :n].cpu().numpy()
    assert np.linalg.norm(W - Wref) / np.linalg.norm(Wref) <= 100 * cond * 2.0 ** -53
    # padding must stay the identity
    npad = fac.np
    if npad > n:
        Wp = fac.full().cpu().numpy()
        assert np.array_equal(Wp[n:, n:], np.eye(npad - n)) and not Wp[n:, :n].any() and not Wp[:n, n:].any()


def test_ozaki_gemm_k_balancing_recovers_opposite_profiles():
    """Operands whose magnitudes run in opposite directions along K (what L21 and inv(L11) look like in a merge): every
    term a_ik b_jk is of order one, but without balancing the small factors are truncated against their row maximum."""
    torch, dv, _lib, mt = _mods()
    lib = _lib.load()
    m = n = 256
    K = 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    prof = torch.pow(10.0, torch.linspace(0, -12, K, dtype=torch.float64, device="cuda"))
    A = torch.randn(m, K, dtype=torch.float64, device="cuda", generator=g) * prof
    B = torch.randn(n, K, dtype=torch.float64, device="cuda", generator=g) / prof
    ref = (A.cpu().numpy().astype(np.longdouble) @ B.cpu().numpy().T.astype(np.longdouble)).astype(np.float64)
    wk = dv.work(lib.socks_dev_ozaki_gemm_work_bytes(m, n, K) + 2048)
    base = (wk.data_ptr() + 1023) // 1024 * 1024
    err = {}
    for name, flags in (("plain", 0), ("balanced", 4)):
        C = torch.zeros(m, n, dtype=torch.float64, device="cuda")
        _lib.call("socks_dev_ozaki_gemm_nt", A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), n, m, n, K, flags, 0, base,
                  dv.stream())
        torch.cuda.synchronize()
        err[name] = np.abs(-C.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err["balanced"] < 1e-14 and err["plain"] > 1e3 * err["balanced"]


def test_ill_conditioned_matrix_on_the_integer_engine_and_its_guard():
    """cond(G) = 5e10: the merge products' operands are balanced along K before the digit split, so the integer inverse is
    as accurate as the FP64 one; lowering COND_LIMIT_INT8 exercises the safety net (inverse redone on FP64)."""
    torch, dv, _lib, mt = _mods()
    rng = np.random.default_rng(3)
    n, sigma, lam = 3000, 1.0, 1e-11
    X = rng.uniform(-1.1, 1.1, (n, 2))
    G = orc.rbf_kernel(X, X, sigma)
    G[np.diag_indices(n)] += lam * n
    res = {}
    limit = mt.COND_LIMIT_INT8
    try:
        for name, engine, lim in (("fp64", "fp64", limit), ("int8", "int8", limit), ("guard", "int8", 1e5)):
            mt.set_factorization_engine(engine)
            mt.COND_LIMIT_INT8 = lim
            fac = mt.factorize(X, None, mt.RBFSpec(sigma=sigma), lam)
            fac.check()
            assert fac.cond_estimate > 1e6
            assert fac._fp64_inverse == (name != "int8")
            res[name] = np.abs(fac.full()[:n, :n].cpu().numpy() @ G - np.eye(n)).max()
    finally:
        mt.COND_LIMIT_INT8 = limit
        mt.set_factorization_engine("int8")
    assert res["int8"] <= 2 * res["fp64"] and res["guard"] <= 2 * res["fp64"]


def test_factorization_engine_switch():
    """metrics.set_factorization_engine: the FP64 DMMA engine and the integer tensor-core engine give the same inverse."""
    torch, dv, _lib, mt = _mods()
    rng = np.random.default_rng(11)
    n = 3000
    X = rng.uniform(-1.1, 1.1, (n, 2))
    res = {}
    try:
        for engine in ("fp64", "int8"):
            mt.set_factorization_engine(engine)
            fac = mt.factorize(X, None, mt.RBFSpec(sigma=0.2), 1e-4)
            fac.check()
            res[engine] = (fac.diag().cpu().numpy(), fac.full()[:n, :n].cpu().numpy())
    finally:
        mt.set_factorization_engine("int8")
    with pytest.raises(ValueError):
        mt.set_factorization_engine("fp32")
    G = orc.rbf_kernel(X, X, 0.2)
    G[np.diag_indices(n)] += 1e-4 * n
    cond = np.linalg.cond(G)
    assert np.max(np.abs(res["int8"][0] - res["fp64"][0]) / res["fp64"][0]) <= 100 * cond * 2.0 ** -53
    for engine in ("fp64", "int8"):
        R = res[engine][1] @ G - np.eye(n)
        assert np.abs(R).max() <= 100 * cond * 2.0 ** -53


@pytest.mark.parametrize("n,nb,keep_l", [(3300, 1024, True), (2900, 512, False), (1400, 256, True), (2100, 1024, False)])
def test_potrf_inv_lookahead_matches_recursion_and_numpy(n, nb, keep_l):
    """socks_potrf_inv's look-ahead factorisation (second stream, block columns of `nb`, ragged last block) against the
    single-stream recursion (nb = 0) and float64 numpy: L residual, M L = I, diag(inv G)."""
    torch, dv, _lib, mt = _mods()
    lib = _lib.load()
    rng = np.random.default_rng(n)
    X = rng.uniform(-1.1, 1.1, (n, 2))
    G = orc.rbf_kernel(X, X, 0.2)
    G[np.diag_indices(n)] += 1e-4 * n
    cond = np.linalg.cond(G)
    res = {}
    old = lib.socks_potrf_inv_set_block(nb)
    try:
        for mode, block in (("lookahead", nb), ("lookahead_dmma", nb), ("recursion", 0)):
            lib.socks_potrf_inv_set_block(block)
            lib.socks_potrf_inv_set_block(-3 if mode == "lookahead_dmma" else -4)  # trailing updates: FP64 DMMA / int8
            fac = mt.factorize(X, None, mt.RBFSpec(sigma=0.2), 1e-4, keep_l=keep_l)
            fac.check()
            res[mode] = (np.tril(fac.inverse_factor()[:n, :n].cpu().numpy()), fac.diag().cpu().numpy(),
                         np.tril(fac.L[:n, :n].cpu().numpy()) if keep_l else None)
    finally:
        lib.socks_potrf_inv_set_block(old)
        lib.socks_potrf_inv_set_block(-4)
    Md, wd, _ = res["lookahead_dmma"]
    assert np.linalg.norm(res["lookahead"][0] - Md) / np.linalg.norm(Md) <= 100 * cond * 2.0 ** -53
    M, w, L = res["lookahead"]
    M0, w0, L0 = res["recursion"]
    Lref = np.linalg.cholesky(G)
    if keep_l:
        assert np.linalg.norm(L @ L.T - G) / np.linalg.norm(G) < 1e-14
        assert np.linalg.norm(L - L0) / np.linalg.norm(L0) < 1e-13
    assert np.linalg.norm(M @ Lref - np.eye(n)) / np.sqrt(n) < 50 * np.sqrt(cond) * 2.0 ** -53 * n ** 0.5
    assert np.linalg.norm(M - M0) / np.linalg.norm(M0) <= 100 * cond * 2.0 ** -53
    wref = np.diag(np.linalg.inv(G))
    assert np.max(np.abs(w - wref) / wref) <= 100 * cond * 2.0 ** -53
    assert np.max(np.abs(w - w0) / w0) <= 100 * cond * 2.0 ** -53


@pytest.mark.parametrize("m,n,K,syrk,lower,digits", [(256, 128, 96, False, False, 7), (384, 384, 1024, True, True, 7),
                                                     (128, 256, 16384 + 64, False, False, 7), (256, 128, 32, False, False, 7),
                                                     (256, 256, 2048, False, False, 8), (384, 384, 14336 + 96, True, True, 8),
                                                     (256, 256, 14336 + 96, False, False, 8)])
def test_ozaki_gemm_is_fp64_accurate(m, n, K, syrk, lower, digits):
    """C -= A B^T on the integer tensor cores (7 balanced base-256 digits, tcgen05.mma kind::i8; csrc/ozaki.cu) against
    float64 with operands spanning 12 decades per row; K > 16384 exercises the chunked int32 accumulation."""
    torch, dv, _lib, mt = _mods()
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(m + n + K)
    A = torch.randn(m, K, dtype=torch.float64, device="cuda", generator=g)
    A = A * torch.pow(10.0, torch.empty(m, 1, dtype=torch.float64, device="cuda").uniform_(-6, 6, generator=g))
    B = A if syrk else torch.randn(n, K, dtype=torch.float64, device="cuda", generator=g)
    C0 = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    C = C0.clone()
    wk = dv.work(lib.socks_dev_ozaki_gemm_work_bytes(m, n, K) + 2048)
    base = (wk.data_ptr() + 1023) // 1024 * 1024
    _lib.call("socks_dev_ozaki_gemm_nt", A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), n, m, n, K,
              (1 if syrk else 0) | (2 if digits == 8 else 0), 1 if lower else 0, base, dv.stream())
    torch.cuda.synchronize()
    A_, B_, C0_, C_ = A.cpu().numpy(), B.cpu().numpy(), C0.cpu().numpy(), C.cpu().numpy()
    ref = C0_ - (A_.astype(np.longdouble) @ B_.T.astype(np.longdouble)).astype(np.float64)
    # error model of the digit scheme: operands are truncated at 2^-54 of their ROW maximum and digit pairs with
    # s + t >= 7 are dropped (<= 6 * 2^-54 per sample, worst case), so the bound is relative to K * rowmax(A) * rowmax(B)
    # (worst case 13 * 2^-53 of it: 1 from the truncations, 12 from the dropped pairs) -- plus the rounding of C itself.
    bound = K * np.abs(A_).max(axis=1)[:, None] * np.abs(B_).max(axis=1)[None, :]
    # C is rounded once per pass and accumulation chunk (2 passes x up to 2 chunks here), `ref` once
    err = (np.abs(C_ - ref) - 6 * 2.0 ** -53 * np.maximum(np.abs(ref), np.abs(C0_))) / bound
    if lower:  # only 128 x 64 tiles touching the lower triangle are written
        i, j = np.indices((m, n))
        keep = (j // 64) * 64 <= (i // 128) * 128 + 127
        assert np.array_equal(C_[~keep], C0_[~keep])
        err = err[keep]
    assert err.max() < 13 * 2.0 ** -53 / (256 if digits == 8 else 1)  # one more digit: 256x smaller truncation


def test_not_positive_definite_reports():
    torch, dv, _lib, mt = _mods()
    n = 130
    X = np.zeros((n, 2))  # all points identical: G = ones + tiny ridge -> numerically singular
    with pytest.raises((np.linalg.LinAlgError, AssertionError)):
        fac = mt.Factorization(dv.to_dev(X), (-1.0, 0), -0.5)  # negative shift: pivot <= 0
        fac.check()


def test_gemv_and_argmin_primitives():
    torch, dv, _lib, mt = _mods()
    rng = np.random.default_rng(5)
    for n, m in [(1, 1), (70, 33), (1000, 257), (513, 1000)]:
        B = rng.standard_normal((n, m))
        V = rng.standard_normal((2, n))
        Bd, Vd = dv.to_dev(B), dv.to_dev(V)
        y = dv.empty(2, m)
        wk = dv.work(_lib.load().socks_reduce_work_bytes(m))
        _lib.call("socks_gemv_t", dv.ptr(Bd), n, m, m, dv.ptr(Vd), n, 2, dv.ptr(y), m, dv.ptr(wk), dv.stream())
        want = V @ B
        assert np.max(np.abs(y.cpu().numpy() - want)) <= 1e-12 * max(1.0, np.abs(want).max())
    idx = torch.zeros(1, dtype=torch.int32, device="cuda")
    for P in (1, 7, 256, 1000):
        C = rng.standard_normal(P)
        C[rng.integers(P)] = C.min()  # tie: first occurrence wins
        D = rng.standard_normal(P)
        for Dm in (None, D, np.abs(D) + 1):
            Cd = dv.to_dev(C)
            Dd = None if Dm is None else dv.to_dev(Dm)
            _lib.call("socks_argmin_masked", dv.ptr(Cd), dv.ptr(Dd), P, dv.ptr(idx), dv.stream())
            assert int(idx.item()) == orc.heuristic_solution(C, Dm)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0)])
@pytest.mark.parametrize("cfg", [0, 1])
def test_gemm_f64_forms(ta, tb, cfg):
    """Every operand form / tile configuration of the DMMA GEMM against numpy, incl. alpha/beta, the
    triangular k-range flags and lower-only output."""
    torch, dv, _lib, mt = _mods()
    rng = np.random.default_rng(10 * ta + tb)
    m, n, K = 256, 384, 272
    A = rng.standard_normal((K, m) if ta else (m, K))
    B = rng.standard_normal((n, K) if tb else (K, n))
    C0 = rng.standard_normal((m, n))
    opA = A.T if ta else A
    opB = B.T if tb else B
    Ad, Bd = dv.to_dev(A), dv.to_dev(B)
    for alpha, beta in [(1.0, 0.0), (-1.0, 1.0), (0.5, -2.0)]:
        Cd = dv.to_dev(C0)
        _lib.call("socks_gemm_f64", ta, tb, dv.ptr(Ad), A.shape[1], dv.ptr(Bd), B.shape[1], dv.ptr(Cd), n, m, n, K,
                  alpha, beta, 0, 0, cfg, dv.stream())
        want = alpha * (opA @ opB) + beta * C0
        assert np.max(np.abs(Cd.cpu().numpy() - want)) <= 1e-12 * np.abs(want).max()
    # triangular operands (square case): A lower (k <= i) times B lower in (k, j) sense (k >= j), lower output only
    s = 384
    L1 = np.tril(rng.standard_normal((s, s)))
    L2 = np.tril(rng.standard_normal((s, s)))
    if not ta and not tb:
        Ad, Bd, Cd = dv.to_dev(L1), dv.to_dev(L2), dv.zeros(s, s)
        _lib.call("socks_gemm_f64", 0, 0, dv.ptr(Ad), s, dv.ptr(Bd), s, dv.ptr(Cd), s, s, s, s, 1.0, 0.0, 1 | 4, 1, cfg,
                  dv.stream())
        got, want = Cd.cpu().numpy(), L1 @ L2
        for bi in range(3):
            for bj in range(bi + 1):
                blk = (slice(128 * bi, 128 * bi + 128), slice(128 * bj, 128 * bj + 128))
                assert np.max(np.abs(got[blk] - want[blk])) <= 1e-12 * np.abs(want).max()
        assert not got[:128, 128:].any()  # tiles above the block diagonal untouched
    if ta and not tb:  # lauum form: M^T M with M lower, k >= max(i, j)
        Ad, Cd = dv.to_dev(L1), dv.zeros(s, s)
        _lib.call("socks_gemm_f64", 1, 0, dv.ptr(Ad), s, dv.ptr(Ad), s, dv.ptr(Cd), s, s, s, s, 1.0, 0.0, 2 | 4, 1, cfg,
                  dv.stream())
        got, want = np.tril(Cd.cpu().numpy()), np.tril(L1.T @ L1)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.abs(want).max()


def test_potrs_against_numpy_solve():
    torch, dv, _lib, mt = _mods()
    rng = np.random.default_rng(9)
    n, q = 700, 5
    X = rng.uniform(-1, 1, (n, 2))
    fac = mt.factorize(X, None, mt.RBFSpec(sigma=0.3), 1e-3)
    fac.check()
    G = orc.rbf_kernel(X, X, 0.3)
    G[np.diag_indices(n)] += 1e-3 * n
    Bh = rng.standard_normal((n, q))
    npad = fac.np
    B = dv.zeros(npad, 128)
    B[:n, :q].copy_(dv.to_dev(Bh))
    wk = dv.empty(npad, 128)
    _lib.call("socks_potrs", dv.ptr(fac.inverse_factor()), n, npad, dv.ptr(B), 128, 128, dv.ptr(wk), dv.stream())
    want = np.linalg.solve(G, Bh)
    got = B[:n, :q].cpu().numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 100 * np.linalg.cond(G) * 2.0 ** -53


# ---------------------------------------------------------------------------------------------- kernel families
def test_sklearn_kernels_work_with_regularized_inverse():
    """Port of gym_socks/kernel/tests/test_kernel.py:132-147 (`test_sklearn_kernels`): the four sklearn pairwise
    kernels must be accepted by regularized_inverse(Y, Y, kernel_fn=...) -- and here also give the reference's matrix."""
    from sklearn.metrics.pairwise import laplacian_kernel, linear_kernel, polynomial_kernel, rbf_kernel

    from socks_b200.kernel.metrics import regularized_inverse

    Y = np.arange(4).reshape((2, 2))
    for kernel_fn in [linear_kernel, polynomial_kernel, rbf_kernel, laplacian_kernel]:
        W = regularized_inverse(Y, Y, kernel_fn=kernel_fn)
        want = np.linalg.inv(kernel_fn(Y, Y) + (1 / 4) * 2 * np.identity(2))  # metrics.py:251-252, 278-280
        assert W.shape == (2, 2) and np.allclose(W, want, rtol=1e-12, atol=0), kernel_fn.__name__


@pytest.mark.parametrize("name", ["linear", "poly_default", "poly_params", "laplacian", "laplacian_gamma"])
def test_kernel_families_inverse_and_product_gram(name):
    """regularized_inverse with each family, with and without the product Gram K(U,U) * K(X,X) (metrics.py:268-276),
    against numpy on sklearn's matrices (cond-aware, as for the RBF W)."""
    from functools import partial

    from sklearn.metrics.pairwise import laplacian_kernel, linear_kernel, polynomial_kernel

    from socks_b200.kernel.metrics import regularized_inverse

    fn = {"linear": linear_kernel, "poly_default": polynomial_kernel,
          "poly_params": partial(polynomial_kernel, gamma=0.3, degree=2, coef0=0.5), "laplacian": laplacian_kernel,
          "laplacian_gamma": partial(laplacian_kernel, gamma=2.5)}[name]
    rng = np.random.default_rng(11)
    X = rng.uniform(-1, 1, (300, 3))
    U = rng.uniform(-1, 1, (300, 3))  # same width as X: sklearn's gamma=None is 1 / n_features of its argument
    lam = 1e-3
    for Uarg in (None, U):
        G = fn(X, X) if Uarg is None else fn(Uarg, Uarg) * fn(X, X)
        G = G + lam * 300 * np.identity(300)
        want = np.linalg.inv(G)
        got = regularized_inverse(X, U=Uarg, regularization_param=lam, kernel_fn=fn)
        tol = max(1e-10, 100 * np.linalg.cond(G) * 2.0 ** -53)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) <= tol


def test_kernel_family_rejections():
    from sklearn.metrics.pairwise import sigmoid_kernel

    from socks_b200.kernel.metrics import regularized_inverse

    with pytest.raises(TypeError):
        regularized_inverse(np.eye(3), kernel_fn=sigmoid_kernel)
    with pytest.raises(TypeError):
        regularized_inverse(np.eye(3), kernel_fn=lambda X, Y=None: X @ X.T)
    # an indefinite matrix is reported, not silently "inverted" through a failed Cholesky
    from functools import partial

    from sklearn.metrics.pairwise import polynomial_kernel

    with pytest.raises(np.linalg.LinAlgError):
        regularized_inverse(np.array([[1.0, 0.0], [0.0, 2.0], [3.0, 1.0]]), regularization_param=1e-9,
                            kernel_fn=partial(polynomial_kernel, degree=1, coef0=-5.0))


@pytest.mark.parametrize("n,m,d,squared", [(37, 23, 3, True), (40, 40, 2, False), (301, 7, 4, True), (1, 1, 2, True),
                                            (2000, 1500, 2, True)])
def test_median_select_matches_numpy(n, m, d, squared):
    """socks_median_sq_dist (radix select of both middle order statistics) against np.median of the oracle's matrix:
    the sigma=None bandwidth of metrics.py:130-131 / :199-200, bit for bit."""
    torch, dv, _lib, mt = _mods()
    rng = np.random.default_rng(n + m)
    X, Y = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    if n == 40:
        Y = X.copy()  # zeros on the diagonal take part in the median (test_kernel.py:102-115)
    D = orc.sq_dist(X, Y)
    want = np.median(D if squared else np.sqrt(D))
    got = mt._median_sqdist(dv.to_dev(X), dv.to_dev(Y), squared=squared)
    assert got == want or abs(got - want) <= 4 * np.finfo(float).eps * want  # sqrt: device vs numpy rounding


def test_context_workspace_and_stream():
    import ctypes

    torch, dv, _lib, mt = _mods()
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    _lib.call("socks_ctx_create", torch.cuda.current_device(), ctypes.byref(ctx))
    assert lib.socks_ctx_stream(ctx)
    p1, p2, p3 = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    _lib.call("socks_ctx_workspace", ctx, 1 << 20, ctypes.byref(p1))
    _lib.call("socks_ctx_workspace", ctx, 1 << 10, ctypes.byref(p2))   # smaller: same arena
    assert p1.value and p1.value == p2.value
    _lib.call("socks_ctx_workspace", ctx, 1 << 22, ctypes.byref(p3))   # larger: reallocated
    assert p3.value
    # a Gram matrix computed on the context's stream into the context's workspace
    X = dv.to_dev(np.arange(4.0).reshape(2, 2))
    torch.cuda.synchronize()
    _lib.call("socks_gram_rbf", dv.ptr(X), 2, dv.ptr(X), 2, 2, -2.0, 1, None, p3, 2, lib.socks_ctx_stream(ctx))
    _lib.call("socks_ctx_synchronize", ctx)
    K = torch.empty(4, dtype=torch.float64, device="cuda")
    _lib.call("socks_copy2d", K.data_ptr(), 32, p3, 32, 32, 1, 3, None)
    torch.cuda.synchronize()
    assert np.allclose(K.cpu().numpy().reshape(2, 2), [[1.0, np.exp(-4.0)], [np.exp(-4.0), 1.0]])
    _lib.call("socks_ctx_destroy", ctx)


def test_composite_sr_entry_points_match_the_wrappers():
    """socks_sr_fit and socks_sr_predict (pack + predict in one call, arbitrary 16-byte aligned workspace) called directly
    through ctypes give the bits KernelSR gives; predict()'s host paths (direct pinned-host stores, overlapped column
    blocks, forced block counts) agree bit for bit."""
    torch, dv, _lib, mt = _mods()
    from functools import partial

    from socks_b200 import synthetic as syn
    from socks_b200.algorithms.reach.common import PROBLEM_CODE, pack_tubes
    from socks_b200.algorithms.reach.kernel_sr import KernelSR

    lib = _lib.load()
    n, T, d, m = 777, 9, 2, 5001
    X, U, Y = syn.integrator_sample(n, d, seed=4)
    pts_h = syn.uniform_points(m, d, seed=5, low=-1.05, high=1.05)
    ct, tt = syn.integrator_tubes(T, d)
    alg = KernelSR(time_horizon=T, constraint_tube=ct, target_tube=tt, problem="FHT",
                   kernel_fn=partial(mt.rbf_kernel, sigma=0.15), regularization_param=1e-4)
    alg.fit_arrays(X, Y)
    Xd, Yd, pts = dv.to_dev(X), dv.to_dev(Y), dv.to_dev(pts_h)
    tubes = pack_tubes(ct, tt, T, d)
    scale = -2 * 0.15 ** 2
    vf = dv.empty(T, n)
    wk = dv.work(lib.socks_sr_fit_work_bytes(n))
    _lib.call("socks_sr_fit", dv.ptr(Xd), dv.ptr(Yd), dv.ptr(alg._w), n, d, scale, 1, dv.ptr(tubes), PROBLEM_CODE["FHT"], T,
              dv.ptr(vf), n, None, dv.ptr(wk), dv.stream())
    assert torch.equal(vf, alg._vf)
    out = dv.empty(T, m)
    work = dv.work(lib.socks_sr_predict_work_bytes(n, m) + 64)
    for off in (0, 16, 48):  # the entry point only requires 16-byte alignment
        _lib.call("socks_sr_predict", dv.ptr(Xd), dv.ptr(alg._w), dv.ptr(vf), n, n, d, scale, 1, dv.ptr(pts), m,
                  dv.ptr(tubes), PROBLEM_CODE["FHT"], T, dv.ptr(out), m, work.data_ptr() + off, 0, dv.stream())
        assert torch.equal(out, alg.predict_dev(pts))
    ref = alg.predict(pts_h)
    assert np.array_equal(ref, out.cpu().numpy())
    for chunks, direct in ((1, False), (3, False), (None, True), (None, False)):
        alg.predict_chunks, alg.predict_direct_host = chunks, direct
        assert np.array_equal(alg.predict(pts_h), ref)
