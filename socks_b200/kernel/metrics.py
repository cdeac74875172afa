"""`gym_socks.kernel.metrics` drop-in: `rbf_kernel`, `regularized_inverse` (GPU, fp64).

Same names, argument order, defaults, assertion behaviour and return types as
`gym_socks/kernel/metrics.py:69-138` and `:209-280`; the arithmetic runs in libsocksb200
(gram.cu, chol.cu).  `kernel_fn` follows the reference's callable protocol
(`kernel_fn(X, Y=None) -> (n, m) ndarray`, metrics.py:239-241) but must be recognisable as an RBF
kernel (this package's or gym_socks's `rbf_kernel` / `abel_kernel`, or sklearn's `rbf_kernel`, `linear_kernel`,
`polynomial_kernel`, `laplacian_kernel` -- the four the reference's own test feeds to `regularized_inverse`,
`gym_socks/kernel/tests/test_kernel.py:18,132-147` -- bare or wrapped in `functools.partial` with keyword
parameters): arbitrary Python callables cannot run on the GPU and there is no CPU fallback, so anything else
raises TypeError.
"""
from functools import partial

import numpy as np
import torch

from socks_b200 import _device as dv
from socks_b200 import _lib

__all__ = ["rbf_kernel", "abel_kernel", "euclidean_distance", "regularized_inverse", "RBFSpec", "resolve_kernel",
           "Factorization"]

KIND_RBF, KIND_ABEL, KIND_LINEAR, KIND_POLY, KIND_LAPLACIAN = 0, 1, 2, 3, 4  # SOCKS_KERNEL_* of include/socks_b200.h


# ----------------------------------------------------------------------------- kernel_fn protocol
class RBFSpec:
    """Resolved RBF kernel: k(x, y) = exp(-gamma |x - y|^2).

    `sigma`-style (gym_socks semantics): gamma = 1 / (2 sigma^2) (metrics.py:135); sigma=None is
    the median heuristic sigma = median(squared distances) evaluated per call (metrics.py:130-131).
    `gamma`-style (sklearn semantics): gamma used as is; gamma=None -> 1 / n_features.
    """

    def __init__(self, sigma=None, gamma=None, style="sigma", kind=KIND_RBF, coef0=1.0, degree=3.0):
        self.sigma, self.gamma, self.style, self.kind = sigma, gamma, style, kind
        self.coef0, self.degree = float(coef0), float(degree)  # polynomial_kernel only
        if style == "sigma" and sigma is not None:
            assert sigma > 0, "sigma must be a strictly positive real value."

    def params_for(self, Xd, Yd):
        """(scale, divide) of the C ABI for one kernel evaluation between device arrays Xd (n, d), Yd (m, d):
        k = exp(d2 / scale) with scale = -2 sigma^2 (the reference divides, metrics.py:135) or
        k = exp(d2 * scale) with scale = -gamma (sklearn multiplies)."""
        if self.kind == KIND_LINEAR:
            return (1.0, 0)
        if self.style == "gamma":  # sklearn: gamma=None -> 1 / n_features
            g = float(self.gamma) if self.gamma is not None else 1.0 / Xd.shape[1]
            return (g, 0) if self.kind == KIND_POLY else (-g, 0)
        if self.kind == KIND_ABEL:  # exp(dist / -sigma), sigma=None -> median of the distances (metrics.py:199-204)
            sigma = float(self.sigma) if self.sigma is not None else _median_sqdist(Xd, Yd, squared=False)
            return (-sigma, 1)
        sigma = float(self.sigma) if self.sigma is not None else _median_sqdist(Xd, Yd)
        return (-2 * (sigma ** 2), 1)

    def gamma_for(self, Xd, Yd):
        """gamma such that k = exp(-gamma d2) (for reporting / rescaling)."""
        scale, divide = self.params_for(Xd, Yd)
        return -1.0 / scale if divide else -scale

    @property
    def extra(self):
        """(coef0, degree) of the C ABI (read by SOCKS_KERNEL_POLY only)."""
        return (self.coef0, self.degree)

    @property
    def separable(self):
        """k([x|u], [x'|u']) == k(x, x') k(u, u') for one shared scale (product Gram as ONE kernel on [X | U])."""
        return self.kind == KIND_RBF

    @property
    def data_dependent(self):
        return self.style == "sigma" and self.sigma is None


_SKLEARN_KINDS = {"rbf_kernel": KIND_RBF, "linear_kernel": KIND_LINEAR, "polynomial_kernel": KIND_POLY,
                  "laplacian_kernel": KIND_LAPLACIAN}
_SKLEARN_KEYWORDS = {"rbf_kernel": {"gamma"}, "linear_kernel": {"dense_output"},
                     "polynomial_kernel": {"gamma", "degree", "coef0"}, "laplacian_kernel": {"gamma"}}


def _is_rbf_function(fn):
    return callable(fn) and getattr(fn, "__name__", "") in ("rbf_kernel", "abel_kernel", "linear_kernel",
                                                            "polynomial_kernel", "laplacian_kernel")


def resolve_kernel(kernel_fn, default_sigma):
    """Map the reference's `kernel_fn` argument to an RBFSpec (TypeError if it is not an RBF)."""
    if kernel_fn is None:
        return RBFSpec(sigma=default_sigma)
    if isinstance(kernel_fn, RBFSpec):
        return kernel_fn
    fn, kw = kernel_fn, {}
    if isinstance(kernel_fn, partial):
        if kernel_fn.args:
            raise TypeError("kernel_fn partials must bind keyword arguments only (sigma= or gamma=).")
        fn, kw = kernel_fn.func, dict(kernel_fn.keywords or {})
    if not _is_rbf_function(fn):
        raise TypeError(
            "socks_b200 runs the kernel on the GPU and accepts the kernel families it implements there: "
            "rbf_kernel / abel_kernel (this package's or gym_socks's, sigma=...) and sklearn's rbf_kernel, "
            f"linear_kernel, polynomial_kernel, laplacian_kernel; got {kernel_fn!r}. There is no CPU fallback."
        )
    module = getattr(fn, "__module__", "") or ""
    kw.pop("distance_fn", None)  # accepted and ignored by the reference (metrics.py:114-117)
    if module.startswith("sklearn") or fn.__name__ in ("linear_kernel", "polynomial_kernel", "laplacian_kernel"):
        extra = set(kw) - _SKLEARN_KEYWORDS[fn.__name__]
        if extra:
            raise TypeError(f"unsupported keyword(s) for sklearn {fn.__name__}: {sorted(extra)}")
        return RBFSpec(gamma=kw.get("gamma"), style="gamma", kind=_SKLEARN_KINDS[fn.__name__],
                       coef0=kw.get("coef0", 1), degree=kw.get("degree", 3))
    extra = set(kw) - {"sigma"}
    if extra:
        raise TypeError(f"unsupported keyword(s) for {fn.__name__}: {sorted(extra)}")
    kind = KIND_ABEL if fn.__name__ == "abel_kernel" else KIND_RBF
    return RBFSpec(sigma=kw.get("sigma"), style="sigma", kind=kind)


# ----------------------------------------------------------------------------- device primitives
def gram_dev(Xd, Yd, kparams, rowscale=None, out=None, ld=None, kind=KIND_RBF, extra=(1.0, 3.0)):
    """K = rowscale[:, None] * k(X_i, Y_j) on the device, kparams = (scale, divide); returns an (n, ld) tensor
    whose first m columns hold K (ld = m rounded up to even so rows stay 16-byte aligned)."""
    n, d = Xd.shape
    m = Yd.shape[0]
    if ld is None:
        ld = dv.even(m)
    if out is None:
        out = dv.empty(n, ld)
    _lib.call("socks_gram_family", int(kind), dv.ptr(Xd), n, dv.ptr(Yd), m, d, float(kparams[0]), int(kparams[1]),
              float(extra[0]), float(extra[1]), dv.ptr(rowscale), dv.ptr(out), ld, dv.stream())
    return out


def sqdist_dev(Xd, Yd, squared=True):
    n, d = Xd.shape
    m = Yd.shape[0]
    out = dv.empty(n, m)
    _lib.call("socks_sqdist" if squared else "socks_dist", dv.ptr(Xd), n, dv.ptr(Yd), m, d, dv.ptr(out), m, dv.stream())
    return out


def _median_sqdist(Xd, Yd, squared=True):
    """np.median of the (squared) distance matrix (metrics.py:130-131, :199-200): mean of the two middle order
    statistics for an even count, by radix select on the device (csrc/compose.cu: socks_median_sq_dist)."""
    if Xd.shape[0] * Yd.shape[0] * 8 > MEDIAN_BYTES_LIMIT:
        raise ValueError(f"sigma=None takes the median of the full {Xd.shape[0]} x {Yd.shape[0]} distance matrix "
                         "(metrics.py:130-131), which does not fit the device workspace limit; pass an explicit sigma.")
    n, d = Xd.shape
    m = Yd.shape[0]
    lib = _lib.load()
    wk = dv.work(lib.socks_median_work_bytes(n, m))
    res = dv.empty(1)
    _lib.call("socks_median_sq_dist", dv.ptr(Xd), n, dv.ptr(Yd), m, d, 1 if squared else 0, dv.ptr(wk), dv.ptr(res),
              dv.stream())
    return float(res.item())


MEDIAN_BYTES_LIMIT = 16 << 30  # largest distance matrix the median heuristic may materialise on the device


def sharded_bandwidth_guard(spec):
    """sigma=None makes the bandwidth a function of ALL test points: a rank that only sees its shard would use a
    different kernel than the single-GPU / reference run.  Raise instead of silently diverging."""
    import torch.distributed as dist

    if spec.data_dependent and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        raise ValueError("sigma=None (median heuristic) depends on ALL test points: compute the bandwidth once on the "
                         "full set and pass kparams to predict_dev on every rank.")


_ENGINE = {"factorization": "int8"}
# e = (max l_ii / min l_ii)^2, read off the diagonal of the Cholesky factor: a (loose, 100-5000x) lower bound of cond_2(G).
# Above this value Factorization.check() redoes the inverse with its GEMMs on the FP64 tensor cores (the trailing updates
# stay on the integer ones).  A safety net beyond what has been measured: with the merge products balanced along K
# (csrc/ozaki.cu) the integer inverse's residual max|W G - I| equals the FP64 engine's up to e = 1.7e8 (cond 8e11), the
# worst case tried (tools/illcond_check.py); without the balancing it was 7x / 19x / 500x worse at e = 3.3e5 / 3.3e7 / 1.7e8.
# Every BASELINE config has e <= 3.9e4 (tools/cond_estimates.py).
COND_LIMIT_INT8 = 1e10


def set_factorization_engine(engine="int8"):
    """Which tensor cores the large GEMMs of the Cholesky + inverse run on (process-wide, csrc/chol.cu):
    "int8" (default): fp64-accurate digit GEMMs on tcgen05.mma kind::i8 -- 7 digits in the trailing updates, 8 in the
    merges of inv(L); same residuals as fp64 on every matrix measured, 2.2x faster.
    "fp64": the FP64 DMMA GEMM everywhere, for matrices whose rows span more dynamic range than the digit scheme carries.
    Not part of the reference API (gym_socks has no such choice); call it before fitting."""
    if engine not in ("int8", "fp64"):
        raise ValueError(f"engine must be 'int8' or 'fp64', not {engine!r}")
    _ENGINE["factorization"] = engine
    lib = _lib.load()
    lib.socks_potrf_inv_set_block(-4 if engine == "int8" else -3)
    lib.socks_potrf_inv_set_block(-6 if engine == "int8" else -5)


class Factorization:
    """G = K(Z,Z) + lambda N I = L L^T on the device, with the pieces of inv(G) the path needs.

    metrics.py:278-280 (`inv(K + regularization_param * num_rows * I)`) computed as Cholesky + triangular
    inverse in one recursion (socks_potrf_inv); `diag()` is diag(inv(G)) (all the SR algorithms read,
    kernel_sr.py:45), `full()` is inv(G) itself.  `keep_l=True` also leaves the complete factor L in `self.L`;
    `two_pass=True` uses the separate socks_potrf + socks_trtri entry points instead of the fused recursion."""

    def __init__(self, Zd, kparams, diag_add, kind=KIND_RBF, keep_l=False, two_pass=False, Ud=None, extra=(1.0, 3.0)):
        n, d = Zd.shape
        self.n, self.np = n, dv.padded(n)
        self._recipe = ("gram", Zd, kparams, float(diag_add), int(kind), bool(keep_l), bool(two_pass), Ud, extra)
        self._w = None
        self._W = None
        self._fp64_inverse = _ENGINE["factorization"] != "int8"
        self._factor()

    def _factor(self):
        """Build G and run the factorisation recorded in self._recipe (again, with the merges of inv(L) on the FP64
        tensor cores, when check() finds the matrix too ill-conditioned for the digit GEMMs: see COND_LIMIT_INT8)."""
        npad, n = self.np, self.n
        st = dv.stream()
        lib = _lib.load()
        what = self._recipe[0]
        keep_l = two_pass = False
        G = dv.empty(npad, npad)
        dv.mark("start")
        if what == "gram":
            _, Zd, kparams, diag_add, kind, keep_l, two_pass, Ud, extra = self._recipe
            d = Zd.shape[1]
            if Ud is None and kind in (KIND_RBF, KIND_ABEL):
                _lib.call("socks_gram_sym_kernel", int(kind), dv.ptr(Zd), n, d, float(kparams[0]), int(kparams[1]),
                          float(diag_add), dv.ptr(G), npad, st)
            else:  # any family, optional second factor k(U, U) with the same kernel (metrics.py:268-276)
                _lib.call("socks_gram_sym_family", int(kind), dv.ptr(Zd), n, d, dv.ptr(Ud), 0 if Ud is None else Ud.shape[1],
                          float(kparams[0]), int(kparams[1]), float(extra[0]), float(extra[1]), float(diag_add),
                          dv.ptr(G), npad, st)
            device = Zd.device
        else:  # caller-supplied matrix
            _, Gd, diag_add = self._recipe
            _lib.call("socks_pad_spd", dv.ptr(Gd), n, Gd.stride(0), float(diag_add), dv.ptr(G), npad, st)
            device = Gd.device
        self.info = torch.zeros(1, dtype=torch.int32, device=device)
        self._lrange = torch.zeros(2, dtype=torch.float64, device=device)
        self._M = dv.empty(npad, npad)
        wk = dv.work(lib.socks_trtri_work_bytes(n))
        dv.mark("gram_sym")
        if two_pass:
            dinv = dv.work(lib.socks_potrf_work_bytes(n))
            _lib.call("socks_potrf", dv.ptr(G), n, npad, dv.ptr(dinv), dv.ptr(self.info), st)
            _lib.call("socks_trtri", dv.ptr(G), n, npad, dv.ptr(dinv), dv.ptr(self._M), npad, dv.ptr(wk), st)
            self._fp64_inverse = True
        else:
            fallback = _ENGINE["factorization"] == "int8" and self._fp64_inverse  # only check() asks for this
            if fallback:
                lib.socks_potrf_inv_set_block(-5)  # merges of inv(L) on the FP64 tensor cores for this call
            try:
                _lib.call("socks_potrf_inv", dv.ptr(G), n, npad, dv.ptr(self._M), npad, dv.ptr(wk), dv.ptr(self.info),
                          1 if keep_l else 0, st)
            finally:
                if fallback:
                    lib.socks_potrf_inv_set_block(-6)
        # the diagonal of L survives in G's diagonal blocks either way: (max / min)^2 <= cond_2(G)
        _lib.call("socks_diag_range", dv.ptr(G), n, npad, dv.ptr(self._lrange), st)
        self.L = G if (keep_l or two_pass) else None  # otherwise only scratch is left in it
        dv.mark("potrf_inv")
        if self._w is not None:  # a refactorisation: refresh what was already handed out, in place
            w, self._w = self._w, None
            w.copy_(self.diag())
            self._w = w
        self._W = None

    @classmethod
    def from_matrix(cls, Gd, diag_add):
        """Factor a caller-supplied symmetric positive definite matrix (device, n x n) + diag_add * I."""
        self = cls.__new__(cls)
        n = Gd.shape[0]
        self.n, self.np = n, dv.padded(n)
        self._recipe = ("matrix", Gd, float(diag_add))
        self._w = None
        self._W = None
        self._fp64_inverse = _ENGINE["factorization"] != "int8"
        self._factor()
        return self

    def check(self):
        info = int(self.info.item())
        if info != 0:
            raise np.linalg.LinAlgError(
                f"regularised Gram matrix is not positive definite (pivot {info - 1}); the reference's LU inverse "
                "would return a meaningless matrix here"
            )
        lo, hi = self._lrange.tolist()
        self.cond_estimate = (hi / lo) ** 2 if lo > 0 else float("inf")
        if not self._fp64_inverse and self.cond_estimate > COND_LIMIT_INT8:
            # the rows of inv(L) span more dynamic range than 8 digits carry: redo with the merges (and W = M^T M) on the
            # FP64 tensor cores; the trailing updates stay on the integer ones (measured as accurate as FP64 up to
            # cond 1e12: tools/illcond_check.py)
            self._fp64_inverse = True
            self._factor()
            info = int(self.info.item())
            if info != 0:
                raise np.linalg.LinAlgError(f"regularised Gram matrix is not positive definite (pivot {info - 1})")

    def inverse_factor(self):
        if self._M is None:
            raise RuntimeError("the inverse factor was released")
        return self._M

    def diag(self):
        """w = diag(inv(G)) as a device vector of length n."""
        if self._w is None:
            M = self.inverse_factor()
            self._w = dv.empty(self.n)
            wk = dv.work(_lib.load().socks_reduce_work_bytes(self.np))
            _lib.call("socks_diag_inv", dv.ptr(M), self.n, self.np, dv.ptr(self._w), dv.ptr(wk), dv.stream())
            dv.mark("diag_inv")
        return self._w

    def full(self):
        """inv(G) as a device (np, np) tensor; the logical matrix is [:n, :n]."""
        if self._W is None:
            M = self.inverse_factor()
            self._W = dv.empty(self.np, self.np)
            if not self._fp64_inverse and self.np > 2048:
                wk = dv.work(_lib.load().socks_lauum_i8_work_bytes(self.n))
                _lib.call("socks_lauum_i8", dv.ptr(M), self.n, self.np, dv.ptr(self._W), self.np, dv.ptr(wk), dv.stream())
            else:
                _lib.call("socks_lauum", dv.ptr(M), self.n, self.np, dv.ptr(self._W), self.np, dv.stream())
        return self._W

    def release_factors(self):
        """Drop L and M once only diag()/full() are needed (3.2 GB each at N = 20k)."""
        self.L = None
        self._M = None


# ----------------------------------------------------------------------------- public API
def rbf_kernel(X, Y=None, sigma=None, distance_fn=None):
    """RBF kernel Gram matrix, `gym_socks/kernel/metrics.py:69-138`.

    `Y=None` means `Y = X` (:119-120); `sigma=None` uses the median of the squared distances
    (:130-131); `sigma <= 0` raises AssertionError (:133); `distance_fn` is accepted and ignored
    exactly as in the reference (:114-117).  Returns a float64 ndarray of shape (len(X), len(Y)).
    """
    Xh = dv.as2d(X)
    Yh = Xh if Y is None else dv.as2d(Y)
    if sigma is not None:
        assert sigma > 0, "sigma must be a strictly positive real value."
    n, m = Xh.shape[0], Yh.shape[0]
    if n == 0 or m == 0:
        return np.zeros((n, m))
    Xd = dv.to_dev(Xh)
    Yd = Xd if Y is None else dv.to_dev(Yh)
    K = gram_dev(Xd, Yd, RBFSpec(sigma=sigma).params_for(Xd, Yd))
    return K[:, :m].cpu().numpy()


def abel_kernel(X, Y=None, sigma=None, distance_fn=None):
    """Abel kernel exp(-|x - y| / sigma), `gym_socks/kernel/metrics.py:174-206` (sigma=None: median of the
    distances, `:199-200`).  A user-supplied `distance_fn` cannot run on the GPU and raises TypeError."""
    if distance_fn is not None and getattr(distance_fn, "__name__", "") != "euclidean_distance":
        raise TypeError("abel_kernel on the GPU only supports the Euclidean distance (no CPU fallback)")
    Xh = dv.as2d(X)
    Yh = Xh if Y is None else dv.as2d(Y)
    if sigma is not None:
        assert sigma > 0, "sigma must be a strictly positive real value."
    n, m = Xh.shape[0], Yh.shape[0]
    if n == 0 or m == 0:
        return np.zeros((n, m))
    Xd = dv.to_dev(Xh)
    Yd = Xd if Y is None else dv.to_dev(Yh)
    spec = RBFSpec(sigma=sigma, kind=KIND_ABEL)
    return gram_dev(Xd, Yd, spec.params_for(Xd, Yd), kind=KIND_ABEL)[:, :m].cpu().numpy()


def euclidean_distance(X, Y=None, squared=False):
    """Pairwise (squared) Euclidean distance by direct differences, `gym_socks/kernel/metrics.py:26-66`."""
    Xh = dv.as2d(X)
    Yh = Xh if Y is None else dv.as2d(Y)
    n, m = Xh.shape[0], Yh.shape[0]
    if n == 0 or m == 0:
        return np.zeros((n, m))
    Xd = dv.to_dev(Xh)
    Yd = Xd if Y is None else dv.to_dev(Yh)
    return sqdist_dev(Xd, Yd, squared=bool(squared)).cpu().numpy()


def regularized_inverse(X, Y=None, U=None, V=None, regularization_param=None, kernel_fn=None):
    """Regularised inverse `inv(K + lambda N I)`, `gym_socks/kernel/metrics.py:209-280`.

    Defaults and assertions follow the reference: lambda = 1/N^2 (:251-252) else `> 0` (:254-256);
    kernel sigma = 1 (:258-259); `X.shape == Y.shape` (:261-262); product Gram with `U` (:268-276).
    As every caller in the reference, `Y` and `V` must be `X` and `U` themselves (the matrix has to be
    symmetric positive definite for the Cholesky-based solve); other values raise ValueError.
    """
    Xh = dv.as2d(X)
    if Y is not None:
        Yh = dv.as2d(Y)
        assert Xh.shape == Yh.shape, "Parameters %r, %r must have the same shape." % (X, Y)
        if not np.array_equal(Xh, Yh):
            raise ValueError("regularized_inverse: only Y = X (symmetric Gram) is supported on the GPU path.")
    n = Xh.shape[0]
    if regularization_param is None:
        regularization_param = 1 / (n ** 2)
    else:
        assert regularization_param > 0, "regularization_param must be a strictly positive real value."
    spec = resolve_kernel(kernel_fn, default_sigma=1)
    if U is not None:
        Uh = dv.as2d(U)
        assert Xh.shape[0] == Uh.shape[0], "Parameters %r, %r must have the same sample size." % (X, U)
        if V is not None:
            Vh = dv.as2d(V)
            assert Uh.shape == Vh.shape, "Parameters %r, %r must have the same shape." % (U, V)
            if not np.array_equal(Uh, Vh):
                raise ValueError("regularized_inverse: only V = U is supported on the GPU path.")
    fac = factorize(Xh, None if U is None else Uh, spec, regularization_param)
    W = fac.full()
    fac.check()
    return W[:n, :n].cpu().numpy()


def factorize(Xh, Uh, spec, regularization_param, keep_l=False, two_pass=False):
    """Host arrays -> device Factorization of K(X,X) [* K(U,U)] + lambda N I.

    The product of two RBF kernels with the same bandwidth is the RBF kernel of the concatenated
    vectors, so the product Gram of metrics.py:268-276 is one kernel evaluation on Z = [X | U].
    With the median heuristic each factor has its own bandwidth; the columns are rescaled so a single
    gamma applies.  The other kernel families multiply the two factors explicitly."""
    n = Xh.shape[0]
    Xd = dv.to_dev(Xh)
    kp = spec.params_for(Xd, Xd)
    Ud = None
    if Uh is None:
        Zd = Xd
    else:
        Ud0 = dv.to_dev(Uh)
        ku = spec.params_for(Ud0, Ud0)  # what the reference's kernel_fn(U, V) call would use
        if spec.separable:
            if ku == kp:
                Zd = dv.to_dev(np.concatenate([Xh, Uh], axis=1))
            else:  # data-dependent bandwidths differ: rescale U so one scale applies
                gx, gu = spec.gamma_for(Xd, Xd), spec.gamma_for(Ud0, Ud0)
                Zd = dv.to_dev(np.concatenate([Xh, Uh * np.sqrt(gu / gx)], axis=1))
        else:  # general product Gram k(X,X) * k(U,U) (socks_gram_sym_family), one parameter set for both factors
            if ku != kp:
                raise NotImplementedError(f"product Gram with different kernel parameters for states and actions "
                                          f"({kp} vs {ku}): pass explicit kernel parameters")
            Zd, Ud = Xd, Ud0
    return Factorization(Zd, kp, regularization_param * n, kind=spec.kind, keep_l=keep_l, two_pass=two_pass, Ud=Ud,
                         extra=spec.extra)
