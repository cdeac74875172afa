"""Separating kernel classifier on the GPU: drop-in for
`gym_socks/algorithms/reach/separating_kernel.py` (`SeparatingKernelClassifier`), SURVEY §8(f) rank 2.

Reference: W = inv(K + lambda N I) with the Abel kernel (`:86-88`), tau = 1 - min diag(K^T W K / n) (`:90`),
predict diag(K_T^T W K_T / n) >= 1 - tau (`:108-111`).  With G = L L^T and M = inv(L):
k^T W k = |M k|^2, so the N x N x M triple product becomes one triangular DMMA product Z = M K(X,T)
(socks_gemm_f64 with the lower-triangular k-range clipping) and one column sum-of-squares pass
(socks_colsumsq); W itself is only formed if the `.W` attribute is read.
"""
import logging

import numpy as np

from socks_b200 import _device as dv
from socks_b200 import _lib
from socks_b200.kernel.metrics import RBFSpec, factorize, gram_dev, resolve_kernel

logger = logging.getLogger(__name__)

__all__ = ["SeparatingKernelClassifier"]


class SeparatingKernelClassifier:
    """`gym_socks/algorithms/reach/separating_kernel.py:20-114`; defaults: abel_kernel(sigma=0.1), lambda = 1 (`:55-62`)."""

    predict_chunk = 16384  # test points per product (multiple of 128)

    def __init__(self, kernel_fn=None, regularization_param=None, *args, **kwargs):
        self.kernel_fn = kernel_fn
        self.regularization_param = regularization_param

    def _validate_params(self, S=None):
        if self.kernel_fn is None:
            self._spec = RBFSpec(sigma=0.1, kind=1)
        else:
            self._spec = resolve_kernel(self.kernel_fn, default_sigma=0.1)
        if self.regularization_param is None:
            self.regularization_param = 1

    def _validate_data(self, X):
        if X is None:
            raise ValueError("Must supply a sample.")

    def _scores_dev(self, pts):
        """diag(K^T W K) / n for the columns K = k(X, pts): device vector of len(pts)."""
        n, npad = self._n, self._fac.np
        m = pts.shape[0]
        lib = _lib.load()
        st = dv.stream()
        M = self._fac.inverse_factor()
        out = dv.empty(m)
        mc = min(self.predict_chunk, -(-m // 128) * 128)
        Kmat, Z = dv.zeros(npad, mc), dv.empty(npad, mc)
        wk = dv.work(lib.socks_reduce_work_bytes(mc))
        kp = self._spec.params_for(self._Xd, pts)
        for s in range(0, m, mc):
            cnt = min(mc, m - s)
            if cnt < mc:
                Kmat.zero_()
            gram_dev(self._Xd, pts[s:s + cnt], kp, out=Kmat, ld=mc, kind=self._spec.kind, extra=self._spec.extra)
            # Z = M K, M lower triangular: k-range of each output tile clipped at its row block (flag 1)
            _lib.call("socks_gemm_f64", 0, 0, dv.ptr(M), npad, dv.ptr(Kmat), mc, dv.ptr(Z), mc, npad, mc, npad, 1.0, 0.0,
                      1, 0, 0, st)
            _lib.call("socks_colsumsq", dv.ptr(Z), n, cnt, mc, dv.ptr(out[s:]), dv.ptr(wk), st)
        return out / n  # scalar scaling of an m-vector (tensor plumbing, not on the O(N^2 M) path)

    def fit(self, X):
        self._validate_data(X)
        self._validate_params(X)
        Xh = dv.as2d(X)
        self.X = np.array(X)
        self._n = Xh.shape[0]
        self._Xd = dv.to_dev(Xh)
        logger.debug("Computing matrix inverse.")
        self._fac = factorize(Xh, None, self._spec, self.regularization_param)
        self._fac.check()
        C = self._scores_dev(self._Xd)
        self.tau = 1 - float(C.min())
        return self

    def scores(self, T):
        self._validate_data(T)
        return self._scores_dev(dv.to_dev(dv.as2d(T))).cpu().numpy()

    def predict(self, T):
        return self.scores(T) >= 1 - self.tau

    @property
    def W(self):
        return self._fac.full()[: self._n, : self._n].cpu().numpy()

    def score(self):
        raise NotImplementedError
