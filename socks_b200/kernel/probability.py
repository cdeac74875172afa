"""`gym_socks.kernel.probability` drop-in for the sample statistics: maximum mean discrepancy and witness function
(`gym_socks/kernel/probability.py:9-93`), SURVEY §8(f) rank 3.  Gram matrices come from the CUDA kernels and are
reduced on the device with the column-sum kernel (socks_gemv_t with V = 1); only m + n numbers reach the host."""
from functools import partial

import numpy as np

from socks_b200 import _device as dv
from socks_b200 import _lib
from socks_b200.kernel.metrics import gram_dev, resolve_kernel

__all__ = ["maximum_mean_discrepancy", "witness_function"]


def _colsums(Xd, Yd, spec):
    """Column sums of k(X, Y): device vector of len(Y)."""
    n, m = Xd.shape[0], Yd.shape[0]
    K = gram_dev(Xd, Yd, spec.params_for(Xd, Yd), kind=spec.kind, extra=spec.extra)
    y = dv.empty(1, m)
    wk = dv.work(_lib.load().socks_reduce_work_bytes(m))
    _lib.call("socks_gemv_t", dv.ptr(K), n, m, K.stride(0), 0, 0, 1, dv.ptr(y), m, dv.ptr(wk), dv.stream())
    return y[0]


def maximum_mean_discrepancy(X, Y, kernel_fn=None, biased=False, squared=False):
    """MMD between two samples, `gym_socks/kernel/probability.py:9-73` (default kernel rbf sigma=1 `:49-50`;
    unbiased estimator drops the diagonals `:63-69`; returns the square root unless `squared`)."""
    spec = resolve_kernel(kernel_fn, default_sigma=1)
    Xh, Yh = dv.as2d(X), dv.as2d(Y)
    m, n = len(Xh), len(Yh)
    Xd, Yd = dv.to_dev(Xh), dv.to_dev(Yh)
    A = float(_colsums(Xd, Xd, spec).cpu().numpy().sum())
    B = float(_colsums(Yd, Yd, spec).cpu().numpy().sum())
    C = float(_colsums(Xd, Yd, spec).cpu().numpy().sum())
    if biased is True:
        mmd = A / (m ** 2) + B / (n ** 2) - 2 * C / (m * n)
    else:  # k(x, x) = 1 on the diagonal of either kernel family
        mmd = (A - m) / (m * (m - 1)) + (B - n) / (n * (n - 1)) - 2 * C / (m * n)
    return mmd if squared else np.sqrt(mmd)


def witness_function(X, Y, t, kernel_fn=None):
    """Witness function mean_i k(x_i, t) - mean_j k(y_j, t), `gym_socks/kernel/probability.py:76-93`."""
    spec = resolve_kernel(kernel_fn, default_sigma=1)
    Xd, Yd, Td = dv.to_dev(dv.as2d(X)), dv.to_dev(dv.as2d(Y)), dv.to_dev(dv.as2d(t))
    A = _colsums(Xd, Td, spec).cpu().numpy()
    B = _colsums(Yd, Td, spec).cpu().numpy()
    return A / Xd.shape[0] - B / Yd.shape[0]
