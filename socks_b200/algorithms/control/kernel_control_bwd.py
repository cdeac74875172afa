"""Backward-in-time (dynamic programming) kernel optimal control on the GPU: drop-in for
`gym_socks/algorithms/control/kernel_control_bwd.py` (`kernel_control_bwd`, `KernelControlBwd`).

Recursion (`:58-122`): out[T-1] = f32(cost_{T-1}); for t = T-2..0: Z = out[t+1] @ W, C = K(X,Y)^T diag(Z) CUA,
optional D from constraint_t @ W, V = row-min of C over {D <= 0}, out[t] = cost_t + V.  The reference's
N x N x P tensor `beta` (`:67`) is never formed: each step is one symmetric GEMV over the resident W
(socks_gemv_t), a pack (socks_scale_pack), one DMMA GEMM (socks_gemm_tn) and a fused masked row-min
(socks_rowmin_masked).  `__call__` (`:412-440`) is KernelControlFwd's with the value function as the cost.
"""
import logging
from functools import partial

import numpy as np
import torch

from socks_b200 import _device as dv
from socks_b200 import _lib
from socks_b200.algorithms.control.common import compute_solution
from socks_b200.kernel.metrics import factorize, gram_dev, resolve_kernel
from socks_b200.sampling.transform import transpose_sample

logger = logging.getLogger(__name__)

__all__ = ["kernel_control_bwd", "KernelControlBwd"]


def kernel_control_bwd(S, A, time_horizon=None, cost_fn=None, constraint_fn=None, heuristic=False,
                       regularization_param=None, kernel_fn=None, batch_size=None, verbose=True):
    """`gym_socks/algorithms/control/kernel_control_bwd.py:225-266`: returns the trained policy."""
    alg = KernelControlBwd(time_horizon=time_horizon, cost_fn=cost_fn, constraint_fn=constraint_fn,
                           heuristic=heuristic, regularization_param=regularization_param, kernel_fn=kernel_fn,
                           batch_size=batch_size, verbose=verbose)
    alg.train(S, A)
    return alg


class KernelControlBwd:
    """`gym_socks/algorithms/control/kernel_control_bwd.py:269-440`."""

    def __init__(self, time_horizon=None, cost_fn=None, constraint_fn=None, heuristic=False,
                 regularization_param=None, kernel_fn=None, batch_size=None, verbose=True, *args, **kwargs):
        self.time_horizon = time_horizon
        self.cost_fn = cost_fn
        self.constraint_fn = constraint_fn
        self.heuristic = heuristic
        self.kernel_fn = kernel_fn
        self.regularization_param = regularization_param
        self.batch_size = batch_size  # accepted for API parity; the GEMM form needs no batching
        self.verbose = verbose

    def _validate_params(self, S=None, A=None):
        self._spec = resolve_kernel(self.kernel_fn, default_sigma=0.1)
        if self.regularization_param is None:
            self.regularization_param = 1
        if S is None:
            raise ValueError("Must supply a sample.")
        if A is None:
            raise ValueError("Must supply a sample.")
        if self.cost_fn is None:
            raise ValueError("Must supply a cost function.")
        self._constrained = self.constraint_fn is not None

    def _validate_data(self, S):
        if S is None:
            raise ValueError("Must supply a sample.")

    def train(self, S, A):
        self._validate_data(S)
        self._validate_data(A)
        self._validate_params(S, A)
        X, U, Y = transpose_sample(S)
        return self.train_arrays(X, U, Y, A)

    def train_arrays(self, X, U, Y, A):
        if not hasattr(self, "_spec"):
            self._validate_params(True, True)
        Xh, Uh, Yh, Ah = dv.as2d(X), dv.as2d(U), dv.as2d(Y), dv.as2d(A)
        n, d = Xh.shape
        P = Ah.shape[0]
        T = int(self.time_horizon)
        self._n, self._d, self._P, self._Xh, self.A = n, d, P, Xh, np.array(A)
        lib = _lib.load()
        st = dv.stream()
        logger.debug("Computing matrix inverse.")
        fac = factorize(Xh, Uh, self._spec, self.regularization_param)
        self._Wd = fac.full()
        fac.check()
        fac.release_factors()
        npad = self._np = fac.np
        Xd, Yd, Ud, Ad = dv.to_dev(Xh), dv.to_dev(Yh), dv.to_dev(Uh), dv.to_dev(Ah)
        self._Xd = Xd
        logger.debug("Computing covariance matrices.")
        self._CUA = gram_dev(Ud, Ad, self._spec.params_for(Ud, Ad), kind=self._spec.kind, extra=self._spec.extra, ld=P)
        Kmat = dv.zeros(npad, npad)  # K(X, Y), the GEMM's (K x m) left operand, zero padded
        gram_dev(Xd, Yd, self._spec.params_for(Xd, Yd), kind=self._spec.kind, extra=self._spec.extra, out=Kmat, ld=npad)
        Yarr = np.array(Y)
        self.cost_fn = partial(self.cost_fn, state=Yarr)  # kernel_control_bwd.py:387
        if self.constraint_fn is not None:
            self.constraint_fn = partial(self.constraint_fn, state=Yarr)  # :390-391
        logger.debug("Computing value functions.")
        R = 2 if self._constrained else 1
        ldbm = -(-R * P // 128) * 128
        Bm, Cm = dv.empty(npad, ldbm), dv.empty(npad, ldbm)
        wk = dv.work(max(lib.socks_reduce_work_bytes(n), lib.socks_reduce_work_bytes(P)))
        vf = dv.empty(max(T, 1), n)
        last = np.array(self.cost_fn(time=T - 1), dtype=np.float32).astype(np.float64)  # :76-78
        vf[T - 1].copy_(dv.to_dev(last))
        rows, z = dv.empty(R, n), dv.empty(R, n)
        for t in range(T - 2, -1, -1):
            rows[0].copy_(vf[t + 1])
            if self._constrained:
                rows[1].copy_(dv.to_dev(np.asarray(self.constraint_fn(time=t), dtype=np.float64)))  # :96
            # Z = rows @ W (W symmetric), :90,96
            _lib.call("socks_gemv_t", dv.ptr(self._Wd), n, n, npad, dv.ptr(rows), n, R, dv.ptr(z), n, dv.ptr(wk), st)
            _lib.call("socks_scale_pack", dv.ptr(self._CUA), n, P, dv.ptr(z), n, R, dv.ptr(Bm), npad, ldbm, st)
            _lib.call("socks_gemm_tn", dv.ptr(Kmat), npad, dv.ptr(Bm), ldbm, dv.ptr(Cm), ldbm, npad, ldbm, npad, st)
            cost_t = dv.to_dev(np.asarray(self.cost_fn(time=t), dtype=np.float64))
            _lib.call("socks_rowmin_masked", dv.ptr(Cm), ldbm, n, P, int(self._constrained), dv.ptr(cost_t),
                      dv.ptr(vf[t]), st)  # :99-115
        self._vf = vf[:T]
        self._vf_host = None
        self._wk = wk
        self._idx = torch.zeros(1, dtype=torch.int32, device=Xd.device)
        return self

    def scores(self, time=0, state=None):
        """Device (R, P): row 0 = C (value-function scores), row 1 = D (constraint scores)."""
        n, d, P = self._n, self._d, self._P
        st = dv.stream()
        sd = dv.to_dev(np.asarray(state, dtype=np.float64).reshape(-1))
        if sd.numel() != d:
            raise ValueError("KernelControlBwd evaluates one state per call (kernel_control_bwd.py:424).")
        rows = [np.array(self.value_functions[time], dtype=np.float32)]  # float32 rounding, :427
        if self.constraint_fn is not None:
            rows.append(np.array(self.constraint_fn(time=time), dtype=np.float32))  # :432-434
        R = len(rows)
        c = dv.to_dev(np.stack(rows).astype(np.float64))
        z, q, y = dv.empty(R, n), dv.empty(R, n), dv.empty(R, P)
        _lib.call("socks_gemv_t", dv.ptr(self._Wd), n, n, self._np, dv.ptr(c), n, R, dv.ptr(z), n, dv.ptr(self._wk), st)
        scale, divide = self._spec.params_for(self._Xd, sd.reshape(1, d))
        _lib.call("socks_fwd_weights", dv.ptr(self._Xd), n, d, float(scale), int(divide), dv.ptr(sd), dv.ptr(z), n, R,
                  dv.ptr(q), n, st)
        _lib.call("socks_gemv_t", dv.ptr(self._CUA), n, P, P, dv.ptr(q), n, R, dv.ptr(y), P, dv.ptr(self._wk), st)
        return y

    def __call__(self, time=0, state=None, *args, **kwargs):
        if state is None:
            print("Must supply a state to the policy.")
            return None
        y = self.scores(time=time, state=state)
        if self.heuristic is True:
            Dp = dv.ptr(y[1]) if y.shape[0] > 1 else 0
            _lib.call("socks_argmin_masked", dv.ptr(y[0]), Dp, self._P, dv.ptr(self._idx), dv.stream())
            idx = int(self._idx.item())
        else:
            yh = y.cpu().numpy()
            sol = compute_solution(yh[0], yh[1] if yh.shape[0] > 1 else None, heuristic=False)
            idx = int(np.argmax(sol))
        return np.array(self.A[idx], dtype=np.float32)

    @property
    def X(self):
        return self._Xh

    @property
    def W(self):
        return self._Wd[: self._n, : self._n].cpu().numpy()

    @property
    def CUA(self):
        return self._CUA.cpu().numpy()

    @property
    def value_functions(self):
        if self._vf_host is None:
            self._vf_host = self._vf.cpu().numpy()
        return self._vf_host
