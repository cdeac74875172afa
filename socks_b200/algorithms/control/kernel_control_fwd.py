"""Forward-in-time kernel optimal control on the GPU: drop-in for
`gym_socks/algorithms/control/kernel_control_fwd.py` (`kernel_control_fwd`, `KernelControlFwd`).

Reference per call (:217-238):  beta = W @ (K(X,s) * CUA)  (an N x N x P dgemm),  C = f32(cost) @ beta.
Because W is symmetric the same numbers are  C = (W c)^T (K(X,s) * CUA) = sum_i z_i k(x_i,s) CUA[i,:]
with z = W c: one symmetric GEMV over the resident W, one kernel-evaluation pass and one GEMV over CUA,
all HBM-bound (csrc/sr.cu: socks_gemv_t, socks_fwd_weights), then a device-side masked argmin.
"""
import logging

import numpy as np
import torch

from socks_b200 import _device as dv
from socks_b200 import _lib
from socks_b200.algorithms.control.common import compute_solution
from socks_b200.kernel.metrics import factorize, gram_dev, resolve_kernel
from socks_b200.sampling.transform import transpose_sample

logger = logging.getLogger(__name__)

__all__ = ["kernel_control_fwd", "KernelControlFwd"]


def kernel_control_fwd(S, A, cost_fn=None, constraint_fn=None, heuristic=False, regularization_param=None,
                       kernel_fn=None, verbose=True):
    """`gym_socks/algorithms/control/kernel_control_fwd.py:52-93`: returns the trained policy."""
    alg = KernelControlFwd(cost_fn=cost_fn, constraint_fn=constraint_fn, heuristic=heuristic,
                           regularization_param=regularization_param, kernel_fn=kernel_fn, verbose=verbose)
    alg.train(S, A)
    return alg


class KernelControlFwd:
    """`gym_socks/algorithms/control/kernel_control_fwd.py:95-238`.  Extra keyword arguments are swallowed
    as in the reference (`examples/control/tracking.py:137` passes `time_horizon=`)."""

    def __init__(self, cost_fn=None, constraint_fn=None, heuristic=False, regularization_param=None, kernel_fn=None,
                 verbose=True, *args, **kwargs):
        self.cost_fn = cost_fn
        self.constraint_fn = constraint_fn
        self.heuristic = heuristic
        self.kernel_fn = kernel_fn
        self.regularization_param = regularization_param
        self.verbose = verbose

    def _validate_params(self, S=None, A=None):
        if S is None:
            raise ValueError("Must supply a sample.")
        if A is None:
            raise ValueError("Must supply a sample.")
        self._spec = resolve_kernel(self.kernel_fn, default_sigma=0.1)
        if self.regularization_param is None:
            self.regularization_param = 1
        if self.cost_fn is None:
            raise ValueError("Must supply a cost function.")
        self._constrained = self.constraint_fn is not None

    def _validate_data(self, S):
        if S is None:
            raise ValueError("Must supply a sample.")

    def train(self, S, A):
        self._validate_data(S)
        self._validate_data(A)
        self._validate_params(S=S, A=A)
        X, U, Y = transpose_sample(S)
        return self.train_arrays(X, U, Y, A)

    def train_arrays(self, X, U, Y, A):
        if not hasattr(self, "_spec"):
            self._validate_params(S=True, A=True)
        from functools import partial

        Xh, Uh, Yh, Ah = dv.as2d(X), dv.as2d(U), dv.as2d(Y), dv.as2d(A)
        n, d = Xh.shape
        self._n, self._d, self._P = n, d, Ah.shape[0]
        self._Xh, self.A = Xh, np.array(A)
        logger.debug("Computing matrix inverse.")
        fac = factorize(Xh, Uh, self._spec, self.regularization_param)
        self._fac = fac
        self._Wd = fac.full()  # (np, np), symmetric; kernel_control_fwd.py:192-197
        fac.check()
        fac.release_factors()
        self._np = fac.np
        self._Xd = dv.to_dev(Xh)
        Ud, Ad = dv.to_dev(Uh), dv.to_dev(Ah)
        logger.debug("Computing covariance matrix.")
        self._CUA = gram_dev(Ud, Ad, self._spec.params_for(Ud, Ad), kind=self._spec.kind, extra=self._spec.extra, ld=self._P)  # :200
        Yarr = np.array(Y)
        self.cost_fn = partial(self.cost_fn, state=Yarr)  # :203
        if self.constraint_fn is not None:
            self.constraint_fn = partial(self.constraint_fn, state=Yarr)  # :206-207
        lib = _lib.load()
        self._wk = dv.work(lib.socks_fwd_scores_work_bytes(n, self._P))
        self._idx = torch.zeros(1, dtype=torch.int32, device=self._Xd.device)
        self._idx_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        # per-call inputs travel in ONE pinned buffer -> ONE host-to-device copy: [state (d, padded to 8) | cost | constraint]
        self._ldc = dv.even(n)
        self._stage_h = torch.zeros(8 + 2 * self._ldc, dtype=torch.float64).pin_memory()
        self._stage_d = dv.zeros(8 + 2 * self._ldc)
        self._scores = dv.zeros(2, self._P)
        self._h2d_done = None
        return self

    def _enqueue(self, time, state, policy):
        """Upload (state, float32-rounded cost/constraint rows) and enqueue socks_fwd_scores / socks_fwd_policy
        (csrc/compose.cu): z = W c, q = z * K(X, s), scores = q^T CUA [, masked argmin].  Returns R."""
        n, d, P = self._n, self._d, self._P
        sh = np.asarray(state, dtype=np.float64).reshape(-1)
        if sh.size != d:
            raise ValueError("KernelControlFwd evaluates one state per call (kernel_control_fwd.py:222 broadcasts "
                             "an (N, 1) column against CUA).")
        if self._h2d_done is not None:
            self._h2d_done.synchronize()  # the previous call's copy has left the pinned buffer
        host = self._stage_h.numpy()
        host[:d] = sh
        host[8:8 + n] = np.array(self.cost_fn(time=time), dtype=np.float32)  # float32 rounding, :225
        R = 1
        if self.constraint_fn is not None:
            host[8 + self._ldc:8 + self._ldc + n] = np.array(self.constraint_fn(time=time), dtype=np.float32)  # :231
            R = 2
        self._stage_d.copy_(self._stage_h, non_blocking=True)
        self._h2d_done = torch.cuda.Event()
        self._h2d_done.record()
        scale, divide = self._spec.params_for(self._Xd, self._Xd)
        base = self._stage_d.data_ptr()
        args = (dv.ptr(self._Wd), self._np, dv.ptr(self._Xd), dv.ptr(self._CUA), n, d, P, float(scale), int(divide), base,
                base + 64, self._ldc, R, dv.ptr(self._scores), dv.ptr(self._wk))
        if policy:
            _lib.call("socks_fwd_policy", *args, dv.ptr(self._idx), dv.stream())
        else:
            _lib.call("socks_fwd_scores", *args, dv.stream())
        return R

    def scores(self, time=0, state=None):
        """Device (R, P) tensor: row 0 = cost scores C, row 1 = constraint scores D (if constrained)."""
        if self._spec.data_dependent:
            raise NotImplementedError("KernelControlFwd with sigma=None: the bandwidth of K(X, s) would be the median of "
                                      "one column; pass an explicit sigma")
        R = self._enqueue(time, state, policy=False)
        return self._scores[:R].clone()

    def __call__(self, time=0, state=None, *args, **kwargs):
        if state is None:
            print("Must supply a state to the policy.")
            return None
        if self.heuristic is True:
            self._enqueue(time, state, policy=True)
            self._idx_host.copy_(self._idx, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            idx = int(self._idx_host[0])
        else:
            R = self._enqueue(time, state, policy=False)
            yh = self._scores[:R].cpu().numpy()
            sol = compute_solution(yh[0], yh[1] if R > 1 else None, heuristic=False)
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
