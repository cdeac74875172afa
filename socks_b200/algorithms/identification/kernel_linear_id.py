"""Kernel-based linear system identification on the GPU: drop-in for
`gym_socks/algorithms/identification/kernel_linear_id.py` (`kernel_linear_id`, `KernelLinearId`).

fit: the tall-skinny products Z^T Z and Z^T Y of the primal normal equations (`:133-145`, Z = [X | U | 1]) are one pass
over the samples (socks_atb_small); the (d+m+1)^2 solve (`:148-151`) runs through the padded Cholesky path
(socks_pad_spd -> socks_potrf_inv -> socks_potrs).  predict: socks_linear_map (`:173-181`).
"""
import logging

import numpy as np
import torch

from socks_b200 import _device as dv
from socks_b200 import _lib
from socks_b200.sampling.transform import transpose_sample

logger = logging.getLogger(__name__)

__all__ = ["kernel_linear_id", "KernelLinearId"]


def kernel_linear_id(S, regularization_param=None, verbose=False):
    """`kernel_linear_id.py:22-49`: returns the fitted model."""
    alg = KernelLinearId(regularization_param=regularization_param, verbose=verbose)
    alg.fit(S)
    return alg


class KernelLinearId:
    """`kernel_linear_id.py:52-185`: `fit(S)`, `predict(T, U=None)`, properties `state_matrix`, `input_matrix`,
    `disturbance_mean`; default regularization_param = 1 (`:99-100`)."""

    _estimator_type = "regressor"

    def __init__(self, regularization_param=None, verbose=False, *args, **kwargs):
        self.regularization_param = regularization_param
        self.verbose = verbose
        self._state_matrix = None
        self._input_matrix = None
        self._disturbance_mean = None

    @property
    def state_matrix(self):
        return self._state_matrix

    @property
    def input_matrix(self):
        return self._input_matrix

    @property
    def disturbance_mean(self):
        return self._disturbance_mean

    def _validate_params(self, S):
        if self.regularization_param is None:
            self.regularization_param = 1

    def _validate_data(self, S):
        if S is None:
            raise ValueError("Must supply a sample.")

    def fit(self, S):
        self._validate_data(S)
        self._validate_params(S)
        X, U, Y = transpose_sample(S)
        return self.fit_arrays(X, U, Y)

    def fit_arrays(self, X, U, Y):
        self._validate_params(None)
        Xh, Uh, Yh = dv.as2d(X), dv.as2d(U), dv.as2d(Y)
        n, x_dim = Xh.shape
        u_dim = Uh.shape[1]
        Zh = np.concatenate((Xh, Uh, np.ones((n, 1))), axis=1)  # :136
        z_dim = Zh.shape[1]
        if z_dim > 64 or Yh.shape[1] > 64:
            raise ValueError("KernelLinearId on the GPU supports up to 64 regressor / output columns")
        if n < z_dim:
            logger.warning(f"The sample size is less than the dim of H: {n} < {z_dim}.")
        lib, st = _lib.load(), dv.stream()
        Zd, Yd = dv.to_dev(Zh), dv.to_dev(Yh)
        y_dim = Yh.shape[1]
        wk = dv.work(lib.socks_atb_small_work_bytes(n, z_dim, max(z_dim, y_dim)))
        CZZ = dv.empty(z_dim, z_dim)
        _lib.call("socks_atb_small", dv.ptr(Zd), z_dim, z_dim, dv.ptr(Zd), z_dim, z_dim, n, dv.ptr(CZZ), z_dim, dv.ptr(wk), st)
        npad = dv.padded(z_dim)
        B = dv.zeros(npad, 128)  # right-hand sides C_YZ^T = Z^T Y in the first y_dim columns, zero padding
        _lib.call("socks_atb_small", dv.ptr(Zd), z_dim, z_dim, dv.ptr(Yd), y_dim, y_dim, n, dv.ptr(B), 128, dv.ptr(wk), st)
        # solve(C_ZZ^T, C_YZ^T) (:148-151); C_ZZ = Z^T Z + lambda N I is symmetric positive definite
        G = dv.empty(npad, npad)
        _lib.call("socks_pad_spd", dv.ptr(CZZ), z_dim, z_dim, float(self.regularization_param) * n, dv.ptr(G), npad, st)
        M = dv.empty(npad, npad)
        info = torch.zeros(1, dtype=torch.int32, device=Zd.device)
        w1 = dv.work(lib.socks_trtri_work_bytes(z_dim))
        _lib.call("socks_potrf_inv", dv.ptr(G), z_dim, npad, dv.ptr(M), npad, dv.ptr(w1), dv.ptr(info), 0, st)
        w2 = dv.empty(npad, 128)
        _lib.call("socks_potrs", dv.ptr(M), z_dim, npad, dv.ptr(B), 128, 128, dv.ptr(w2), st)
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError("Z^T Z + lambda N I is not positive definite")
        H = B[:z_dim, :y_dim].t().contiguous()  # (y_dim, z_dim)
        self._H = H
        Hh = H.cpu().numpy()
        self._state_matrix = Hh[:x_dim, :x_dim]  # :153-155
        self._input_matrix = Hh[:x_dim, x_dim:x_dim + u_dim]
        self._disturbance_mean = Hh[:x_dim, -1:]
        self._x_dim, self._u_dim = x_dim, u_dim
        return self

    def predict(self, T, U=None):
        """`:159-183`: state_matrix @ T.T (+ input_matrix @ U.T) + squeeze(disturbance_mean.T), shape (x_dim, len(T)).
        The last addition is numpy's broadcast of an (x_dim,) vector against (x_dim, len(T)) exactly as the reference
        writes it (it adds along the LAST axis, so it only broadcasts for len(T) == x_dim or x_dim == 1); it is applied
        with numpy on the returned array so that shapes which raise in the reference raise here too."""
        self._validate_data(T)
        Th = dv.as2d(np.array(T))
        x_dim, u_dim = self._x_dim, self._u_dim
        m = Th.shape[0]
        Td = dv.to_dev(Th)
        Ud = None
        if U is not None:
            self._validate_data(U)
            Ud = dv.to_dev(dv.as2d(np.array(U)))
        out = dv.empty(x_dim, max(m, 1))
        z = self._H.shape[1]
        _lib.call("socks_linear_map", dv.ptr(self._H), z, x_dim, self._H.data_ptr() + 8 * x_dim, z, u_dim, x_dim,
                  dv.ptr(Td), dv.ptr(Ud), m, dv.ptr(out), out.stride(0), dv.stream())
        result = out[:, :m].cpu().numpy()
        result += np.squeeze(self._disturbance_mean.T)  # :181
        return result

    def score(self):
        raise NotImplementedError
