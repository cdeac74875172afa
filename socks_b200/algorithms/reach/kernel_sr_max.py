"""Maximal kernel-based stochastic reachability on the GPU: drop-in for
`gym_socks/algorithms/reach/kernel_sr_max.py` (`kernel_sr_max`, `KernelMaximalSR`).

The reference's N x M x P tensor (kernel_sr_max.py:45-47) is never formed: each recursion step is a
DMMA GEMM  K(X,pts)^T [diag(w) CUA | diag(vf w) CUA]  followed by a fused divide + max + step
(csrc/srmax.cu).  In predict all T-1 steps share one GEMM per chunk of test points.
"""
import logging

import numpy as np

from socks_b200 import _device as dv
from socks_b200 import _lib
from socks_b200.algorithms.reach.common import PROBLEM_CODE, pack_tubes
from socks_b200.kernel.metrics import factorize, gram_dev, resolve_kernel, sharded_bandwidth_guard
from socks_b200.sampling.transform import transpose_sample

logger = logging.getLogger(__name__)

__all__ = ["kernel_sr_max", "KernelMaximalSR"]


def kernel_sr_max(S, A, T, time_horizon=None, constraint_tube=None, target_tube=None, problem="THT",
                  regularization_param=None, kernel_fn=None, batch_size=None, verbose=False):
    """One-shot wrapper, `gym_socks/algorithms/reach/kernel_sr_max.py:142-199`."""
    alg = KernelMaximalSR(time_horizon=time_horizon, constraint_tube=constraint_tube, target_tube=target_tube,
                          problem=problem, regularization_param=regularization_param, kernel_fn=kernel_fn,
                          batch_size=batch_size, verbose=verbose)
    alg.fit(S, A)
    return alg.predict(T)


class KernelMaximalSR:
    """`gym_socks/algorithms/reach/kernel_sr_max.py:202-387`.  Note the keyword order of the reference:
    `regularization_param` precedes `kernel_fn` here (:231-232)."""

    predict_chunk = 8192  # test points per GEMM (multiple of 128)
    # Contraction engine of predict: "i8" = Ozaki splitting on the integer tensor cores (tcgen05.mma kind::i8, csrc/ozaki.cu;
    # exact digit products, num/den good to ~1e-11), with the FP64 DMMA GEMM for the points it flags; "f64" = DMMA only.
    predict_engine = "i8"
    predict_slices = 7   # 49-bit fixed point: |num, den error| ~ 4e-12 of the operand scale at N = 20 000 (measured)
    predict_tau = 1e-2   # points with min_a den[a] < tau * scale go to the FP64 path (relative error there > ~4e-10)

    def __init__(self, time_horizon=None, constraint_tube=None, target_tube=None, problem="THT",
                 regularization_param=None, kernel_fn=None, batch_size=None, verbose=False, *args, **kwargs):
        self.time_horizon = time_horizon
        self.constraint_tube = constraint_tube
        self.target_tube = target_tube
        self.problem = problem
        self.kernel_fn = kernel_fn
        self.regularization_param = regularization_param
        self.batch_size = batch_size
        self.verbose = verbose

    def _validate_params(self, S=None, A=None):
        self._spec = resolve_kernel(self.kernel_fn, default_sigma=0.1)
        if self.regularization_param is None:
            self.regularization_param = 1
        if self.time_horizon is None:
            raise ValueError("Must supply a time horizon.")
        assert self.time_horizon >= 0 and isinstance(self.time_horizon, (int, np.integer))
        if self.constraint_tube is None:
            raise ValueError("Must supply constraint tube.")
        if self.target_tube is None:
            raise ValueError("Must supply target tube.")
        if self.problem not in ("FHT", "THT"):
            raise ValueError(f"problem is not in {'THT', 'FHT'}")

    def _validate_data(self, S=None):
        if S is None:
            raise ValueError("Must supply a sample.")

    # -- fitted state for the multi-GPU helpers (socks_b200/distributed.py): N (d + 1 + T + P) doubles
    def state_tensors(self):
        return {"CUA": self._CUA, "X": self._Xd, "vf": self._vf, "w": self._w}

    def allocate_state(self, n, d, P):
        """Receiving side of a broadcast: empty device state of the right shapes (no fit on this rank)."""
        self._validate_params()
        T = int(self.time_horizon)
        self._n, self._d, self._P, self._Xh, self._fac = int(n), int(d), int(P), None, None
        self._Xd, self._w, self._vf, self._CUA = dv.empty(n, d), dv.empty(n), dv.empty(T, n), dv.empty(n, P)
        self._vf_host = None
        self._tubes = pack_tubes(self.constraint_tube, self.target_tube, T, d)
        return self.state_tensors()

    @staticmethod
    def _kmul_negative(kparams):
        scale, divide = kparams
        return (1.0 / scale if divide else scale) < 0.0

    def fit(self, S, A):
        self._validate_params(S, A)
        self._validate_data(S)
        self._validate_data(A)
        X, U, Y = transpose_sample(S)
        return self.fit_arrays(X, U, Y, A)

    def fit_arrays(self, X, U, Y, A):
        if not hasattr(self, "_spec"):
            self._validate_params()
        Xh, Uh, Yh, Ah = dv.as2d(X), dv.as2d(U), dv.as2d(Y), dv.as2d(A)
        n, d = Xh.shape
        P = Ah.shape[0]
        T = int(self.time_horizon)
        self._n, self._d, self._P, self._Xh = n, d, P, Xh
        npad = dv.padded(n)
        logger.debug("Computing matrix inverse.")
        fac = factorize(Xh, Uh, self._spec, self.regularization_param)
        self._fac = fac
        w = fac.diag()
        fac.check()
        Xd, Yd, Ud, Ad = dv.to_dev(Xh), dv.to_dev(Yh), dv.to_dev(Uh), dv.to_dev(Ah)
        self._Xd, self._w = Xd, w
        self._tubes = pack_tubes(self.constraint_tube, self.target_tube, T, d)
        lib = _lib.load()
        st = dv.stream()
        prob = PROBLEM_CODE[self.problem]
        vf = dv.empty(max(T, 1), n)
        kp_xy, kp_ua = self._spec.params_for(Xd, Yd), self._spec.params_for(Ud, Ad)
        if self._spec.kind == 0 and kp_xy == kp_ua and T >= 1:
            # one enqueue (csrc/compose.cu: socks_srmax_fit): CUA, K(X,Y), normaliser GEMM, T-1 x (pack, GEMM, max/step)
            logger.debug("Computing covariance matrices and value functions.")
            CUA = dv.empty(n, P)
            if self.predict_engine == "i8" and d <= 8 and self._kmul_negative(kp_xy):
                # the backward recursion on the integer tensor cores; samples outside the fixed-point range (counted
                # on the device) send the whole fit to the FP64 engine below
                import torch

                wk = dv.work(lib.socks_srmax_fit_i8_work_bytes(n, P, int(self.predict_slices)) + 1024)
                base = (wk.data_ptr() + 1023) // 1024 * 1024
                flagged = torch.zeros(1, dtype=torch.int32, device=Xd.device)
                _lib.call("socks_srmax_fit_i8", dv.ptr(Xd), dv.ptr(Yd), dv.ptr(Ud), dv.ptr(Ad), dv.ptr(w), n, d, Ud.shape[1],
                          P, float(kp_xy[0]), int(kp_xy[1]), dv.ptr(self._tubes), prob, T, dv.ptr(CUA), dv.ptr(vf), n, base,
                          int(self.predict_slices), float(self.predict_tau), dv.ptr(flagged), st)
                self.last_fit_flagged = int(flagged.item())
                if self.last_fit_flagged == 0:
                    self._CUA = CUA
                    self._vf = vf[:T]
                    self._vf_host = None
                    return self
            wk = dv.work(lib.socks_srmax_fit_work_bytes(n, P))
            _lib.call("socks_srmax_fit", dv.ptr(Xd), dv.ptr(Yd), dv.ptr(Ud), dv.ptr(Ad), dv.ptr(w), n, d, Ud.shape[1], P,
                      float(kp_xy[0]), int(kp_xy[1]), dv.ptr(self._tubes), prob, T, dv.ptr(CUA), dv.ptr(vf), n,
                      dv.ptr(wk), st)
            self._CUA = CUA
            self._vf = vf[:T]
            self._vf_host = None
            return self
        logger.debug("Computing covariance matrices.")
        CUA = gram_dev(Ud, Ad, kp_ua, kind=self._spec.kind, extra=self._spec.extra, ld=P)  # (n, P) contiguous, kernel_sr_max.py:332
        self._CUA = CUA
        # K(X, Y) as the GEMM's left operand: (K = npad rows) x (m = npad columns), zero padding
        Kmat = dv.zeros(npad, npad)
        gram_dev(Xd, Yd, kp_xy, kind=self._spec.kind, extra=self._spec.extra, out=Kmat, ld=npad)
        logger.debug("Computing value functions.")
        ldbm = -(-P // 128) * 128
        Bm = dv.empty(npad, ldbm)
        Cm = dv.empty(npad, ldbm)
        Den = dv.empty(npad, ldbm)
        # vf[T-1] = 1_target(Y) first (kernel_sr_max.py:59): the first pack below reads it
        _lib.call("socks_srmax_step", dv.ptr(Cm), ldbm, 0, 0, n, P, 1, 0, dv.ptr(Yd), d, dv.ptr(self._tubes), prob, T,
                  dv.ptr(vf), n, 1, st)
        if T > 1:  # the normaliser sum_i w_i C_ij CUA_ik does not depend on t: one GEMM for all steps
            _lib.call("socks_srmax_pack", dv.ptr(CUA), n, P, dv.ptr(w), 0, n, 0, 0, 1, dv.ptr(Bm), npad, ldbm, st)
            _lib.call("socks_gemm_tn", dv.ptr(Kmat), npad, dv.ptr(Bm), ldbm, dv.ptr(Den), ldbm, npad, ldbm, npad, st)
        for t in range(T - 2, -1, -1):
            # numerator row: vf[t+1] * w                                   (kernel_sr_max.py:65-72)
            _lib.call("socks_srmax_pack", dv.ptr(CUA), n, P, dv.ptr(w), dv.ptr(vf), n, t, 1, 1, dv.ptr(Bm), npad, ldbm, st)
            _lib.call("socks_gemm_tn", dv.ptr(Kmat), npad, dv.ptr(Bm), ldbm, dv.ptr(Cm), ldbm, npad, ldbm, npad, st)
            _lib.call("socks_srmax_step", dv.ptr(Cm), ldbm, dv.ptr(Den), ldbm, n, P, 2, t, dv.ptr(Yd), d,
                      dv.ptr(self._tubes), prob, T, dv.ptr(vf), n, 0, st)
        self._vf = vf[:T]
        self._vf_host = None
        return self

    def predict(self, T):
        self._validate_data(T)
        return self.predict_dev(dv.to_dev(dv.as2d(T))).cpu().numpy()

    def predict_dev(self, pts, kparams=None):
        m = pts.shape[0]
        Th = int(self.time_horizon)
        n, d, P = self._n, self._d, self._P
        npad = dv.padded(n)
        out = dv.empty(max(Th, 1), max(m, 1))[:Th, :m]
        if m == 0 or Th == 0:
            return out
        st = dv.stream()
        prob = PROBLEM_CODE[self.problem]
        if kparams is None:
            sharded_bandwidth_guard(self._spec)
            kparams = self._spec.params_for(self._Xd, pts)
        if self._spec.kind == 0 and self.predict_engine == "i8" and d <= 8 and Th >= 2 and kparams[0] < 0:
            return self._predict_i8(pts, out, kparams)
        if self._spec.kind == 0:  # one enqueue for the whole test set (csrc/compose.cu: socks_srmax_predict)
            chunk = min(self.predict_chunk, -(-m // 128) * 128)
            wk = self._cached_work(_lib.load().socks_srmax_predict_work_bytes(n, P, Th, chunk))
            _lib.call("socks_srmax_predict", dv.ptr(self._Xd), dv.ptr(self._w), dv.ptr(self._vf), n, dv.ptr(self._CUA), n, d,
                      P, float(kparams[0]), int(kparams[1]), dv.ptr(pts), m, dv.ptr(self._tubes), prob, Th, dv.ptr(out),
                      out.stride(0), dv.ptr(wk), chunk, st)
            return out
        # other kernel families: the same steps with the family Gram builder
        R = Th  # all rows at once: r = 0 normaliser, r = 1..T-1 time steps 0..T-2
        ldbm = -(-R * P // 128) * 128
        Bm = dv.empty(npad, ldbm)
        _lib.call("socks_srmax_pack", dv.ptr(self._CUA), n, P, dv.ptr(self._w), dv.ptr(self._vf), n, 0, 0, R, dv.ptr(Bm),
                  npad, ldbm, st)
        mc = min(self.predict_chunk, -(-m // 128) * 128)
        Kmat = dv.zeros(npad, mc)
        Cm = dv.empty(mc, ldbm)
        for s in range(0, m, mc):
            cnt = min(mc, m - s)
            if cnt < mc:
                Kmat.zero_()
            p = pts[s:s + cnt]
            gram_dev(self._Xd, p, kparams, out=Kmat, ld=mc, kind=self._spec.kind, extra=self._spec.extra)
            _lib.call("socks_gemm_tn", dv.ptr(Kmat), mc, dv.ptr(Bm), ldbm, dv.ptr(Cm), ldbm, mc, ldbm, npad, st)
            _lib.call("socks_srmax_step", dv.ptr(Cm), ldbm, 0, 0, cnt, P, R, 0, dv.ptr(p), d, dv.ptr(self._tubes), prob,
                      Th, dv.ptr(out[:, s:]), out.stride(0), 1, st)
        return out

    def _predict_i8(self, pts, out, kparams):
        """Integer tensor-core contraction (socks_srmax_predict_i8) + FP64 recomputation of the points it flags."""
        import torch

        m, Th = pts.shape[0], int(self.time_horizon)
        n, d, P = self._n, self._d, self._P
        lib, st, prob = _lib.load(), dv.stream(), PROBLEM_CODE[self.problem]
        chunk = min(self.predict_chunk, -(-m // 128) * 128)
        S = int(self.predict_slices)
        nbytes = lib.socks_srmax_predict_i8_work_bytes(n, P, Th, chunk, S)
        wk = self._cached_work(nbytes + 2048)
        base = (wk.data_ptr() + 1023) // 1024 * 1024
        flagged = torch.empty(m + 1, dtype=torch.int32, device=pts.device)
        _lib.call("socks_srmax_predict_i8", dv.ptr(self._Xd), dv.ptr(self._w), dv.ptr(self._vf), n, dv.ptr(self._CUA), n, d,
                  P, float(kparams[0]), int(kparams[1]), dv.ptr(pts), m, dv.ptr(self._tubes), prob, Th, dv.ptr(out),
                  out.stride(0), base, chunk, S, float(self.predict_tau), dv.ptr(flagged), st)
        cnt = int(flagged[0].item())
        self.last_flagged = cnt
        if cnt:  # far-away points: the exact FP64 path, scattered back into their columns
            idx = flagged[1:1 + cnt].long()
            sub = pts.index_select(0, idx).contiguous()
            tmp = dv.empty(Th, cnt)
            c2 = min(self.predict_chunk, -(-cnt // 128) * 128)
            w2 = dv.work(lib.socks_srmax_predict_work_bytes(n, P, Th, c2))
            _lib.call("socks_srmax_predict", dv.ptr(self._Xd), dv.ptr(self._w), dv.ptr(self._vf), n, dv.ptr(self._CUA), n, d,
                      P, float(kparams[0]), int(kparams[1]), dv.ptr(sub), cnt, dv.ptr(self._tubes), prob, Th, dv.ptr(tmp),
                      tmp.stride(0), dv.ptr(w2), c2, st)
            out.index_copy_(1, idx, tmp)
        return out

    def _cached_work(self, nbytes):
        need = (int(nbytes) + 7) // 8 + 2
        t = self.__dict__.get("_work")
        if t is None or t.numel() < need:
            t = self.__dict__["_work"] = dv.work(nbytes)
        return t

    @property
    def X(self):
        if self._Xh is None:
            self._Xh = self._Xd.cpu().numpy()
        return self._Xh

    @property
    def W(self):
        if self._fac is None:
            raise AttributeError("this rank received the fitted state by broadcast; W lives on the fitting rank")
        return self._fac.full()[: self._n, : self._n].cpu().numpy()

    @property
    def CUA(self):
        return self._CUA.cpu().numpy()

    @property
    def value_functions(self):
        if self._vf_host is None:
            self._vf_host = self._vf.cpu().numpy()
        return self._vf_host

    def score(self):
        raise NotImplementedError
