"""Kernel-based stochastic reachability on the GPU: drop-in for
`gym_socks/algorithms/reach/kernel_sr.py` (`kernel_sr`, `KernelSR`).

fit  = Gram + Cholesky + triangular inverse -> w = diag(W); B = diag(w) K(X,Y) resident in HBM;
       T-1 GEMV passes with the FHT/THT step fused (socks_sr_recursion).
predict = one fused kernel over the test points, K(X,T) never materialised (socks_sr_predict).
"""
import logging

import numpy as np

from socks_b200 import _device as dv
from socks_b200 import _lib
from socks_b200.algorithms.reach.common import PROBLEM_CODE, pack_tubes
from socks_b200.kernel.metrics import factorize, gram_dev, resolve_kernel, sharded_bandwidth_guard
from socks_b200.sampling.transform import transpose_sample

logger = logging.getLogger(__name__)

__all__ = ["kernel_sr", "KernelSR"]


def kernel_sr(S, T, time_horizon=None, constraint_tube=None, target_tube=None, problem="THT",
              regularization_param=None, kernel_fn=None, batch_size=None, verbose=False):
    """One-shot wrapper, `gym_socks/algorithms/reach/kernel_sr.py:116-170`.
    Returns safety probabilities of shape (time_horizon, len(T))."""
    alg = KernelSR(time_horizon=time_horizon, constraint_tube=constraint_tube, target_tube=target_tube,
                   problem=problem, regularization_param=regularization_param, kernel_fn=kernel_fn,
                   batch_size=batch_size, verbose=verbose)
    alg.fit(S)
    return alg.predict(T)


class KernelSR:
    """`gym_socks/algorithms/reach/kernel_sr.py:173-347`: same constructor, `fit(S)`, `predict(T)`,
    attributes `X`, `W`, `value_functions`.  `batch_size` is accepted for API parity and does not change
    results (test points are independent columns; the fused predict kernel streams them anyway)."""

    # 0: automatic (factored form on the fp64 tensor cores), 1: FMA-pipe contraction, 2: direct form on the fp64 tensor
    # cores, 3 / 4: forced CTA shapes, 6: integer tensor-core engine (tcgen05.mma kind::i8; slower, kept as a cross-check)
    # -- csrc/sr_predict.cu
    predict_variant = 0

    def __init__(self, time_horizon=None, constraint_tube=None, target_tube=None, problem="THT", kernel_fn=None,
                 regularization_param=None, batch_size=None, verbose=False, *args, **kwargs):
        self.time_horizon = time_horizon
        self.constraint_tube = constraint_tube
        self.target_tube = target_tube
        self.problem = problem
        self.kernel_fn = kernel_fn
        self.regularization_param = regularization_param
        self.batch_size = batch_size
        self.verbose = verbose

    # -- validation: same order and exception types as kernel_sr.py:224-264
    def _validate_params(self, S):
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

    def _validate_data(self, S):
        if S is None:
            raise ValueError("Must supply a sample.")

    # -- fitted state for the multi-GPU helpers (socks_b200/distributed.py): ~ N (d + 1 + T) doubles
    def state_tensors(self):
        return {"X": self._Xd, "vf": self._vf, "w": self._w}

    def allocate_state(self, n, d):
        """Receiving side of a broadcast: empty device state of the right shapes (no fit on this rank)."""
        self._validate_params(None)
        T = int(self.time_horizon)
        self._n, self._d, self._Xh = int(n), int(d), None
        self._Xd, self._w, self._vf = dv.empty(n, d), dv.empty(n), dv.empty(T, n)
        self._vf_host, self._fac = None, None
        self._tubes = pack_tubes(self.constraint_tube, self.target_tube, T, d)
        return self.state_tensors()

    def fit(self, S):
        self._validate_data(S)
        self._validate_params(S)
        X, U, Y = transpose_sample(S)
        return self.fit_arrays(X, Y)

    def fit_arrays(self, X, Y):
        """fit() on already transposed samples (avoids building the list of tuples at N = 20k+)."""
        if not hasattr(self, "_spec"):
            self._validate_params(None)
        Xh, Yh = dv.as2d(X), dv.as2d(Y)
        n, d = Xh.shape
        T = int(self.time_horizon)
        self._n, self._d = n, d
        self._Xh = Xh
        logger.debug("Computing matrix inverse.")
        fac = factorize(Xh, None, self._spec, self.regularization_param)
        self._fac = fac
        w = fac.diag()
        fac.check()
        if not self._keep_factors:
            fac.release_factors()
        Xd, Yd = dv.to_dev(Xh), dv.to_dev(Yh)
        self._Xd, self._w = Xd, w
        self._tubes = pack_tubes(self.constraint_tube, self.target_tube, T, d)
        lib = _lib.load()
        vf = dv.empty(max(T, 1), n)
        kparams = self._spec.params_for(Xd, Yd)
        fused_pack = self._spec.kind == 0 and not self._spec.data_dependent and T >= 1
        if self._spec.kind == 0 and T >= 1 and dv.stage_marks is None:
            # one enqueue: beta = diag(w) K(X,Y), in-place recursion, predict operand (csrc/compose.cu: socks_sr_fit)
            logger.debug("Computing covariance matrix and value functions.")
            wk = dv.work(lib.socks_sr_fit_work_bytes(n))
            pack = dv.work(lib.socks_sr_pack_bytes(n, d, T)) if fused_pack else None
            _lib.call("socks_sr_fit", dv.ptr(Xd), dv.ptr(Yd), dv.ptr(w), n, d, float(kparams[0]), int(kparams[1]),
                      dv.ptr(self._tubes), PROBLEM_CODE[self.problem], T, dv.ptr(vf), n, dv.ptr(pack), dv.ptr(wk),
                      dv.stream())
            self._vf = vf[:T]
            if fused_pack:
                self._pack = pack
                self._pack_key = (float(kparams[0]), int(kparams[1]), self._vf.data_ptr(), self._w.data_ptr())
            self._vf_host = None
            return self
        # the same kernels enqueued one by one: other kernel families, and when stage timing is on (bench.py)
        logger.debug("Computing covariance matrix.")
        ldb = dv.even(n)
        B = gram_dev(Xd, Yd, kparams, kind=self._spec.kind, extra=self._spec.extra, rowscale=w, ld=ldb)  # un-normalised beta, kernel_sr.py:45
        dv.mark("gram_cross")
        logger.debug("Computing value functions.")
        wk = dv.work(lib.socks_reduce_work_bytes(n) + 8 * (n + 2))
        if T >= 1:
            _lib.call("socks_sr_recursion", dv.ptr(B), n, n, ldb, dv.ptr(Yd), d, dv.ptr(self._tubes),
                      PROBLEM_CODE[self.problem], T, dv.ptr(vf), n, dv.ptr(vf), n, dv.ptr(wk), dv.stream())
        dv.mark("fit_recursion")
        self._vf = vf[:T]
        self._vf_host = None
        del B
        if fused_pack:
            self._packed(kparams)  # predict operand, once per fit
            dv.mark("pack")
        return self

    _keep_factors = True  # `.W` stays available; set False to free L and M right after fit

    # How predict() returns the (T, M) result to the host, measured on B200 (profiles/r2_e2e_small.json):
    #   M < 100 000: ONE launch whose finish kernel stores straight into the pinned host buffer (zero-copy over PCIe);
    #   M >= 100 000: two column blocks, the device-to-host copy of block 1 overlapping the kernel of block 2.
    predict_chunks = None       # None: automatic; an int forces that many column blocks (and disables direct stores)
    predict_direct_host = None  # None: automatic; True / False forces

    @staticmethod
    def _chunks_for(m):
        return 2 if m >= 100000 else 1

    def predict(self, T):
        """Host in, host out (`kernel_sr.py:311-344`): (m, d) points -> (time_horizon, m) float64 ndarray."""
        self._validate_data(T)
        pts_h = dv.as2d(T)
        m, Th = pts_h.shape[0], int(self.time_horizon)
        if m == 0 or Th == 0:
            return np.empty((Th, m))
        import torch

        src = torch.from_numpy(pts_h)
        pts = src.to(dv.device(), non_blocking=src.is_pinned())
        host, finish = dv.host_pool.get(Th, m)
        kparams = self._kparams_for(pts)  # once for the whole call (sigma=None: median over K(X,T))
        direct = self.predict_direct_host
        if direct is None:
            direct = self.predict_chunks is None and m < 100000
        if direct and self._spec.kind == 0:
            # pinned host memory is device-addressable (UVA): the finish kernel's coalesced row stores go over PCIe as
            # posted writes, no copy is enqueued at all
            for s in range(0, m, self.max_launch_points):
                e = min(m, s + self.max_launch_points)
                self._predict_launch(pts[s:e], host.data_ptr() + 8 * s, m, kparams)
            torch.cuda.current_stream().synchronize()
            return finish()
        out = self._cached("out", (Th, m))
        nchunks = self.predict_chunks or self._chunks_for(m)
        chunk = -(-(-(-m // nchunks)) // 256) * 256
        cs = dv.copy_stream()
        for s in range(0, m, chunk):
            e = min(m, s + chunk)
            self._predict_into(pts[s:e], out[:, s:e], kparams)
            ev = self._event(s // chunk)
            ev.record()
            cs.wait_event(ev)
            # Th row segments of this column block -> pinned host, one strided enqueue on the copy stream
            _lib.call("socks_copy2d", host.data_ptr() + 8 * s, 8 * m, out.data_ptr() + 8 * s, 8 * out.stride(0),
                      8 * (e - s), Th, 2, cs.cuda_stream)
        cs.synchronize()
        return finish()

    def _event(self, i):
        import torch

        evs = self.__dict__.setdefault("_events", [])
        while len(evs) <= i:
            evs.append(torch.cuda.Event())
        return evs[i]

    def predict_dev(self, pts, kparams=None):
        """Device in, device out: (m, d) test points -> (time_horizon, m) safety probabilities."""
        m = pts.shape[0]
        Th = int(self.time_horizon)
        out = dv.empty(max(Th, 1), max(m, 1))[:Th, :m]
        if m == 0 or Th == 0:
            return out
        self._predict_into(pts, out, kparams)
        return out

    def _cached(self, name, shape):
        cache = self.__dict__.setdefault("_buf_cache", {})
        t = cache.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = cache[name] = dv.empty(*shape)
        return t

    max_launch_points = 1 << 20  # bounds the per-split partial-sum workspace (16 * splits * 8 B per point)

    def _packed(self, kparams):
        """Packed operand of the predict kernels (socks_sr_pack): a function of the fitted state and the bandwidth,
        so it is built once per fit (once per call only when sigma=None makes the bandwidth data dependent)."""
        key = (float(kparams[0]), int(kparams[1]), self._vf.data_ptr(), self._w.data_ptr())
        if getattr(self, "_pack_key", None) != key:
            Th = int(self.time_horizon)
            self._pack = dv.work(_lib.load().socks_sr_pack_bytes(self._n, self._d, Th))
            _lib.call("socks_sr_pack", dv.ptr(self._Xd), dv.ptr(self._w), dv.ptr(self._vf), self._n, self._n, self._d,
                      key[0], key[1], Th, dv.ptr(self._pack), dv.stream())
            self._pack_key = key
        return self._pack

    def _kparams_for(self, pts):
        """(scale, divide) for K(X, pts).  With sigma=None the reference takes the median over the whole
        len(X) x len(T) distance matrix (metrics.py:130-131), which this path would have to materialise."""
        sharded_bandwidth_guard(self._spec)
        return self._spec.params_for(self._Xd, pts)

    def _predict_into(self, pts, out, kparams=None):
        m = pts.shape[0]
        if kparams is None:
            kparams = self._kparams_for(pts)
        if m > self.max_launch_points:
            for s in range(0, m, self.max_launch_points):
                e = min(m, s + self.max_launch_points)
                self._predict_into(pts[s:e], out[:, s:e], kparams)
            return
        if self._spec.kind != 0:
            return self._predict_materialised(pts, out, kparams)
        self._predict_launch(pts, dv.ptr(out), out.stride(0), kparams)

    def _predict_launch(self, pts, out_ptr, ldout, kparams):
        if self._spec.kind != 0:
            raise NotImplementedError("direct host stores are implemented for the fused RBF predict only")
        m = pts.shape[0]
        scale, divide = kparams
        pack = self._packed(kparams)
        nbytes = _lib.load().socks_sr_predict_packed_work_bytes(self._n, m)
        work = self._cached("work", ((nbytes + 7) // 8 + 2,))
        _lib.call("socks_sr_predict_packed", dv.ptr(pack), dv.ptr(self._Xd), self._n, self._d, float(scale), int(divide),
                  dv.ptr(pts), m, dv.ptr(self._tubes), PROBLEM_CODE[self.problem], int(self.time_horizon), int(out_ptr),
                  int(ldout), dv.ptr(work), int(self.predict_variant), dv.stream())

    def _predict_materialised(self, pts, out, kparams):
        """Kernel families other than RBF (abel, and sklearn's linear / polynomial / laplacian): the fused predict
        kernel evaluates exp(-g |x - p|^2) only, so these take the path the fit uses -- B = diag(w) K(X, chunk) resident
        in HBM, T-1 GEMV passes with the step fused (socks_sr_recursion), chunk by chunk (kernel_sr.py:329-342)."""
        m, n = pts.shape[0], self._n
        lib = _lib.load()
        chunk = max(256, min(m, (1 << 28) // max(n, 1)))  # <= 2 GiB of B per chunk
        for s in range(0, m, chunk):
            p = pts[s:s + chunk]
            cnt = p.shape[0]
            ldb = dv.even(cnt)
            B = gram_dev(self._Xd, p, kparams, kind=self._spec.kind, extra=self._spec.extra, rowscale=self._w, ld=ldb)
            wk = dv.work(lib.socks_reduce_work_bytes(cnt) + 8 * (cnt + 2))
            o = out[:, s:s + cnt]
            _lib.call("socks_sr_recursion", dv.ptr(B), n, cnt, ldb, dv.ptr(p), self._d, dv.ptr(self._tubes),
                      PROBLEM_CODE[self.problem], int(self.time_horizon), dv.ptr(self._vf), n, dv.ptr(o),
                      out.stride(0), dv.ptr(wk), dv.stream())

    # -- attributes the reference exposes (kernel_sr.py:286-309)
    @property
    def X(self):
        if self._Xh is None:
            self._Xh = self._Xd.cpu().numpy()
        return self._Xh

    @property
    def W(self):
        """inv(G) as an ndarray; computed on first access (the algorithm itself only needs diag(W))."""
        if self._fac is None or (self._fac._M is None and self._fac._W is None):
            raise AttributeError("factors were released (keep_factors=False); W is unavailable")
        return self._fac.full()[: self._n, : self._n].cpu().numpy()

    @property
    def value_functions(self):
        if self._vf_host is None:
            self._vf_host = self._vf.cpu().numpy()
        return self._vf_host

    def score(self):
        raise NotImplementedError
