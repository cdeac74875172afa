"""CPU oracle: numpy restatement of the reference's kernel-embedding hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; nothing under
`socks_b200/` does, and the product path raises when its CUDA library is missing.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference
(`/root/reference/gym_socks`, imported with the test-only `oracle/gym_shim`) on seeded float64
inputs and commits its outputs to `tests/golden/*.npz`; `tests/test_oracle.py` checks every
function below against those vectors and against the reference's own known-answer tests
(`gym_socks/kernel/tests/test_kernel.py:46-130`, `gym_socks/utils/tests/test_utils.py:10-54`).

Each function cites the reference lines it restates.  The restatement keeps the reference's
arithmetic (direct-difference squared distance, explicit LU inverse, diag(W) + column
normalisation) but never materialises the N x N x P tensor of `kernel_sr_max.py:45-47`, and it
evaluates large cross-Grams in column chunks so the sizes of BASELINE.json fit in host RAM.
"""

import numpy as np

__all__ = [
    "rbf_kernel",
    "regularized_inverse",
    "normalize",
    "indicator",
    "fht_step",
    "tht_step",
    "sr_recursion",
    "KernelSROracle",
    "KernelMaximalSROracle",
    "KernelControlFwdOracle",
    "KernelControlBwdOracle",
    "heuristic_solution",
    "abel_kernel",
    "SeparatingKernelOracle",
    "maximum_mean_discrepancy",
    "witness_function",
    "conditional_embedding_predict",
]


# --------------------------------------------------------------------------- kernel/metrics.py
def sq_dist(X, Y):
    """Pairwise squared distance by direct differences, accumulated dimension by dimension
    (`gym_socks/kernel/metrics.py:122-128`: K += (Y_i - X_i)^2 for i in range(d))."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    D = np.zeros((X.shape[0], Y.shape[0]))
    for i in range(X.shape[1]):
        diff = Y[None, :, i] - X[:, None, i]
        D += diff * diff
    return D


def rbf_kernel(X, Y=None, sigma=None):
    """exp(-|x - y|^2 / (2 sigma^2)); sigma=None uses the median of the SQUARED distances
    (`gym_socks/kernel/metrics.py:119-138`; median heuristic `:130-131`)."""
    if Y is None:
        Y = X
    K = sq_dist(X, Y)
    if sigma is None:
        sigma = np.median(K)
    else:
        assert sigma > 0, "sigma must be a strictly positive real value."
    K /= -2 * (sigma ** 2)
    np.exp(K, K)
    return K


def regularized_inverse(X, U=None, regularization_param=None, sigma=1.0):
    """inv(K(X,X) [* K(U,U)] + lambda * N * I), explicit LU inverse
    (`gym_socks/kernel/metrics.py:248-280`; default lambda = 1/N^2 `:251-252`,
    default kernel sigma = 1 `:258-259`, product Gram `:268-276`)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    if regularization_param is None:
        regularization_param = 1 / (n ** 2)
    else:
        assert regularization_param > 0
    K = rbf_kernel(X, X, sigma)
    if U is not None:
        U = np.asarray(U, dtype=np.float64)
        assert U.shape[0] == n
        K = rbf_kernel(U, U, sigma) * K
    K[np.diag_indices(n)] += regularization_param * n
    return np.linalg.inv(K)


# --------------------------------------------------------------------------- utils/__init__.py
def normalize(v):
    """Column L1 normalisation, no guard: 0/0 -> NaN (`gym_socks/utils/__init__.py:22`)."""
    return v / np.sum(v, axis=0)


def indicator(points, box):
    """Closed-box membership against the box's (float32) bounds
    (`gym_socks/utils/__init__.py:42-47`)."""
    lo, hi = box.low, box.high
    return np.array(np.all(points >= lo, axis=1) & np.all(points <= hi, axis=1), dtype=bool)


# --------------------------------------------------------------------------- reach/common.py
def fht_step(pts, V, constraint, target):
    """1_T + 1_{K \\ T} * V (`gym_socks/algorithms/reach/common.py:36-39`)."""
    in_k = indicator(pts, constraint)
    in_t = indicator(pts, target)
    return in_t + (in_k & ~in_t) * V


def tht_step(pts, V, constraint, target):
    """1_K * V (`gym_socks/algorithms/reach/common.py:67-69`)."""
    return indicator(pts, constraint) * V


_STEP = {"FHT": fht_step, "THT": tht_step}


# --------------------------------------------------------------------------- reach/kernel_sr.py
def sr_recursion(pts, w_diag, C, time_horizon, constraint_tube, target_tube, value_functions, problem, out=None):
    """Backward recursion of `gym_socks/algorithms/reach/kernel_sr.py:43-69`.

    beta = normalize(diag(W)[:, None] * C) (`:45`), out[T-1] = 1_target(pts) (`:57`),
    out[t] = step(pts, vf[t+1] @ beta, K_t, T_t) (`:61-63`).  When `out is value_functions`
    (fit) the recursion feeds on its own output exactly as the reference does (`:297-309`)."""
    step = _STEP[problem]
    beta = normalize(w_diag[:, None] * C)
    if out is None:
        out = np.empty((time_horizon, len(pts)))
    out[time_horizon - 1, :] = indicator(pts, target_tube[time_horizon - 1])
    for t in range(time_horizon - 2, -1, -1):
        V = value_functions[t + 1, :] @ beta
        out[t, :] = step(pts, V, constraint_tube[t], target_tube[t])
    return out


class KernelSROracle:
    """`KernelSR.fit/predict` (`gym_socks/algorithms/reach/kernel_sr.py:266-344`) with the
    rbf kernel; defaults sigma=0.1 (`:227`), lambda=1 (`:230`)."""

    def __init__(self, time_horizon, constraint_tube, target_tube, problem="THT", sigma=0.1,
                 regularization_param=1, chunk=4096):
        if problem not in ("FHT", "THT"):
            raise ValueError("problem is not in {'THT', 'FHT'}")
        self.T = time_horizon
        self.ctube, self.ttube = constraint_tube, target_tube
        self.problem, self.sigma, self.lam, self.chunk = problem, sigma, regularization_param, chunk

    def fit(self, X, U, Y):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        self.X = X
        self.W = regularized_inverse(X, regularization_param=self.lam, sigma=self.sigma)
        self.w_diag = np.diag(self.W).copy()
        CXY = rbf_kernel(X, Y, self.sigma)
        vf = np.empty((self.T, len(Y)))
        self.value_functions = sr_recursion(Y, self.w_diag, CXY, self.T, self.ctube, self.ttube, vf,
                                            self.problem, out=vf)
        return self

    def predict(self, pts):
        pts = np.asarray(pts, dtype=np.float64)
        out = np.empty((self.T, len(pts)))
        for s in range(0, len(pts), self.chunk):  # columns are independent (SURVEY Appx B.10)
            p = pts[s:s + self.chunk]
            C = rbf_kernel(self.X, p, self.sigma)
            out[:, s:s + self.chunk] = sr_recursion(p, self.w_diag, C, self.T, self.ctube, self.ttube,
                                                    self.value_functions, self.problem)
        return out


# --------------------------------------------------------------------------- reach/kernel_sr_max.py
def srmax_recursion(pts, w_diag, C, CUA, time_horizon, constraint_tube, target_tube, value_functions,
                    problem, out=None):
    """Backward recursion of `gym_socks/algorithms/reach/kernel_sr_max.py:43-79`.

    The reference forms beta[i,j,k] = W_ii C_ij CUA_ik / sum_i(...) (`:45-47`) and
    wX[j,k] = sum_i vf[t+1,i] beta[i,j,k] (`:65-70`), VX = max_k wX (`:72`).  The same
    numbers as two contractions: den = C^T diag(w) CUA, num_t = C^T diag(vf[t+1] * w) CUA,
    wX = num_t / den."""
    step = _STEP[problem]
    den = C.T @ (w_diag[:, None] * CUA)
    if out is None:
        out = np.zeros((time_horizon, len(pts)))
    out[time_horizon - 1, :] = indicator(pts, target_tube[time_horizon - 1])
    for t in range(time_horizon - 2, -1, -1):
        num = C.T @ ((value_functions[t + 1, :] * w_diag)[:, None] * CUA)
        VX = np.max(num / den, axis=1)
        out[t, :] = step(pts, VX, constraint_tube[t], target_tube[t])
    return out


class KernelMaximalSROracle:
    """`KernelMaximalSR.fit/predict` (`gym_socks/algorithms/reach/kernel_sr_max.py:295-384`):
    product-Gram inverse (`:323-328`), CUA = K(U, A) (`:332`)."""

    def __init__(self, time_horizon, constraint_tube, target_tube, problem="THT", sigma=0.1,
                 regularization_param=1, chunk=2048):
        if problem not in ("FHT", "THT"):
            raise ValueError("problem is not in {'THT', 'FHT'}")
        self.T = time_horizon
        self.ctube, self.ttube = constraint_tube, target_tube
        self.problem, self.sigma, self.lam, self.chunk = problem, sigma, regularization_param, chunk

    def fit(self, X, U, Y, A):
        X = np.asarray(X, dtype=np.float64)
        U = np.asarray(U, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        A = np.asarray(A, dtype=np.float64)
        self.X = X
        self.W = regularized_inverse(X, U=U, regularization_param=self.lam, sigma=self.sigma)
        self.w_diag = np.diag(self.W).copy()
        CXY = rbf_kernel(X, Y, self.sigma)
        self.CUA = rbf_kernel(U, A, self.sigma)
        vf = np.empty((self.T, len(Y)))
        self.value_functions = srmax_recursion(Y, self.w_diag, CXY, self.CUA, self.T, self.ctube,
                                               self.ttube, vf, self.problem, out=vf)
        return self

    def predict(self, pts):
        pts = np.asarray(pts, dtype=np.float64)
        out = np.empty((self.T, len(pts)))
        for s in range(0, len(pts), self.chunk):
            p = pts[s:s + self.chunk]
            C = rbf_kernel(self.X, p, self.sigma)
            out[:, s:s + self.chunk] = srmax_recursion(p, self.w_diag, C, self.CUA, self.T, self.ctube,
                                                       self.ttube, self.value_functions, self.problem)
        return out


# --------------------------------------------------------------------------- control/
def heuristic_solution(C, D=None):
    """Index chosen by the heuristic branches of `compute_solution`
    (`gym_socks/algorithms/control/common.py:96-100` unconstrained argmin;
    `:166-174` argmin over {D <= 0}, global argmin if that set is empty)."""
    if D is None:
        return int(np.argmin(C))
    ok = np.where(D <= 0)[0]
    if len(ok) == 0:
        return int(np.argmin(C))
    return int(ok[np.argmin(C[ok])])


class KernelControlFwdOracle:
    """`KernelControlFwd.train/__call__` (`gym_socks/algorithms/control/kernel_control_fwd.py:
    164-238`), heuristic solution.  Scores C = f32(cost_t(Y)) @ (W @ (K(X,s) * CUA)) (`:220-225`),
    optional D (`:228-232`), action = f32(A[argmin]) (`:236-238`)."""

    def __init__(self, cost_fn, constraint_fn=None, sigma=0.1, regularization_param=1):
        self.cost_fn, self.constraint_fn = cost_fn, constraint_fn
        self.sigma, self.lam = sigma, regularization_param

    def train(self, X, U, Y, A):
        self.X = np.asarray(X, dtype=np.float64)
        self.Y = np.asarray(Y, dtype=np.float64)
        self.A = np.asarray(A, dtype=np.float64)
        U = np.asarray(U, dtype=np.float64)
        self.W = regularized_inverse(self.X, U=U, regularization_param=self.lam, sigma=self.sigma)
        self.CUA = rbf_kernel(U, self.A, self.sigma)
        return self

    def scores(self, time, state):
        CXT = rbf_kernel(self.X, np.asarray(state, dtype=np.float64).reshape(1, -1), self.sigma)
        beta = self.W @ (CXT * self.CUA)
        C = np.matmul(np.array(self.cost_fn(time=time, state=self.Y), dtype=np.float32), beta)
        D = None
        if self.constraint_fn is not None:
            D = np.matmul(np.array(self.constraint_fn(time=time, state=self.Y), dtype=np.float32), beta)
        return C, D

    def __call__(self, time=0, state=None):
        C, D = self.scores(time, state)
        return np.array(self.A[heuristic_solution(C, D)], dtype=np.float32)


class KernelControlBwdOracle:
    """`KernelControlBwd.train/__call__` (`gym_socks/algorithms/control/kernel_control_bwd.py:58-122, 351-440`).

    Recursion (`:70-120`): out[T-1] = f32(cost_{T-1}(Y)); for t = T-2..0: Z = out[t+1] @ W (`:90`),
    C[j,k] = sum_i Z_i CXY[i,j] CUA[i,k] (`:67,91`: einsum over beta = CXY (x) CUA), optional
    D from R = constraint_t(Y) @ W (`:96-97`), V_j = min_k C[j,k] over {D[j,k] <= 0} (all k if empty, `:99-107`),
    out[t] = cost_t(Y) + V (`:115`).  The N x N x P tensor `beta` is never formed: C = CXY^T diag(Z) CUA."""

    def __init__(self, time_horizon, cost_fn, constraint_fn=None, sigma=0.1, regularization_param=1):
        self.T, self.cost_fn, self.constraint_fn = time_horizon, cost_fn, constraint_fn
        self.sigma, self.lam = sigma, regularization_param

    def train(self, X, U, Y, A):
        self.X = np.asarray(X, dtype=np.float64)
        self.Y = np.asarray(Y, dtype=np.float64)
        self.A = np.asarray(A, dtype=np.float64)
        U = np.asarray(U, dtype=np.float64)
        self.W = regularized_inverse(self.X, U=U, regularization_param=self.lam, sigma=self.sigma)
        self.CUA = rbf_kernel(U, self.A, self.sigma)
        CXY = rbf_kernel(self.X, self.Y, self.sigma)
        n, T = len(self.Y), self.T
        out = np.empty((T, n))
        out[T - 1, :] = np.array(self.cost_fn(time=T - 1, state=self.Y), dtype=np.float32)
        for t in range(T - 2, -1, -1):
            Z = np.matmul(out[t + 1, :], self.W)
            C = CXY.T @ (Z[:, None] * self.CUA)
            if self.constraint_fn is not None:
                R = np.matmul(self.constraint_fn(time=t, state=self.Y), self.W)
                D = CXY.T @ (R[:, None] * self.CUA)
                V = np.empty(n)
                for i in range(n):
                    CA = C[i][D[i] <= 0]
                    V[i] = np.min(C[i]) if CA.size == 0 else np.min(CA)
            else:
                V = np.min(C, axis=1)
            out[t, :] = self.cost_fn(time=t, state=self.Y) + V
        self.value_functions = out
        return self

    def scores(self, time, state):
        CXT = rbf_kernel(self.X, np.asarray(state, dtype=np.float64).reshape(1, -1), self.sigma)
        beta = self.W @ (CXT * self.CUA)
        C = np.matmul(np.array(self.value_functions[time], dtype=np.float32), beta)
        D = None
        if self.constraint_fn is not None:
            D = np.matmul(np.array(self.constraint_fn(time=time, state=self.Y), dtype=np.float32), beta)
        return C, D

    def __call__(self, time=0, state=None):
        C, D = self.scores(time, state)
        return np.array(self.A[heuristic_solution(C, D)], dtype=np.float32)


# --------------------------------------------------------------------------- abel kernel / separating classifier
def abel_kernel(X, Y=None, sigma=None):
    """exp(-|x - y| / sigma); sigma=None uses the median of the (unsquared) distances
    (`gym_socks/kernel/metrics.py:174-206`, distances from `euclidean_distance` `:26-66`)."""
    if Y is None:
        Y = X
    K = np.sqrt(sq_dist(X, Y))
    if sigma is None:
        sigma = np.median(K)
    else:
        assert sigma > 0
    K /= -sigma
    np.exp(K, K)
    return K


class SeparatingKernelOracle:
    """`SeparatingKernelClassifier.fit/predict` (`gym_socks/algorithms/reach/separating_kernel.py:67-111`):
    W = inv(K + lambda N I) with the Abel kernel, tau = 1 - min diag(K^T W K / n), predict diag(K_T^T W K_T / n) >= 1 - tau."""

    def __init__(self, sigma=0.1, regularization_param=1):
        self.sigma, self.lam = sigma, regularization_param

    def fit(self, X):
        X = np.asarray(X, dtype=np.float64)
        self.X = X
        n = len(X)
        K = abel_kernel(X, X, self.sigma)
        G = K.copy()
        G[np.diag_indices(n)] += self.lam * n
        self.W = np.linalg.inv(G)
        self.tau = 1 - np.min(np.diagonal((1 / n) * K.T @ self.W @ K))
        return self

    def scores(self, T):
        K = abel_kernel(self.X, np.asarray(T, dtype=np.float64), self.sigma)
        return np.einsum("ij,ij->j", K, self.W @ K) / len(self.X)

    def predict(self, T):
        return self.scores(T) >= 1 - self.tau


# --------------------------------------------------------------------------- kernel/probability.py, algorithms/kernel.py
def maximum_mean_discrepancy(X, Y, sigma=1.0, biased=False, squared=False):
    """`gym_socks/kernel/probability.py:9-73` with the rbf kernel."""
    m, n = len(X), len(Y)
    GX, GY, GXY = rbf_kernel(X, X, sigma), rbf_kernel(Y, Y, sigma), rbf_kernel(X, Y, sigma)
    if biased:
        mmd = GX.sum() / m ** 2 + GY.sum() / n ** 2 - 2 * GXY.sum() / (m * n)
    else:
        mmd = ((GX.sum() - np.trace(GX)) / (m * (m - 1)) + (GY.sum() - np.trace(GY)) / (n * (n - 1))
               - 2 * GXY.sum() / (m * n))
    return mmd if squared else np.sqrt(mmd)


def witness_function(X, Y, t, sigma=1.0):
    """`gym_socks/kernel/probability.py:76-93`."""
    return rbf_kernel(X, t, sigma).sum(axis=0) / len(X) - rbf_kernel(Y, t, sigma).sum(axis=0) / len(Y)


def conditional_embedding_predict(G, lam, y, K):
    """`ConditionalEmbedding.fit/predict` (`gym_socks/algorithms/kernel.py:43-81`): (y^T solve(G + lam M I, K))^T."""
    W = G + lam * len(G) * np.eye(len(G))
    return (np.asarray(y).T @ np.linalg.solve(W, K)).T
