"""LP / heuristic selection of the control action, host side.

`compute_solution` mirrors `gym_socks/algorithms/control/common.py:25-174`.  The LP (`scipy.optimize.linprog`
over P variables) is third-party host code in the reference and stays on the host here (SURVEY §2 row 4:
out of scope for the GPU path); the heuristic branches used with `heuristic=True` run on the device
(`socks_argmin_masked`), this numpy version is the fallback the reference takes when the LP fails.
"""
import logging

import numpy as np

logger = logging.getLogger(__name__)


def _heuristic(C, D=None):
    sol = np.zeros((len(C),))
    if D is None:
        idx = np.argmin(C)  # common.py:96-100
    else:
        ok = np.where(D <= 0)  # common.py:166-174
        idx = C.argmin() if len(ok[0]) == 0 else ok[0][C[ok].argmin()]
    sol[idx] = 1
    return sol


def compute_solution(C, D=None, heuristic=False):
    """Probability vector over the admissible actions (common.py:25-43)."""
    if heuristic is False:
        from scipy.optimize import linprog

        kw = dict(A_eq=np.ones((1, len(C))), b_eq=1)
        if D is not None:
            Dm = D.reshape(-1, 1) if D.ndim == 1 else D
            kw.update(A_ub=Dm.T, b_ub=0)  # common.py:140-151
        sol = linprog(C.T, **kw)  # common.py:77-81
        if sol.success is True:
            return sol.x
        logger.debug("No solution found via scipy.optimize.linprog; returning heuristic solution.")
    return _heuristic(C, D)
