"""Pure (h, J, g) canonical-form numerics.  Restates src/beliefupdates.jl:11-83,
187-200 and the PDMats 0.11 operations they call (PDMats is a third-party
dependency, compat-bounded in Project.toml:28 and not vendored: `PDMat(S)` =
dense Cholesky of the upper triangle, `X_invA_Xt(a,x)` = (x/U)(x/U)', `a\\x`,
`logdet(a)` = 2 sum(log U_ii)).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np

LOG2PI = math.log(2.0 * math.pi)
EPS = float(np.finfo(np.float64).eps)


class BPPosDefException(Exception):
    """src/beliefupdates.jl:11-22.  info = 1-based failing pivot (LAPACK)."""

    def __init__(self, msg, info):
        super().__init__(msg)
        self.msg, self.info = msg, info


def chol_upper(A):
    """Upper Cholesky factor U (A = U'U) reading only A's upper triangle, as
    LAPACK dpotrf('U') does; raises BPPosDefException(info) at the first
    non-positive (or NaN) pivot."""
    n = A.shape[0]
    U = np.zeros((n, n))
    for j in range(n):
        d = A[j, j] - float(U[:j, j] @ U[:j, j])
        if not (d > 0.0):
            raise BPPosDefException("matrix is not positive definite", j + 1)
        U[j, j] = math.sqrt(d)
        if j + 1 < n:
            U[j, j + 1:] = (A[j, j + 1:] - U[:j, j] @ U[:j, j + 1:]) / U[j, j]
    return U


def _solve_ut(U, b):
    """U' x = b (forward substitution), U upper."""
    from scipy.linalg import solve_triangular
    return solve_triangular(U, b, trans="T", lower=False, check_finite=False)


def _solve_u(U, b):
    from scipy.linalg import solve_triangular
    return solve_triangular(U, b, lower=False, check_finite=False)


def marginalize(h, J, g, keep_index, integrate_index=None, metadata=""):
    """src/beliefupdates.jl:51-83.  0-based index arrays."""
    keep_index = np.asarray(keep_index, dtype=np.int64)
    if integrate_index is None:
        integrate_index = np.setdiff1d(np.arange(h.size), keep_index)
    integrate_index = np.asarray(integrate_index, dtype=np.int64)
    if integrate_index.size == 0:
        return h, J, g
    Ji = J[np.ix_(integrate_index, integrate_index)]
    Jk = J[np.ix_(keep_index, keep_index)]
    Jki = J[np.ix_(keep_index, integrate_index)]
    hi = h[integrate_index]
    hk = h[keep_index]
    if np.all(np.abs(Ji) <= EPS) and np.all(np.abs(hi) <= EPS) and np.all(np.abs(Jki) <= EPS):
        return hk, Jk, g
    try:
        U = chol_upper(Ji)
    except BPPosDefException as ex:
        raise BPPosDefException(f"belief {metadata}, integrating {list(integrate_index + 1)}", ex.info)
    if keep_index.size:
        Z = _solve_ut(U, Jki.T).T  # Jki / U
        messageJ = Jk - Z @ Z.T
    else:
        messageJ = Jk
    mui = _solve_u(U, _solve_ut(U, hi))
    messageh = hk - Jki @ mui
    ni = integrate_index.size
    logdet = 2.0 * float(np.sum(np.log(np.diag(U))))
    messageg = g + (ni * LOG2PI - logdet + float(hi @ mui)) / 2
    return messageh, messageJ, messageg


def integratebelief(h, J, g):
    """src/beliefupdates.jl:187-200 -> (mu, norm)."""
    if not np.any(h) and not np.any(J):
        return np.full(h.shape, np.inf), g
    U = chol_upper(J)
    mu = _solve_u(U, _solve_ut(U, h))
    n = h.size
    logdet = 2.0 * float(np.sum(np.log(np.diag(U))))
    return mu, g + (n * LOG2PI - logdet + float(h @ mu)) / 2
