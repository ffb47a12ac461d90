"""Evolutionary models -> linear-Gaussian factors.  Restates src/evomodels/*.jl.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Every model exposes
    ntraits, mu (root prior mean vector), v (root prior variance: scalar for
    univariate models, vector for MvDiag, matrix otherwise)
    isrootfixed()                          src/evomodels/evomodels.jl:41
    factor_treeedge(edge) -> (h, J, g)     child block first, then parent block
    factor_hybridnode(parent_edges)        child block, then parents in the
                                           order the edges are given
    factor_root()                          src/evomodels/evomodels.jl:377-396
Degenerate branches (t == 0, GeneralizedBelief) are out of scope
 and raise.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

LOG2PI = math.log(2.0 * math.pi)


def _sym_logdet(a: np.ndarray) -> float:
    sign, ld = np.linalg.slogdet(a)
    if sign <= 0:
        raise np.linalg.LinAlgError("matrix is not positive definite")
    return float(ld)


def factor_generic(q: np.ndarray, omega: Optional[np.ndarray], j: np.ndarray, nparents: int,
                   ntraits: int, g0: float):
    """src/evomodels/evomodels.jl:214-245:  X0 | X_pa ~ N(q x_pa + omega, j^-1).
    J = [j, -j q; -q'j, q'j q];  h = [j w; -q'j w];  g = g0 - w'j w/2."""
    jq = -j @ q
    qjq = -q.T @ jq
    J = np.block([[j, jq], [jq.T, qjq]])
    ntot = ntraits * (1 + nparents)
    if omega is None:
        return np.zeros(ntot), J, float(g0)
    jomega = j @ omega
    h = np.concatenate([jomega, jq.T @ omega])
    g = g0 - float(omega @ jomega) / 2
    return h, J, float(g)


class EvolutionaryModel:
    ntraits: int

    def isrootfixed(self):
        return bool(np.all(np.asarray(self.v) == 0))

    def rootpriormeanvector(self):
        return np.atleast_1d(np.asarray(self.mu, dtype=float))

    def rootpriorvariance(self):
        v = np.asarray(self.v, dtype=float)
        if v.ndim == 0:
            return v.reshape(1, 1)
        if v.ndim == 1:
            return np.diag(v)
        return v

    def factor_root(self):
        """src/evomodels/evomodels.jl:379-396."""
        V = self.rootpriorvariance()
        mu = self.rootpriormeanvector()
        p = self.ntraits
        if np.any(np.isinf(np.diag(V))):
            return np.zeros(p), np.zeros((p, p)), 0.0
        j = np.linalg.inv(V)
        h = j @ mu
        g = (-p * LOG2PI + _sym_logdet(j) - float(mu @ h)) / 2
        return h, j, float(g)

    # generic fall-backs (src/evomodels/evomodels.jl:208-211, 314-330)
    def branch_q_omega_j_g(self, edge):
        raise NotImplementedError

    def branch_q_omega_v(self, edge):
        raise NotImplementedError

    def factor_treeedge(self, edge):
        q, om, j, g0 = self.branch_q_omega_j_g(edge)
        return factor_generic(q, om, j, 1, self.ntraits, g0)

    def factor_hybridnode(self, pae):
        p = self.ntraits
        v = np.zeros((p, p))
        om = np.zeros(p)
        q = np.zeros((p, p * len(pae)))
        for k, e in enumerate(pae):
            qe, oe, ve = self.branch_q_omega_v(e)
            q[:, k * p:(k + 1) * p] = e.gamma * qe
            v += e.gamma ** 2 * ve
            om += e.gamma * oe
        j = np.linalg.inv(v)
        g0 = (-p * LOG2PI + _sym_logdet(j)) / 2
        return factor_generic(q, om, j, len(pae), p, g0)


# --------------------------------------------------------------------------
# homogeneous Brownian motion (src/evomodels/homogeneousbrownianmotion.jl)
# --------------------------------------------------------------------------
class _HomogeneousBM(EvolutionaryModel):
    """R: p x p variance rate; P = R^-1; g0 = -(p log2pi + logdet R)/2."""

    def _setup(self, R, mu, v):
        self.R = np.atleast_2d(np.asarray(R, dtype=float))
        self.ntraits = self.R.shape[0]
        self.P = np.linalg.inv(self.R)
        self.g0 = -(self.ntraits * LOG2PI + _sym_logdet(self.R)) / 2
        self.mu = np.atleast_1d(np.asarray(mu, dtype=float))
        self.v = v

    def factor_treeedge(self, edge):
        """:222-282 (non-degenerate branch): J=[j -j; -j j], j=P/t, h=0,
        g = g0 - p log(t)/2."""
        t = edge.length
        if t == 0:
            raise NotImplementedError("degenerate (t=0) tree edge: out of scope")
        j = self.P / t
        J = np.block([[j, -j], [-j, j]])
        return np.zeros(2 * self.ntraits), J, float(self.g0 - self.ntraits * math.log(t) / 2)

    def factor_hybridnode(self, pae):
        """:288-351: t0 = sum gamma^2 t; c=(1,-gammas); J_uv = c_u c_v P/t0."""
        gam = np.array([e.gamma for e in pae], dtype=float)
        t = np.array([e.length for e in pae], dtype=float)
        t0 = float(np.sum(gam ** 2 * t))
        if t0 == 0:
            raise NotImplementedError("degenerate hybrid: out of scope")
        c = np.concatenate([[1.0], -gam])
        J = np.kron(np.outer(c, c), self.P / t0)
        return (np.zeros(self.ntraits * len(c)), J,
                float(self.g0 - self.ntraits * math.log(t0) / 2))


class UnivariateBrownianMotion(_HomogeneousBM):
    """:16-49.  v=None => fixed root (0)."""

    def __init__(self, sigma2, mu, v=None):
        v = 0.0 if v is None else float(v)
        if v < 0:
            raise ValueError("root variance v must be non-negative")
        self._setup([[float(sigma2)]], [float(mu)], v)
        self.sigma2 = float(sigma2)

    def factor_root(self):
        """:379-384 (univariate)."""
        v = float(self.v)
        m = float(self.mu[0])
        j = 0.0 if math.isinf(v) else 1.0 / v
        g = 0.0 if j == 0.0 else -(LOG2PI + math.log(v) + m * m * j) / 2
        return np.array([m * j]), np.array([[j]]), g


class MvDiagBrownianMotion(_HomogeneousBM):
    """:60-91."""

    def __init__(self, R, mu, v=None):
        R = np.asarray(R, dtype=float)
        mu = np.asarray(mu, dtype=float)
        if R.shape != mu.shape:
            raise ValueError("R and mu have different lengths")
        if not np.all(R > 0):
            raise ValueError("evolutionary variance rates R must all be positive")
        v = np.zeros_like(mu) if v is None else np.asarray(v, dtype=float)
        if v.shape != mu.shape or np.any(v < 0):
            raise ValueError("bad root variance")
        self._setup(np.diag(R), mu, v)


class MvFullBrownianMotion(_HomogeneousBM):
    """:101-128."""

    def __init__(self, R, mu, v=None):
        R = np.asarray(R, dtype=float)
        mu = np.asarray(mu, dtype=float)
        p = mu.size
        if R.shape != (p, p):
            raise ValueError("R and mu have conflicting sizes")
        if not np.array_equal(R, R.T):
            raise ValueError("R should be symmetric")
        v = np.zeros((p, p)) if v is None else np.asarray(v, dtype=float)
        np.linalg.cholesky(R)  # `inv(R)` in the reference fails if not PD
        self._setup(R, mu, v)


# --------------------------------------------------------------------------
# heterogeneous BM (src/evomodels/heterogeneousmodels.jl:70-150)
# --------------------------------------------------------------------------
class HeterogeneousBrownianMotion(EvolutionaryModel):
    """Per-edge 'painted' rate: colors maps edge.number -> 1-based colour,
    default colour 1 (PaintedParameter, :21-33)."""

    def __init__(self, rates: Sequence[np.ndarray], colors: Optional[Dict[int, int]], mu, v=None):
        rates = [np.atleast_2d(np.asarray(R, dtype=float)) for R in rates]
        self.rates = rates
        self.colors = dict(colors or {})
        self.mu = np.atleast_1d(np.asarray(mu, dtype=float))
        p = self.ntraits = self.mu.size
        for R in rates:
            if R.shape != (p, p) or not np.array_equal(R, R.T):
                raise ValueError("R and mu have conflicting sizes / R not symmetric")
        self.v = np.zeros((p, p)) if v is None else np.asarray(v, dtype=float)
        self.invrates = [np.linalg.inv(R) for R in rates]
        self.g0 = [-(p * LOG2PI + _sym_logdet(R)) / 2 for R in rates]

    def color(self, edge):
        return self.colors.get(edge.number, 1) - 1

    def factor_treeedge(self, edge):
        """:128-134 -> generic with q=I, omega=0, j = R_c^-1 / t."""
        p = self.ntraits
        c = self.color(edge)
        j = self.invrates[c] / edge.length
        g = self.g0[c] - p * math.log(edge.length) / 2
        return factor_generic(np.eye(p), None, j, 1, p, g)

    def factor_hybridnode(self, pae):
        """:135-150: Sigma = sum gamma^2 t R_c; j = Sigma^-1; q=[gamma_k I]."""
        p = self.ntraits
        v = np.zeros((p, p))
        q = np.zeros((p, p * len(pae)))
        for k, e in enumerate(pae):
            q[:, k * p:(k + 1) * p] = e.gamma * np.eye(p)
            v += e.gamma ** 2 * (e.length * self.rates[self.color(e)])
        j = np.linalg.inv(v)
        g0 = (-p * LOG2PI + _sym_logdet(j)) / 2
        return factor_generic(q, None, j, len(pae), p, g0)


# --------------------------------------------------------------------------
# univariate OU (src/evomodels/homogeneousornsteinuhlenbeck.jl:18-66)
# --------------------------------------------------------------------------
class UnivariateOrnsteinUhlenbeck(EvolutionaryModel):
    def __init__(self, sigma2, alpha, theta, mu, v=None):
        if sigma2 <= 0 or alpha <= 0:
            raise ValueError("sigma2 and alpha must be positive")
        self.ntraits = 1
        self.gamma2 = sigma2 / (2 * alpha)
        self.alpha, self.theta = float(alpha), float(theta)
        self.mu = np.array([float(mu)])
        self.v = 0.0 if v is None else float(v)
        self.g0 = -(LOG2PI + math.log(self.gamma2)) / 2

    factor_root = UnivariateBrownianMotion.factor_root

    def branch_q_omega_j_g(self, edge):
        q = math.exp(-self.alpha * edge.length)
        fac = 1 - q * q
        j = 1 / self.gamma2 / fac
        om = (1 - q) * self.theta
        g = self.g0 - math.log(fac) / 2
        return np.array([[q]]), np.array([om]), np.array([[j]]), g

    def branch_q_omega_v(self, edge):
        a = math.exp(-self.alpha * edge.length)
        fac = 1 - a * a
        return np.array([[a]]), np.array([(1 - a) * self.theta]), np.array([[self.gamma2 * fac]])
